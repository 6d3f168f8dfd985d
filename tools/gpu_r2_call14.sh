#!/bin/bash
# round 2, GPU call 14: checks of the refitted window model on the N4 curves (bls24-315 at c = 20), lane-parallel tail for the
# 20-limb bw6-633, the defaults of secp256k1 / bw6-761 after the model and tail changes
mkdir -p gpurun_out
T=gpurun_out/r2c14
ab() {
  python - "$@" <<'PY'
import json,sys,subprocess,os
curve,logn=sys.argv[1:3]; extra=sys.argv[3:]
env=dict(os.environ); args=[]
for kv in extra:
    k,v=kv.split("=")
    if k=="c": args+=["--c",v]
    else: env[k]=v
r=subprocess.run(["python","bench.py","--curve",curve,"--logn",logn,"--steps","5","--warmup","3","--no-cpu","--no-e2e","--no-tables","--no-extras"]+args,capture_output=True,text=True,env=env)
try:
    d=json.loads(r.stdout.strip().splitlines()[-1])
    print("%-12s 2^%s %-32s c=%d %.3f ms parity=%s stages %s" % (curve,logn," ".join(extra),d["engine"]["c"],d["ms_per_step"],d["parity"]["closed_form"],{k:round(v,2) for k,v in d["stages_ms"].items()}))
except Exception as e:
    print("%s 2^%s %s FAILED: %s" % (curve,logn,extra,(r.stderr or r.stdout)[-400:]))
PY
}
{
ab bls24315_g1 24; ab bls24315_g1 24 c=17; ab bls24315_g1 20; ab bls24315_g1 20 c=16; ab bls24317_g1 22
ab bw6633_g1 22; ab bw6633_g1 22 GMSM_QUAD=1 GMSM_QUAD_MAX=20000; ab bw6633_g1 18; ab bw6633_g1 18 GMSM_QUAD=1 GMSM_QUAD_MAX=20000
ab secp256k1_g1 24; ab bw6761_g1 22; ab bw6761_g1 18; ab bn254_g1 24; ab bn254_g2 22
} | tee ${T}_sweeps.txt
