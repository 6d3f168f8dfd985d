#!/bin/bash
# round 2, GPU call 15: c = 20 / 21 against the current choices for the 12-limb G1 groups at 2^24 (their c = 19 has a narrow last
# window: K1 takes 5 ms) and for bn254 G1 at the shard sizes of the multi-GPU runs
mkdir -p gpurun_out
T=gpurun_out/r2c15
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
    print("%-12s 2^%s %-8s c=%d %.3f ms parity=%s stages %s" % (curve,logn," ".join(extra),d["engine"]["c"],d["ms_per_step"],d["parity"]["closed_form"],{k:round(v,2) for k,v in d["stages_ms"].items()}))
except Exception as e:
    print("%s 2^%s %s FAILED: %s" % (curve,logn,extra,(r.stderr or r.stdout)[-400:]))
PY
}
{
for c in 19 20 21; do ab bls12381_g1 24 c=$c; done
for c in 17 19 20; do ab bls12377_g1 24 c=$c; done
for c in 17 20; do ab bn254_g1 25 c=$c; done
ab bn254_g1 24 c=20; ab bls12381_g1 23 c=20; ab bls12381_g1 23 c=17; ab bn254_g2 24 c=17; ab bn254_g2 24 c=20
} | tee ${T}_sweeps.txt
