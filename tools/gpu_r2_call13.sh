#!/bin/bash
# round 2, GPU call 13: bls24-315 / bls24-317 G1 and bw6-633 G1 / G2 on hardware (10- and 20-limb fields, 40-byte scalars):
# parity, first timings and width sweeps; the other N4 curves again after the window-model / lane-parallel-tail changes
mkdir -p gpurun_out
T=gpurun_out/r2c13
( time timeout 1200 python -m pytest tests/test_gpu_n4_curves.py tests/test_gpu_ops.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 ) 2>&1 | tee ${T}_pytest.log
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
    print("%-12s 2^%s %-6s c=%d %.3f ms parity=%s stages %s" % (curve,logn," ".join(extra),d["engine"]["c"],d["ms_per_step"],d["parity"]["closed_form"],{k:round(v,2) for k,v in d["stages_ms"].items()}))
except Exception as e:
    print("%s 2^%s %s FAILED: %s" % (curve,logn,extra,(r.stderr or r.stdout)[-400:]))
PY
}
{
for c in 15 16 17 18 19; do ab bls24315_g1 24 c=$c; done
for c in 13 14 15 16; do ab bls24315_g1 20 c=$c; done
ab bls24317_g1 24; ab bls24317_g1 20
for c in 15 16 17 18; do ab bw6633_g1 22 c=$c; done
for c in 12 13 14 15; do ab bw6633_g1 18 c=$c; done
ab bw6633_g2 20
ab secp256k1_g1 24; ab bw6761_g1 22; ab bw6761_g1 18
} | tee ${T}_sweeps.txt
ls -la gpurun_out | grep r2c13
