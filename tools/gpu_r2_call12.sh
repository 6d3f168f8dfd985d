#!/bin/bash
# round 2, GPU call 12: N4 curves -- wider windows for secp256k1 (19, 20), lane-parallel tail (GMSM_QUAD) for the 24-limb bw6-761,
# parity of the binary-GCD inversion on the full-width moduli
mkdir -p gpurun_out
T=gpurun_out/r2c12
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
    print("%-12s 2^%s %-12s c=%d %.3f ms parity=%s stages %s" % (curve,logn," ".join(extra),d["engine"]["c"],d["ms_per_step"],d["parity"]["closed_form"],{k:round(v,2) for k,v in d["stages_ms"].items()}))
except Exception as e:
    print("%s 2^%s %s FAILED: %s" % (curve,logn,extra,(r.stderr or r.stdout)[-400:]))
PY
}
{
for c in 16 19 20; do ab secp256k1_g1 24 c=$c; done
ab secp256k1_g1 22; ab secp256k1_g1 20; ab secp256k1_g1 16
for q in 0 1; do ab bw6761_g1 22 GMSM_QUAD=$q; ab bw6761_g1 18 GMSM_QUAD=$q; ab bw6761_g1 14 GMSM_QUAD=$q; done
for q in 0 1; do ab bls12381_g2 20 GMSM_QUAD=$q; done
ab bw6761_g1 20; ab bw6761_g1 24
} | tee ${T}_sweeps.txt
ls -la gpurun_out | grep r2c12
