#!/bin/bash
# round 2, GPU call 19: k_accumulate of the 10- / 12-limb G1 groups at 3 blocks per SM (168 registers, small spills) against the
# default 2 blocks (190 / 232 registers): parity of the variant build, then the A/B
mkdir -p gpurun_out
T=gpurun_out/r2c19
( GMSM_LIB=b3 timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_n4_curves.py -x -q -m gpu -p no:cacheprovider -k "bls12381_g1 or bls12377_g1 or bls24315 or bls24317" 2>&1 | tail -3 ) | tee ${T}_pytest_b3.log
ab() {
  python - "$@" <<'PY'
import json,sys,subprocess,os
lib,curve,logn=sys.argv[1:4]
env=dict(os.environ)
if lib!="default": env["GMSM_LIB"]=lib
r=subprocess.run(["python","bench.py","--curve",curve,"--logn",logn,"--steps","5","--warmup","3","--no-cpu","--no-e2e","--no-tables","--no-extras"],capture_output=True,text=True,env=env)
try:
    d=json.loads(r.stdout.strip().splitlines()[-1])
    print("lib=%-8s %-12s 2^%s c=%d %.3f ms parity=%s accumulate %.2f" % (lib,curve,logn,d["engine"]["c"],d["ms_per_step"],d["parity"]["closed_form"],d["stages_ms"]["accumulate"]))
except Exception as e:
    print("%s %s 2^%s FAILED: %s" % (lib,curve,logn,(r.stderr or r.stdout)[-300:]))
PY
}
for cfg in "bls12381_g1 24" "bls12377_g1 22" "bls24315_g1 24" "bls24317_g1 22"; do
  set -- $cfg
  ab default $1 $2; ab b3 $1 $2
done | tee ${T}_ab_minblocks3.txt
