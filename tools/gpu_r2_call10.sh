#!/bin/bash
# round 2, GPU call 10: four-product fused reduction for the G2 y-coordinate (GMSM_DOT4) against the default
mkdir -p gpurun_out
T=gpurun_out/r2c10
( GMSM_LIB=g2d4 timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_ops.py tests/test_gpu_tables.py -x -q -m gpu -p no:cacheprovider -k "g2 or G2 or ops or golden or all_window or table" 2>&1 | tail -2 ) | tee ${T}_pytest_dot4.log
ab() {
  python - "$@" <<'PY'
import json,sys,subprocess,os
lib,curve,logn=sys.argv[1:4]
env=dict(os.environ); env["GMSM_LIB"]=lib if lib!="default" else ""
r=subprocess.run(["python","bench.py","--curve",curve,"--logn",logn,"--steps","5","--warmup","3","--no-cpu","--no-e2e","--no-tables","--no-extras"],capture_output=True,text=True,env=env)
try:
    d=json.loads(r.stdout.strip().splitlines()[-1])
    print("lib=%-8s %-12s 2^%s c=%d %.3f ms parity=%s stages %s" % (lib,curve,logn,d["engine"]["c"],d["ms_per_step"],d["parity"]["closed_form"],{k:round(v,2) for k,v in d["stages_ms"].items()}))
except Exception as e:
    print("lib=%s %s 2^%s FAILED: %s" % (lib,curve,logn,(r.stderr or r.stdout)[-300:]))
PY
}
for lib in default g2d4; do ab $lib bn254_g2 22; ab $lib bls12381_g2 20; ab $lib bls12377_g2 20; done | tee ${T}_ab_dot4.txt
for th in 16 32 48; do
GMSM_COPY_THREADS=$th timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-tables --no-extras 2>>${T}.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('GMSM_COPY_THREADS=$th  e2e pinned', round(d['e2e']['ms_per_step'],2), 'e2e_pageable', round(d['e2e_pageable']['ms_per_step'],2))" | tee -a ${T}_copy_threads.txt
done
