#!/bin/bash
# round 2, GPU call 7: k_accumulate without the software prefetch (GMSM_ACC_NOPREFETCH=1: latency covered by the other warps) at 4 / 3
# blocks per SM against the default
mkdir -p gpurun_out
T=gpurun_out/r2c7
for lib in np4; do
  ( GMSM_LIB=$lib timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_tables.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2 ) | tee -a ${T}_pytest_np.log
done
ab() {
  python - "$@" <<'PY'
import json,sys,subprocess,os
lib,curve,logn=sys.argv[1:4]; extra=sys.argv[4:]
env=dict(os.environ); env["GMSM_LIB"]=lib if lib!="default" else ""
args=[]
for kv in extra:
    k,v=kv.split("=")
    if k=="c": args+=["--c",v]
    else: env[k]=v
r=subprocess.run(["python","bench.py","--curve",curve,"--logn",logn,"--steps","5","--warmup","3","--no-cpu","--no-e2e","--no-tables","--no-extras"]+args,capture_output=True,text=True,env=env)
try:
    d=json.loads(r.stdout.strip().splitlines()[-1])
    print("lib=%-8s %-12s 2^%s %s c=%d %.3f ms parity=%s stages %s" % (lib,curve,logn," ".join(extra),d["engine"]["c"],d["ms_per_step"],d["parity"]["closed_form"],{k:round(v,2) for k,v in d["stages_ms"].items()}))
except Exception as e:
    print("lib=%s %s 2^%s FAILED: %s" % (lib,curve,logn,(r.stderr or r.stdout)[-300:]))
PY
}
for lib in default np4 np3; do ab $lib bn254_g1 24; ab $lib bn254_g1 20; ab $lib bls12381_g1 24; ab $lib bn254_g2 22; ab $lib bls12381_g2 20; done | tee ${T}_ab_noprefetch.txt
for lib in g2lazy g2lazys; do
  ( GMSM_LIB=$lib timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_ops.py -x -q -m gpu -p no:cacheprovider -k "g2 or G2 or ops or golden or all_window" 2>&1 | tail -2 ) | tee -a ${T}_pytest_g2lazy.log
done
for lib in default g2dot g2lazy g2lazys; do ab $lib bn254_g2 22; ab $lib bls12381_g2 20; done | tee ${T}_ab_g2lazy.txt
ls -la gpurun_out | grep r2c7
