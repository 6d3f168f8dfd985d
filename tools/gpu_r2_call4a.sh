#!/bin/bash
# round 2, GPU call 4a: TMA-staged prefetch (cp.async.bulk) variants of k_accumulate against the default, parity + timing
mkdir -p gpurun_out
T=gpurun_out/r2c4a
for lib in tma tma4; do
  ( GMSM_LIB=$lib timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_tables.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3 ) | tee -a ${T}_pytest_tma.log
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
    print("lib=%-8s %-12s 2^%s %s %.3f ms parity=%s stages %s" % (lib,curve,logn," ".join(extra),d["ms_per_step"],d["parity"]["closed_form"],{k:round(v,2) for k,v in d["stages_ms"].items()}))
except Exception as e:
    print("lib=%s %s 2^%s FAILED: %s" % (lib,curve,logn,(r.stderr or r.stdout)[-300:]))
PY
}
for lib in default tma tma4; do ab $lib bn254_g1 24; ab $lib bn254_g1 20; done | tee ${T}_ab_tma.txt
for lib in default tma; do ab $lib bls12381_g1 24; ab $lib bn254_g2 22; ab $lib bls12381_g2 20; done | tee -a ${T}_ab_tma.txt
for lib in default tma4; do
  GMSM_LIB=$( [ $lib = default ] && echo "" || echo $lib ) timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 3 -c 1 -f -o /tmp/prof_$lib \
      python bench.py --logn 24 --steps 1 --warmup 1 --no-e2e --no-cpu --no-tables --no-extras > ${T}_prof_$lib.log 2>&1
  ncu -i /tmp/prof_$lib.ncu-rep --page raw --csv > ${T}_ncu_bn254_g1_${lib}_raw.csv 2>/dev/null
done
ls -la gpurun_out | grep r2c4a
