#!/bin/bash
# round 2, GPU call 8: the new default (per-group choices: no-prefetch, Fp2 by fused products) -- parity, full bench line, A/B of the
# Karatsuba product and of the bn254-G2 choices on the bls G2 groups, ncu captures of the final kernels
mkdir -p gpurun_out
T=gpurun_out/r2c8
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5 ) 2>&1 | tee ${T}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > ${T}_bench_full.json 2> ${T}_bench_full.err; echo "bench full rc=$?"; tail -3 ${T}_bench_full.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c8_bench_full.json"))
    print({k:d[k] for k in ("value","ms_per_step")}, d["parity"]["closed_form"], d["engine"]["c"], d["roofline"]["frac"], d["roofline"]["int_pipe"]); print({k:round(v,2) for k,v in d["stages_ms"].items()})
    for k in ("e2e","e2e_pageable","e2e_resident_bases","e2e_resident_tables","cpu_baseline","concurrent3","resident_tables"):
        v=d.get(k); print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a not in ("sample","path","what","stages_ms","note","workload")} if v else None)
    for k,v in (d.get("configs") or {}).items():
        print(k, round(v["ms_per_step"],3), v["parity"]["closed_form"], "c=%d"%v["c"], {a:round(b,2) for a,b in v["stages_ms"].items()}, v.get("vs_random_scalars"))
except Exception as e:
    print("no full line:", e)
PY
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
( GMSM_LIB=kara timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_ops.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2 ) | tee ${T}_pytest_kara.log
for lib in default kara; do ab $lib bn254_g1 24; ab $lib bls12381_g1 24; done | tee ${T}_ab_karatsuba.txt
for lib in default g2x; do ab $lib bls12381_g2 20; ab $lib bls12377_g2 20; ab $lib bn254_g2 22; done | tee ${T}_ab_g2x.txt
ab default bls12377_g1 22 | tee -a ${T}_ab_g2x.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-tables --no-extras > ${T}_prof_bench.log 2>&1; echo "launch list rc=$?"
for cfg in "bn254_g1 24 7" "bls12381_g1 24 7" "bn254_g2 22 3"; do
  set -- $cfg
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s $3 -c 1 -f -o /tmp/prof_$1 \
      python bench.py --curve $1 --logn $2 --steps 1 --warmup 3 --no-e2e --no-cpu --no-tables --no-extras > ${T}_prof_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > ${T}_ncu_$1_raw.csv 2>/dev/null
done
ls -la gpurun_out | grep r2c8
