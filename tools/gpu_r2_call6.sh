#!/bin/bash
# round 2, GPU call 6: two-mode counting sort (plain / rank) + multi-block probe: parity + timing; Fp2-by-dot2 variants for G2
mkdir -p gpurun_out
T=gpurun_out/r2c6
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5 ) 2>&1 | tee ${T}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > ${T}_bench_full.json 2> ${T}_bench_full.err; echo "bench full rc=$?"; tail -3 ${T}_bench_full.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c6_bench_full.json"))
    print({k:d[k] for k in ("value","ms_per_step")}, d["parity"]["closed_form"], d["engine"]["c"]); print({k:round(v,2) for k,v in d["stages_ms"].items()})
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
for lib in g2dot g2dott; do
  ( GMSM_LIB=$lib timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_ops.py -x -q -m gpu -p no:cacheprovider -k "g2 or G2 or ops or golden or all_window" 2>&1 | tail -2 ) | tee -a ${T}_pytest_g2dot.log
done
for lib in default g2dot g2dott; do ab $lib bn254_g2 22; ab $lib bls12381_g2 20; done | tee ${T}_ab_g2dot.txt
ls -la gpurun_out | grep r2c6
