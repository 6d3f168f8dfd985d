#!/bin/bash
# round 2, GPU call 3: parity of the rebuilt default (batched K1 atomics, serial tail + tree sums, per-group multiplier variants,
# binary-GCD inversion, device point decoder), the full bench line, segment-length sweep, window-width sweep
mkdir -p gpurun_out
T=gpurun_out/r2c3
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6 ) 2>&1 | tee ${T}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > ${T}_bench_full.json 2> ${T}_bench_full.err; echo "bench full rc=$?"; tail -3 ${T}_bench_full.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c3_bench_full.json"))
    print({k:d[k] for k in ("value","ms_per_step","parity")}); print({k:round(v,2) for k,v in d["stages_ms"].items()})
    for k in ("e2e","e2e_pageable","e2e_resident_bases","cpu_baseline","concurrent3"):
        v=d.get(k); print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a not in ("sample","path","what")} if v else None)
    for k,v in (d.get("configs") or {}).items():
        print(k, round(v["ms_per_step"],3), v["parity"]["closed_form"], {a:round(b,2) for a,b in v["stages_ms"].items()}, v.get("vs_random_scalars"))
except Exception as e:
    print("no full line:", e)
PY
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
    print("%-12s 2^%s %-14s %.3f ms parity=%s stages %s" % (curve,logn," ".join(extra),d["ms_per_step"],d["parity"]["closed_form"],{k:round(v,2) for k,v in d["stages_ms"].items()}))
except Exception as e:
    print("%s 2^%s %s FAILED: %s" % (curve,logn,extra,(r.stderr or r.stdout)[-300:]))
PY
}
for L in 8 16 32 64; do ab bn254_g1 24 GMSM_SEG_L=$L; ab bn254_g1 20 GMSM_SEG_L=$L; ab bn254_g2 22 GMSM_SEG_L=$L; ab bls12381_g1 22 GMSM_SEG_L=$L; done | tee ${T}_seg_sweep.txt
for cfg in "bn254_g1 24" "bn254_g1 23" "bn254_g1 20" "bn254_g1 16" "bls12381_g1 24" "bn254_g2 22"; do
  set -- $cfg
  for c in 13 14 15 16 17 18 19; do ab $1 $2 c=$c; done
done | tee ${T}_c_sweep.txt
ls -la gpurun_out | grep r2c3
