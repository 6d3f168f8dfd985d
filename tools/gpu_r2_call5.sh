#!/bin/bash
# round 2, GPU call 5: parity of the current default, full bench line, host-copy thread sweep, out-of-sample width sweep for the
# window model, ncu launch list + full captures of k_accumulate for the three graded configurations
mkdir -p gpurun_out
T=gpurun_out/r2c5
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5 ) 2>&1 | tee ${T}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > ${T}_bench_full.json 2> ${T}_bench_full.err; echo "bench full rc=$?"; tail -3 ${T}_bench_full.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c5_bench_full.json"))
    print({k:d[k] for k in ("value","ms_per_step")}, d["parity"]["closed_form"], d["engine"]); print({k:round(v,2) for k,v in d["stages_ms"].items()})
    for k in ("e2e","e2e_pageable","e2e_resident_bases","e2e_resident_tables","cpu_baseline","concurrent3","resident_tables"):
        v=d.get(k); print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a not in ("sample","path","what","stages_ms","note","workload")} if v else None)
    for k,v in (d.get("configs") or {}).items():
        print(k, round(v["ms_per_step"],3), v["parity"]["closed_form"], "c=%d"%v["c"], {a:round(b,2) for a,b in v["stages_ms"].items()}, v.get("vs_random_scalars"))
except Exception as e:
    print("no full line:", e)
PY
for th in 4 8 16 32; do
GMSM_COPY_THREADS=$th timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-tables --no-extras 2>>${T}.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('GMSM_COPY_THREADS=$th  e2e pinned', round(d['e2e']['ms_per_step'],2), 'e2e_pageable', round(d['e2e_pageable']['ms_per_step'],2))" | tee -a ${T}_copy_threads.txt
done
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
    print("%-12s 2^%s %-6s -> c=%d %.3f ms parity=%s stages %s" % (curve,logn," ".join(extra),d["engine"]["c"],d["ms_per_step"],d["parity"]["closed_form"],{k:round(v,2) for k,v in d["stages_ms"].items()}))
except Exception as e:
    print("%s 2^%s %s FAILED: %s" % (curve,logn,extra,(r.stderr or r.stdout)[-300:]))
PY
}
for cfg in "bn254_g1 22" "bn254_g1 18" "bls12381_g1 22" "bls12381_g1 20" "bn254_g2 20" "bls12381_g2 20" "bls12377_g1 22"; do
  set -- $cfg
  for c in 0 14 15 16 17 18 19; do ab $1 $2 c=$c; done
done | tee ${T}_c_sweep_oos.txt
# ncu: launch list of one bench command, then full captures of the dominant kernel (part-2 launch of the 4th step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-tables --no-extras > ${T}_prof_bench.log 2>&1; echo "launch list rc=$?"
for cfg in "bn254_g1 24" "bls12381_g1 24" "bn254_g2 22"; do
  set -- $cfg
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 7 -c 1 -f -o /tmp/prof_$1 \
      python bench.py --curve $1 --logn $2 --steps 1 --warmup 3 --no-e2e --no-cpu --no-tables --no-extras > ${T}_prof_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > ${T}_ncu_$1_raw.csv 2>/dev/null
done
ls -la gpurun_out | grep r2c5
