#!/bin/bash
# round-end order on the final state: gpu tests, smoke, reference arm, bench (as the driver runs them: --steps 20 --warmup 5)
mkdir -p gpurun_out
T=gpurun_out/r2final
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
SEL=tests; [ "$NG" -gt 1 ] && SEL=tests/test_gpu_dist.py     # the N-GPU box is charged N x: only the tests that need it
( time timeout 1500 python -m pytest $SEL -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5 ) 2>&1 | tee ${T}_pytest_n$NG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "$NG" -eq 1 ]; then
  ( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > ${T}_reference_n1.json 2> ${T}_reference_n1.err; tail -c 700 ${T}_reference_n1.json; tail -4 ${T}_reference_n1.err
  ( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > ${T}_bench_n1.json 2> ${T}_bench_n1.err; echo "bench rc=$?"; tail -4 ${T}_bench_n1.err
  python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2final_bench_n1.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("metric","value","ms_per_step","n_gpus","scaling","dtype","gpu_launches")}, d["parity"]["closed_form"], d["engine"]["c"], "frac", round(d["roofline"]["frac"],4), "pipe", round(d["roofline"]["int_pipe"]["multiplier_pipe"]["frac"],3)); print(d["clocks"]); print({k:round(v,2) for k,v in d["stages_ms"].items()})
    for k in ("e2e","e2e_pageable","e2e_resident_bases","e2e_resident_tables","cpu_baseline","concurrent3","two_in_flight","resident_tables"):
        v=d.get(k); print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a not in ("sample","path","what","stages_ms","note","workload")} if v else None)
    for k,v in (d.get("configs") or {}).items():
        print(k, round(v["ms_per_step"],3), v["parity"]["closed_form"], "c=%d"%v["c"], {a:round(b,2) for a,b in v["stages_ms"].items()}, v.get("vs_random_scalars"), "frac", round(v["roofline"]["frac"],4))
except Exception as e:
    print("no line:", e)
PY
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 10 --warmup 3 > ${T}_bench_n$NG.json 2> ${T}_bench_n$NG.err
  echo "N=$NG rc=$?"; tail -3 ${T}_bench_n$NG.err
  python - $NG <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/r2final_bench_n%s.json"%N).read().strip().splitlines()[-1])
    print("N=%s"%N, d["scaling"], d["config"]["workload"], "value %.1f M/s  %.2f ms  parity %s c=%d" % (d["value"]/1e6, d["ms_per_step"], d["parity"], d["engine"]["c"]))
    print("  stages", {k:round(v,2) for k,v in d["stages_ms"].items()})
    for k in ("e2e","e2e_pageable","weak"):
        v=d.get(k)
        if v: print("  ",k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ("value","ms_per_step","parity","n_total","vs_pinned")})
except Exception as e:
    print("N=%s: no line: %s"%(N,e))
PY
fi
ls -la gpurun_out | grep r2final
