#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus N): NCCL parity tests + the driver's bench command at N ranks (strong scaling, 2^26 total)
mkdir -p gpurun_out
T=gpurun_out/r2c4b
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
( time timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 ) 2>&1 | tee ${T}_pytest_dist.log
for N in 2 $NG; do
  [ $N -gt $NG ] && continue
  [ $N -eq 2 ] && [ $NG -eq 2 ] && [ -f ${T}_bench_n2.json ] && continue
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > ${T}_bench_n$N.json 2> ${T}_bench_n$N.err
  echo "N=$N rc=$?"; tail -2 ${T}_bench_n$N.err
  python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/r2c4b_bench_n%s.json"%N).read().strip().splitlines()[-1])
    print("N=%s"%N, d["scaling"], d["config"]["workload"], "value %.1f M/s  %.2f ms  parity %s" % (d["value"]/1e6, d["ms_per_step"], d["parity"]))
    print("  stages", {k:round(v,2) for k,v in d["stages_ms"].items()})
    for k in ("e2e","e2e_pageable","weak"):
        v=d.get(k)
        if v: print("  ",k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ("value","ms_per_step","parity","n_total","vs_pinned")})
except Exception as e:
    print("N=%s: no line: %s"%(N,e))
PY
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > ${T}_reference_n2.json 2> ${T}_reference_n2.err; echo "reference N=2 rc=$?"; tail -c 400 ${T}_reference_n2.json
ls -la gpurun_out | grep r2c4b
