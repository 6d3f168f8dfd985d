#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus N): full parity suite (incl. the NCCL tests), a quick N=1 line of the rebuilt default,
# then the driver's bench command at N ranks (strong scaling, 2^26 total) and the reference arm as the driver launches it
mkdir -p gpurun_out
T=gpurun_out/r2c4b
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5 ) 2>&1 | tee ${T}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-tables --no-extras --no-cpu > ${T}_bench_n1.json 2> ${T}_bench_n1.err; echo "N=1 rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c4b_bench_n1.json"))
    print("N=1 %.2f ms parity %s" % (d["ms_per_step"], d["parity"]["closed_form"]), {k:round(v,2) for k,v in d["stages_ms"].items()})
    for k in ("e2e","e2e_pageable","e2e_resident_bases"):
        v=d.get(k); print("  ",k, round(v["ms_per_step"],2) if v else None)
except Exception as e: print("N=1 no line", e)
PY
for kind in smallvalues redundancy; do :; done
N=$NG
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > ${T}_bench_n$N.json 2> ${T}_bench_n$N.err
echo "N=$N rc=$?"; tail -3 ${T}_bench_n$N.err
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
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > ${T}_reference_n$N.json 2> ${T}_reference_n$N.err; echo "reference N=$N rc=$?"; tail -c 500 ${T}_reference_n$N.json
ls -la gpurun_out | grep r2c4b
