#!/bin/bash
# what the driver runs at round end, in order: gpu tests, smoke, reference arm, bench
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>gpurun_out/ref.err | tail -c 400; echo
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 2>gpurun_out/b.err | tee gpurun_out/bench_default.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('metric','value','ms_per_step','n_gpus','scaling','dtype','gpu_launches')}); print(d['e2e']); print(d['roofline']); print(d['cpu_baseline']); print(d['clocks'])"
tail -3 gpurun_out/b.err
