#!/bin/bash
# first GPU pass: parity tests, then small + default bench.  Everything bounded by `timeout`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > gpurun_out/cpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --logn 20 --steps 3 --warmup 3 > gpurun_out/bench_2e20.json 2> gpurun_out/bench_2e20.err
echo "bench20 rc=$?"; tail -c 3000 gpurun_out/bench_2e20.json; tail -5 gpurun_out/bench_2e20.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_2e24.json 2> gpurun_out/bench_2e24.err
echo "bench24 rc=$?"; tail -c 3000 gpurun_out/bench_2e24.json; tail -5 gpurun_out/bench_2e24.err
