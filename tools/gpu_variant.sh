#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
timeout 300 python bench.py --curve bls12377_g1 --logn 22 --steps 3 --warmup 3 > gpurun_out/bench_bls12377g1_2e22.json 2>gpurun_out/v.err
python -c "
import json; d=json.load(open('gpurun_out/bench_bls12377g1_2e22.json')); print('bls12377_g1 2^22 c', d['config']['c'], 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'cpu', round(d['cpu_baseline']['value']))" || tail -3 gpurun_out/v.err
