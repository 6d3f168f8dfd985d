#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
for CFG in "bn254_g1 24" "bls12381_g1 24" "bn254_g2 22"; do
  set -- $CFG
  timeout 300 python bench.py --curve $1 --logn $2 --steps 5 --warmup 3 > gpurun_out/bench_$1_$2.json 2>gpurun_out/v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$1_$2.json')); print('$1 2^$2 c', d['config']['c'], 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'res', round(d['e2e_resident_bases']['ms_per_step'],2), 'cpu', round(d['cpu_baseline']['value']), {k:round(v,2) for k,v in d['stages_ms'].items()})" || tail -3 gpurun_out/v.err
done
