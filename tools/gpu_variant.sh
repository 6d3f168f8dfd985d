#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu -p no:cacheprovider -k "window_sizes or config1 or infinity" 2>&1 | tail -3
for L in 32 16 8; do
  for CFG in "bn254_g1 24" "bn254_g2 22"; do
    set -- $CFG
    GMSM_SEG_L=$L timeout 300 python bench.py --curve $1 --logn $2 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/v.json 2>gpurun_out/v.err
    python -c "
import json; d=json.load(open('gpurun_out/v.json')); print('L=$L $1 2^$2 c', d['config']['c'], 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stages_ms'].items()})" || tail -3 gpurun_out/v.err
  done
done
