#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -25
for AFF in 1 0; do
  GMSM_AFFINE=$AFF timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/aff$AFF.json 2>gpurun_out/aff$AFF.err
  python -c "
import json; d=json.load(open('gpurun_out/aff$AFF.json')); print('affine=$AFF ms', round(d['ms_per_step'],2), d['config']['c'], {k:round(v,2) for k,v in d['stages_ms'].items()}, d['gpu_launches'])"
  tail -2 gpurun_out/aff$AFF.err
done
