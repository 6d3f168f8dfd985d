#!/bin/bash
mkdir -p gpurun_out
for SW in 2 3 4 6 64; do
  for CFG in "bn254_g1 24" "bn254_g1 20"; do
    set -- $CFG
    GMSM_SPLIT_W=$SW timeout 200 python bench.py --curve $1 --logn $2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/v.json 2>gpurun_out/v.err
    python -c "
import json; d=json.load(open('gpurun_out/v.json')); print('SPLIT_W=$SW $1 2^$2 ms', round(d['ms_per_step'],3), {k:round(v,2) for k,v in d['stages_ms'].items() if k in ('scatter','accumulate')})" || tail -2 gpurun_out/v.err
  done
done
