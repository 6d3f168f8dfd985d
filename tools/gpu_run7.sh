#!/bin/bash
mkdir -p gpurun_out
for C in 0 16 17; do for CH in 2 3 4 6; do
  GMSM_C=$C GMSM_CHUNKS=$CH timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/e2e_sweep.json 2>gpurun_out/e2e.err
  python -c "
import json; d=json.load(open('gpurun_out/e2e_sweep.json')); print('C=$C chunks=$CH e2e_ms', round(d['e2e']['ms_per_step'],2), 'resident_ms', round(d['e2e_resident_bases']['ms_per_step'],2), 'dev', round(d['ms_per_step'],2))" || tail -2 gpurun_out/e2e.err
done; done
