#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -5
for CH in 2 4 6 8; do
  GMSM_CHUNKS=$CH timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/e2e_sweep.json 2>gpurun_out/e2e.err
  python -c "
import json; d=json.load(open('gpurun_out/e2e_sweep.json')); print('chunks=$CH e2e_ms', round(d['e2e']['ms_per_step'],2), 'resident_ms', round(d['e2e_resident_bases']['ms_per_step'],2), 'dev', round(d['ms_per_step'],2))" || tail -2 gpurun_out/e2e.err
done
