#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
for DEVS in 0 0,1 0,1,2,3; do
  GMSM_DEVICES=$DEVS timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/v_md.json 2>gpurun_out/v.err
  python -c "
import json; d=json.load(open('gpurun_out/v_md.json')); print('GMSM_DEVICES=$DEVS dev ms', round(d['ms_per_step'],2), 'one-shot e2e ms', round(d['e2e']['ms_per_step'],2), 'M/s', round(d['e2e']['value']/1e6,1))" || tail -3 gpurun_out/v.err
done
