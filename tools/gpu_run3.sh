#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -x 2>&1 | tail -6
for CH in 1 2 4 8; do
  GMSM_CHUNKS=$CH timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/e2e_ch$CH.json 2>gpurun_out/e2e_ch$CH.err
  python -c "
import json; d=json.load(open('gpurun_out/e2e_ch$CH.json')); print('chunks=$CH value_ms', round(d['ms_per_step'],2), 'e2e_ms', round(d['e2e']['ms_per_step'],2), 'resident_ms', round(d['e2e_resident_bases']['ms_per_step'],2))"
done
NG=$(nvidia-smi -L | wc -l)
if [ $NG -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err
  echo "N=2 rc=$?"; tail -c 900 gpurun_out/scale_n2.json; grep -E "Error|error" gpurun_out/scale_n2.err | tail -3
fi
