#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_cpp_mirror.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err
echo "N=2 rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/scale_n2.json').read().strip().splitlines()[-1]); print('N=2 value', round(d['value']/1e6,1), 'M/s ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'e2e value', round(d['e2e']['value']/1e6,1))"; grep -iE "error" gpurun_out/scale_n2.err | tail -3
