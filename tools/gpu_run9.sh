#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "N=1 rc=$?"
for N in 2 4 8; do
  if [ $N -le $NG ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
    echo "N=$N rc=$?"
  fi
done
for N in 1 2 4 8; do [ -s gpurun_out/scale_n$N.json ] && python -c "
import json
d=json.loads(open('gpurun_out/scale_n$N.json').read().strip().splitlines()[-1]); print('N=$N value', round(d['value']/1e6,1), 'M/s ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'e2e value', round(d['e2e']['value']/1e6,1))"; done
grep -iE "error" gpurun_out/scale_n*.err | tail -5
