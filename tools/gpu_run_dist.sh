#!/bin/bash
# multi-GPU: NCCL parity test + scaling bench N=1,2 (weak, 2^24 per GPU)
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -8
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  if [ $N -le $NG ]; then
    if [ $N -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
    fi
    echo "N=$N rc=$?"; tail -c 1200 gpurun_out/scale_n$N.json; tail -n 3 gpurun_out/scale_n$N.err
  fi
done
