#!/bin/bash
# window-width calibration: device-resident ms per step for forced c around the model's choice
mkdir -p gpurun_out
: > gpurun_out/csweep.txt
run() { # curve logn c
  GMSM_C=$3 timeout 200 python bench.py --curve $1 --logn $2 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/cs.json 2>gpurun_out/cs.err
  python -c "
import json; d=json.load(open('gpurun_out/cs.json')); print('$1 2^$2 c=$3 ->', d['config']['c'], 'ms', round(d['ms_per_step'],3), {k:round(v,2) for k,v in d['stages_ms'].items() if k in ('accumulate','bucket_reduce','scatter','carries')})" | tee -a gpurun_out/csweep.txt
}
for L in 16 18 20 22; do for C in 0 11 12 13 14 15 16 17; do run bn254_g1 $L $C; done; done
for C in 0 15 16 17 18; do run bls12381_g1 22 $C; done
for C in 0 14 15 16 17; do run bn254_g2 20 $C; done
for C in 0 15 16 17 18; do run bn254_g2 22 $C; done
