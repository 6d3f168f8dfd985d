#!/bin/bash
mkdir -p gpurun_out
for LIB in ni; do
  for CFG in "bn254_g1 24" "bls12381_g1 23" "bn254_g2 22"; do
    set -- $CFG
    GMSM_LIB=$LIB timeout 300 python bench.py --curve $1 --logn $2 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/v_${LIB}_$1.json 2>gpurun_out/v.err
    python -c "
import json; d=json.load(open('gpurun_out/v_${LIB}_$1.json')); print('lib=[$LIB] $1 2^$2 ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stages_ms'].items()})" || tail -3 gpurun_out/v.err
  done
done
