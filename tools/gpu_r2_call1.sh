#!/bin/bash
# round 2, GPU call 1: pipe-rate / modmul microbenchmarks, parity + A/B of the multiplier variants, fresh ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/r2c1_smi.txt
timeout 300 tools/ubench/ubench 2>&1 | tee gpurun_out/r2c1_ubench.txt | tail -70
GMSM_LIB=fast timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r2c1_fast_pytest.log
for lib in "" fast; do
  for cfg in "bn254_g1 24" "bn254_g1 20" "bls12381_g1 24" "bn254_g2 22"; do
    set -- $cfg
    GMSM_LIB=$lib timeout 300 python bench.py --curve $1 --logn $2 --steps 5 --warmup 3 --no-cpu --no-e2e --no-tables --no-extras 2>>gpurun_out/r2c1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('lib=%-7s %-12s 2^%s  %.2f ms  stages %s  clocks %s' % ('$lib' or 'default', '$1', '$2', d['ms_per_step'], {k: round(v,2) for k,v in d['stages_ms'].items()}, d['clocks']))" | tee -a gpurun_out/r2c1_ab.txt
  done
done
# fresh full captures of the CURRENT builds: bls12-381 G1 and bn254 G2, default and fast
for lib in "" fast; do
  for cfg in "bls12381_g1 22" "bn254_g2 22" "bn254_g1 24"; do
    set -- $cfg
    GMSM_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 2 -c 1 -f -o /tmp/prof_$1_${lib:-default} \
      python bench.py --curve $1 --logn $2 --steps 1 --warmup 1 --no-e2e --no-cpu --no-tables --no-extras > gpurun_out/r2c1_prof_$1_${lib:-default}.log 2>&1
    ncu -i /tmp/prof_$1_${lib:-default}.ncu-rep --page raw --csv > gpurun_out/r2c1_ncu_$1_${lib:-default}_raw.csv 2>/dev/null
  done
done
ls -la gpurun_out | tail -20
