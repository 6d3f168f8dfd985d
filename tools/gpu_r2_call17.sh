#!/bin/bash
# round 2, GPU call 17: ncu --set full captures of k_accumulate for the groups added / changed late in the round
mkdir -p gpurun_out
T=gpurun_out/r2c17
for cfg in "bls12377_g1 22" "secp256k1_g1 22" "bw6761_g1 20" "bls12381_g1 24"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 7 -c 1 -f -o /tmp/prof_$1 \
      python bench.py --curve $1 --logn $2 --steps 1 --warmup 3 --no-e2e --no-cpu --no-tables --no-extras > ${T}_prof_$1.log 2>&1
  echo "$1 ncu rc=$?"
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > ${T}_ncu_$1_raw.csv 2>/dev/null
done
ls -la gpurun_out | grep r2c17
