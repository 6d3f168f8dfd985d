#!/bin/bash
mkdir -p gpurun_out
for CFG in "bn254_g2 22 g2" "bls12381_g1 22 bls"; do
  set -- $CFG
  timeout 900 ncu --set full --clock-control none -k regex:k_accumulate -s 1 -c 1 -f -o /tmp/prof_$3 \
      python bench.py --curve $1 --logn $2 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_$3.log 2>&1
  echo "$3 rc=$?"
  ncu -i /tmp/prof_$3.ncu-rep --page raw --csv > gpurun_out/prof_acc_$3_raw.csv 2>/dev/null
done
timeout 900 ncu --set full --clock-control none -k regex:"k_scatter_window|k_digits_hist|k_bucket_segments|k_finalize|k_carry_level" -s 20 -c 12 -f -o /tmp/prof_misc \
    python bench.py --logn 24 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_misc.log 2>&1
echo "misc rc=$?"
ncu -i /tmp/prof_misc.ncu-rep --page raw --csv > gpurun_out/prof_misc_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
