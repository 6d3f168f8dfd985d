#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 1 -c 1 -f -o gpurun_out/prof_acc_g2 \
    python bench.py --curve bn254_g2 --logn 22 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_acc_g2.log 2>&1
echo "g2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 1 -c 1 -f -o gpurun_out/prof_acc_bls \
    python bench.py --curve bls12381_g1 --logn 22 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_acc_bls.log 2>&1
echo "bls rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_scatter_window|k_digits_hist|k_bucket_segments|k_finalize|k_carry_level" -s 20 -c 24 -f -o gpurun_out/prof_misc \
    python bench.py --logn 24 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_misc.log 2>&1
echo "misc rc=$?"
