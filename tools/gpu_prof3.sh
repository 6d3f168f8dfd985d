#!/bin/bash
# ncu: launch list + full capture of the affine level-0 kernels; export CSV on the box, drop the .ncu-rep
mkdir -p gpurun_out
export GMSM_AFFINE=${GMSM_AFFINE:-1}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/launches_aff.csv \
    python bench.py --logn 24 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_bench.log 2>&1
echo "launches rc=$?"
timeout 1200 ncu --set full --clock-control none -k regex:"k_aff_forward|k_aff_backward" -s 2 -c 4 -f -o /tmp/prof_aff \
    python bench.py --logn 24 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_aff.log 2>&1
echo "full rc=$?"
ncu -i /tmp/prof_aff.ncu-rep --page raw --csv > gpurun_out/prof_aff_raw.csv 2>/dev/null
ls -la gpurun_out | tail -6
