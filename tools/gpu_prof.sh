#!/bin/bash
# ncu passes (B200_PROFILING.md): launch list of a short bench, then one full capture of the top kernel.
mkdir -p gpurun_out
LOGN=${LOGN:-24}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --logn $LOGN --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_bench.log 2>&1
echo "launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 1 -c 1 -f -o gpurun_out/prof_acc \
    python bench.py --logn $LOGN --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_acc.log 2>&1
echo "full rc=$?"
ls -la gpurun_out | tail -8
