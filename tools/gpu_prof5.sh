#!/bin/bash
# final round-1 ncu evidence for the default path: launch list of the bench command + full capture of k_accumulate
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_bench.log 2>&1
echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 1 -c 1 -f -o /tmp/prof_acc \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_acc.log 2>&1
echo "full rc=$?"
ncu -i /tmp/prof_acc.ncu-rep --page raw --csv > gpurun_out/prof_acc_final_raw.csv 2>/dev/null
ncu -i /tmp/prof_acc.ncu-rep --page source --csv > /tmp/src.csv 2>/dev/null; wc -c /tmp/src.csv; head -c 3000000 /tmp/src.csv > gpurun_out/prof_acc_final_source_head.csv
NG=$(nvidia-smi -L | wc -l)
if [ $NG -ge 2 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err
echo "N=2 rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/scale_n2.json').read().strip().splitlines()[-1]); print('N=2 value', round(d['value']/1e6,1), 'M/s ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'e2e value', round(d['e2e']['value']/1e6,1))"
fi
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_bn254g1_2e24.json 2> gpurun_out/b1.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_bn254g1_2e24.json')); print(round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'resident', round(d['e2e_resident_bases']['ms_per_step'],2), 'cpu', round(d['cpu_baseline']['value']), d['gpu_launches'])"
ls -la gpurun_out | tail -8
