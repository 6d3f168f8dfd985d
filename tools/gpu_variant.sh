#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --logn 26 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_bn254g1_2e26.json 2>gpurun_out/v.err
python -c "
import json; d=json.load(open('gpurun_out/bench_bn254g1_2e26.json')); print('bn254_g1 2^26 c', d['config']['c'], 'ms', round(d['ms_per_step'],2), 'M/s', round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],2), 'res', round(d['e2e_resident_bases']['ms_per_step'],2), {k:round(v,2) for k,v in d['stages_ms'].items()})" || tail -3 gpurun_out/v.err
timeout 600 python bench.py --curve bls12381_g2 --logn 20 --steps 3 --warmup 3 > gpurun_out/bench_bls12381g2_2e20.json 2>gpurun_out/v2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_bls12381g2_2e20.json')); print('bls12381_g2 2^20 c', d['config']['c'], 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'cpu', round(d['cpu_baseline']['value']))" || tail -3 gpurun_out/v2.err
