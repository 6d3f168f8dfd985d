#!/bin/bash
# full parity + final benches of the single-GPU BASELINE configs
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_bn254g1_2e24.json 2> gpurun_out/b1.err; echo "rc=$?"
timeout 600 python bench.py --curve bls12381_g1 --steps 3 --warmup 3 > gpurun_out/bench_bls12381g1_2e24.json 2> gpurun_out/b2.err; echo "rc=$?"
timeout 600 python bench.py --curve bn254_g2 --logn 22 --steps 3 --warmup 3 > gpurun_out/bench_bn254g2_2e22.json 2> gpurun_out/b3.err; echo "rc=$?"
timeout 600 python bench.py --logn 20 --steps 5 --warmup 3 > gpurun_out/bench_bn254g1_2e20.json 2> gpurun_out/b4.err; echo "rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/b5.err; echo "rc=$?"
for f in bench_bn254g1_2e24 bench_bls12381g1_2e24 bench_bn254g2_2e22 bench_bn254g1_2e20; do python -c "
import json; d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['ms_per_step'],2), 'c', d['config']['c'], 'e2e', round(d['e2e']['ms_per_step'],2), 'resident', round(d['e2e_resident_bases']['ms_per_step'],2), 'cpu', round(d['cpu_baseline']['value']), 'frac', round(d['roofline']['frac'],4), d['clocks'])"; done
tail -c 600 gpurun_out/bench_reference.json
