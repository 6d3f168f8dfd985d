#!/bin/bash
# round 2, GPU call 20: the adopted 3-blocks-per-SM build of the bls24 units in the default library: parity + timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_n4_curves.py tests/test_gpu_ops.py -x -q -m gpu -p no:cacheprovider -k "bls24315 or bls24317" 2>&1 | tail -2 | tee gpurun_out/r2c20_pytest.log
for cfg in "bls24315_g1 24" "bls24317_g1 24"; do set -- $cfg; python bench.py --curve $1 --logn $2 --steps 5 --warmup 3 --no-cpu --no-e2e --no-tables --no-extras | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['metric'], d['engine']['c'], round(d['ms_per_step'],3), d['parity']['closed_form'], round(d['roofline']['int_pipe']['multiplier_pipe']['frac'],3))"; done | tee gpurun_out/r2c20_bench.txt
