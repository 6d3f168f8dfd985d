#!/bin/bash
# parity + benches of all BASELINE single-GPU configs + window-width sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_bn254g1_2e24.json 2> gpurun_out/bench_bn254g1_2e24.err
echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_bn254g1_2e24.json'))
print(d['ms_per_step'], d['config']['c'], d['stages_ms'], d['e2e'], d.get('e2e_resident_bases'), d['cpu_baseline']['value'], d['clocks'])
PY
for C in 14 15 16 18 19; do
  GMSM_C=$C timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/sweep_c$C.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/sweep_c$C.json')); print('c=$C', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stages_ms'].items()})"
done
timeout 600 python bench.py --curve bls12381_g1 --steps 3 --warmup 3 > gpurun_out/bench_bls12381g1_2e24.json 2> gpurun_out/bench_bls.err
echo "bls rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_bls12381g1_2e24.json')); print(d['ms_per_step'], d['config']['c'], d['stages_ms'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --curve bn254_g2 --logn 22 --steps 3 --warmup 3 > gpurun_out/bench_bn254g2_2e22.json 2> gpurun_out/bench_g2.err
echo "g2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_bn254g2_2e22.json')); print(d['ms_per_step'], d['config']['c'], d['stages_ms'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'])"
tail -3 gpurun_out/*.err
