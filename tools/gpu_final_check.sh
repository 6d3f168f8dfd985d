#!/bin/bash
# round-end order on the final state: gpu tests, smoke, reference arm, bench; then the first-carry-level knob for information
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r16_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>gpurun_out/r16_ref.err | tee gpurun_out/r16_bench_reference.json | tail -c 300; echo
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 2>gpurun_out/r16_b.err | tee gpurun_out/r16_bench_default.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('metric','value','ms_per_step','n_gpus','scaling','dtype','gpu_launches')}); print(d['e2e']); print(d['roofline']); print(d['cpu_baseline']); print(d['clocks']); print(d.get('resident_tables')); print(d.get('e2e_resident_bases')); print(d.get('e2e_resident_tables'))"
tail -3 gpurun_out/r16_b.err
for k in 16; do
GMSM_K2_FIRST=$k timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>>gpurun_out/r16_b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d.get('resident_tables') or {}
f=lambda s: ' '.join('%s=%.2f'%(k[:4],v) for k,v in s.items())
print('K2_FIRST=$k | plain %.2f ms [%s] | table %.2f ms [%s]' % (d['ms_per_step'], f(d['stages_ms']), t.get('ms_per_step',0), f(t.get('stages_ms',{}))))" | tee -a gpurun_out/r16_k2first.txt
done
