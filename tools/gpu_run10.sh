#!/bin/bash
# window-table mode: parity tests, default bench line (incl. resident_tables), table-width sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tables.py tests/test_cpp_mirror.py "tests/test_gpu_msm.py::test_kzg_commit_over_generated_srs_and_dump_roundtrip" -x -q -m gpu -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/r10_pytest_tables.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 2>gpurun_out/r10_b.err | tee gpurun_out/r10_bench_default.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['stages_ms']); print(d.get('resident_tables')); print(d['e2e']); print(d.get('e2e_resident_bases')); print(d.get('e2e_resident_tables')); print(d['clocks'])"
tail -3 gpurun_out/r10_b.err
for c in 20 21 23 24; do
  timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e --no-cpu --table-c $c 2>>gpurun_out/r10_b.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); t=d['resident_tables']; print('table c=%d W=%d ms=%.2f build_s=%.2f' % (t['c'], t['windows'], t['ms_per_step'], t['table_build_s']), t['stages_ms'])" | tee -a gpurun_out/r10_table_c_sweep.txt
done
