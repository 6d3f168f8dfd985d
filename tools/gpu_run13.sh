#!/bin/bash
# table mode with its final defaults: parity tests, other groups' lines, ncu launch list + full capture of k_accumulate
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tables.py tests/test_cpp_mirror.py "tests/test_gpu_msm.py::test_kzg_commit_over_generated_srs_and_dump_roundtrip" -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/r13_pytest.log
line() {
  python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d.get('resident_tables') or {}
f=lambda s: ' '.join('%s=%.2f'%(k[:4],v) for k,v in s.items())
print('$1 | plain c=%s %.2f ms [%s] | table c=%s W=%s %.2f ms build %.2fs [%s]' % (d['config']['c'], d['ms_per_step'], f(d['stages_ms']), t.get('c'), t.get('windows'), t.get('ms_per_step',0), t.get('table_build_s',0), f(t.get('stages_ms',{}))))"
}
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>>gpurun_out/r13.err | tee gpurun_out/r13_bench_bn254g1_2e24.json | line bn254_g1_2e24 | tee gpurun_out/r13_lines.txt
timeout 300 python bench.py --curve bls12381_g1 --steps 3 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/r13.err | tee gpurun_out/r13_bench_bls12381g1_2e24.json | line bls12381_g1_2e24 | tee -a gpurun_out/r13_lines.txt
timeout 300 python bench.py --curve bn254_g2 --logn 22 --steps 3 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/r13.err | tee gpurun_out/r13_bench_bn254g2_2e22.json | line bn254_g2_2e22 | tee -a gpurun_out/r13_lines.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r13_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r13_prof_bench.log 2>&1
echo "launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 14 -c 2 -f -o /tmp/prof_acc_tab \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r13_prof_acc.log 2>&1
echo "full rc=$?"
ncu -i /tmp/prof_acc_tab.ncu-rep --page raw --csv > gpurun_out/r13_prof_acc_tables_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
