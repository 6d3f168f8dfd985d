#!/bin/bash
# round 2, GPU call 18: last validation of the committed state (full gpu suite incl. the C++ mirror with the N4 curves, smoke) and
# compute-sanitizer memcheck over the N4 curves' kernel paths
mkdir -p gpurun_out
T=gpurun_out/r2c18
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5 ) 2>&1 | tee ${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_n4.py > ${T}_sanitize_memcheck_n4.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|SANITIZE_RUN|Error|Invalid" ${T}_sanitize_memcheck_n4.log | head -8
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras ) > ${T}_bench_n1.json 2> ${T}_bench_n1.err; echo "bench rc=$?"; tail -c 600 ${T}_bench_n1.json
