#!/bin/bash
# Next step for the experimental multiplier routines: build the variant library HERE (nvcc
# cross-compiles), then A/B it against the default library on one B200 box -- parity first, then the bench.
#   bash tools/next_fast_variant.sh build      # on the build container
#   gpurun --timeout 900 -- 'bash tools/next_fast_variant.sh run'
set -e
case "$1" in
  build)
    GMSM_BUILD_TAG=fast GMSM_NVCC_EXTRA="-DGMSM_SQR_DEDICATED=1 -DGMSM_DOT2=1" python gnark-crypto_b200/build.py
    ;;
  run)
    mkdir -p gpurun_out
    GMSM_LIB=fast timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/fast_pytest.log
    for lib in "" fast; do
      for curve in bn254_g1 bls12381_g1; do
        GMSM_LIB=$lib timeout 300 python bench.py --curve $curve --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/fast.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d.get('resident_tables') or {}
print('lib=%-5s %-12s plain %.2f ms (accumulate %.2f) | tables %.2f ms' % ('$lib' or 'default', '$curve', d['ms_per_step'], d['stages_ms']['accumulate'], t.get('ms_per_step', 0)))" | tee -a gpurun_out/fast_ab.txt
      done
    done
    ;;
  *) echo "usage: $0 build|run"; exit 1 ;;
esac
