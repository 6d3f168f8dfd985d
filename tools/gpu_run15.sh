#!/bin/bash
# validate K2_first = 4 + all-or-nothing precompute on the full gpu suite; sweep the one-shot batch schedule (e2e)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r15_pytest_gpu.log
OUT=gpurun_out/r15_schedule.txt
: > $OUT
run() {
  label=$1; shift
  env "$@" python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --no-tables 2>>gpurun_out/r15.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$label | device %.2f ms | e2e one-shot %.2f ms | resident bases %.2f ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_resident_bases']['ms_per_step']))" | tee -a $OUT
}
run default X=1
run s_1_2_3_5_8 GMSM_SCHEDULE=1,2,3,5,8
run s_2_3_4_7 GMSM_SCHEDULE=2,3,4,7
run s_1_2_4_6 GMSM_SCHEDULE=1,2,4,6
run s_1_3_5_7 GMSM_SCHEDULE=1,3,5,7
run s_1_2_3_4_6_9 GMSM_SCHEDULE=1,2,3,4,6,9
run s_2_3_5 GMSM_SCHEDULE=2,3,5
