#!/bin/bash
# table-mode tuning: number of scatter passes, overlap split, segment length
mkdir -p gpurun_out
OUT=gpurun_out/r12_diag.txt
: > $OUT
run() {
  label=$1; shift
  env "$@" python bench.py --gpus 1 --steps 5 --warmup 3 --no-e2e --no-cpu $BARGS 2>>gpurun_out/r12.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d.get('resident_tables') or {}
f=lambda s: ' '.join('%s=%.2f'%(k[:4],v) for k,v in s.items())
print('$label | plain %.2f ms | table c=%s %.2f ms [%s]' % (d['ms_per_step'], t.get('c'), t.get('ms_per_step',0), f(t.get('stages_ms',{}))))" | tee -a $OUT
}
BARGS=""
run p1 GMSM_TABLE_PASSES=1
run p2 GMSM_TABLE_PASSES=2
run p3 GMSM_TABLE_PASSES=3
run p4 GMSM_TABLE_PASSES=4
run p5 GMSM_TABLE_PASSES=5
run p4s1 GMSM_TABLE_PASSES=4 GMSM_SPLIT_W=1
run p6s1 GMSM_TABLE_PASSES=6 GMSM_SPLIT_W=1
run p6s3 GMSM_TABLE_PASSES=6 GMSM_SPLIT_W=3
run p6L128 GMSM_TABLE_PASSES=6 GMSM_TABLE_SEG_L=128
run p6L96 GMSM_TABLE_PASSES=6 GMSM_TABLE_SEG_L=96
