#!/bin/bash
# table-mode diagnostics: prefetch A/B, overlap on/off, scatter passes, segment length, size probe
mkdir -p gpurun_out
OUT=gpurun_out/r11_diag.txt
: > $OUT
run() {  # label, env..., -- bench args
  label=$1; shift
  env "$@" python bench.py --gpus 1 --steps 5 --warmup 3 --no-e2e --no-cpu $BARGS 2>>gpurun_out/r11.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d.get('resident_tables') or {}
f=lambda s: ' '.join('%s=%.2f'%(k[:4],v) for k,v in s.items())
print('$label | plain %.2f ms [%s] | table c=%s %.2f ms [%s]' % (d['ms_per_step'], f(d['stages_ms']), t.get('c'), t.get('ms_per_step',0), f(t.get('stages_ms',{}))))" | tee -a $OUT
}
BARGS=""
run default X=1
run nopf GMSM_LIB=nopf
run default_again X=1
run nosplit GMSM_SPLIT_W=64
run passes6 GMSM_TABLE_PASSES=6
run passes8 GMSM_TABLE_PASSES=8
run passes24 GMSM_TABLE_PASSES=24
run segL16 GMSM_SEG_L=16
run segL64 GMSM_SEG_L=64
BARGS="--logn 22"
run logn22 X=1
BARGS="--logn 20"
run logn20 X=1
