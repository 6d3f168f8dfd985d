#!/bin/bash
# round 2, GPU call 2: parity of the new default (quad tail, staging ring, session pool), the full bench line, A/B of the
# multiplier variants / the quad tail / the staging ring, the reference arm at full size
mkdir -p gpurun_out
T=gpurun_out/r2c2
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6 ) 2>&1 | tee ${T}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > ${T}_bench_full.json 2> ${T}_bench_full.err; echo "bench full rc=$?"; tail -3 ${T}_bench_full.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c2_bench_full.json"))
    print({k:d[k] for k in ("value","ms_per_step","parity","stages_ms")})
    for k in ("e2e","e2e_pageable","e2e_resident_bases","cpu_baseline","concurrent3"):
        print(k, d.get(k))
    for k,v in (d.get("configs") or {}).items():
        print(k, round(v["ms_per_step"],3), v["parity"], {a:round(b,2) for a,b in v["stages_ms"].items()}, v.get("vs_random_scalars"))
except Exception as e:
    print("no full line:", e)
PY
ab() {  # lib env... : one short device-resident line
  python - "$@" <<'PY'
import json,sys,subprocess,os
lib,curve,logn=sys.argv[1:4]; extra=sys.argv[4:]
env=dict(os.environ); env["GMSM_LIB"]=lib if lib!="default" else ""
for kv in extra:
    k,v=kv.split("="); env[k]=v
r=subprocess.run(["python","bench.py","--curve",curve,"--logn",logn,"--steps","5","--warmup","3","--no-cpu","--no-e2e","--no-tables","--no-extras"],capture_output=True,text=True,env=env)
try:
    d=json.loads(r.stdout.strip().splitlines()[-1])
    print("lib=%-8s %-12s 2^%s %s  %.2f ms parity=%s stages %s" % (lib,curve,logn," ".join(extra),d["ms_per_step"],d["parity"]["closed_form"],{k:round(v,2) for k,v in d["stages_ms"].items()}))
except Exception as e:
    print("lib=%s %s 2^%s FAILED: %s" % (lib,curve,logn,(r.stderr or r.stdout)[-300:]))
PY
}
for lib in default sqr fast3; do ab $lib bn254_g1 24; done | tee ${T}_ab_mul.txt
for lib in default sqr fast3; do ab $lib bls12381_g1 24; ab $lib bn254_g2 22; done | tee -a ${T}_ab_mul.txt
for q in 0 1; do ab default bn254_g1 24 GMSM_QUAD=$q; ab default bn254_g1 20 GMSM_QUAD=$q; ab default bn254_g2 22 GMSM_QUAD=$q; ab default bls12381_g1 22 GMSM_QUAD=$q; done | tee ${T}_ab_quad.txt
ab default bn254_g1 20 | tee -a ${T}_ab_quad.txt
# staging ring off: pageable e2e through cudaMemcpyAsync's own bounce buffer
GMSM_STAGING=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-tables --no-extras 2>>${T}.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('GMSM_STAGING=0  e2e', round(d['e2e']['ms_per_step'],2), 'e2e_pageable', round(d['e2e_pageable']['ms_per_step'],2))" | tee ${T}_staging_off.txt
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > ${T}_reference.json 2> ${T}_reference.err; tail -c 600 ${T}_reference.json; tail -4 ${T}_reference.err
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_host_path.py -x -q -k "fault" -p no:cacheprovider 2>&1 | tail -8 | tee ${T}_sanitizer_fault.log
ls -la gpurun_out | tail -12
