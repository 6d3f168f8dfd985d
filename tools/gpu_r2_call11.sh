#!/bin/bash
# round 2, GPU call 11: the N4 remainder on hardware -- secp256k1 G1 (full-width moduli) and bw6-761 G1/G2 (24-limb Fp, 6-word
# scalars): parity tests, bench lines, width sweeps for the window model
mkdir -p gpurun_out
T=gpurun_out/r2c11
( time timeout 1200 python -m pytest tests/test_gpu_n4_curves.py tests/test_gpu_ops.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6 ) 2>&1 | tee ${T}_pytest.log
ab() {
  python - "$@" <<'PY'
import json,sys,subprocess,os
curve,logn=sys.argv[1:3]; extra=sys.argv[3:]
env=dict(os.environ); args=[]
for kv in extra:
    k,v=kv.split("=")
    if k=="c": args+=["--c",v]
    else: env[k]=v
r=subprocess.run(["python","bench.py","--curve",curve,"--logn",logn,"--steps","5","--warmup","3","--no-cpu","--no-e2e","--no-tables","--no-extras"]+args,capture_output=True,text=True,env=env)
try:
    d=json.loads(r.stdout.strip().splitlines()[-1])
    print("%-12s 2^%s %-6s c=%d %.3f ms parity=%s pipe=%.3f stages %s" % (curve,logn," ".join(extra),d["engine"]["c"],d["ms_per_step"],d["parity"]["closed_form"],d["roofline"]["int_pipe"]["multiplier_pipe"]["frac"],{k:round(v,2) for k,v in d["stages_ms"].items()}))
except Exception as e:
    print("%s 2^%s %s FAILED: %s" % (curve,logn,extra,(r.stderr or r.stdout)[-400:]))
PY
}
{
ab secp256k1_g1 24; ab bw6761_g1 22; ab bw6761_g2 20
for c in 15 16 17 18; do ab secp256k1_g1 24 c=$c; done
for c in 12 13 14 15; do ab secp256k1_g1 20 c=$c; done
for c in 14 15 16 17; do ab bw6761_g1 22 c=$c; done
for c in 11 12 13 14; do ab bw6761_g1 18 c=$c; done
} | tee ${T}_sweeps.txt
for cfg in "secp256k1_g1 24" "bw6761_g1 22"; do
  set -- $cfg
  timeout 600 python bench.py --curve $1 --logn $2 --steps 5 --warmup 3 --no-extras > ${T}_bench_$1.json 2> ${T}_bench_$1.err; echo "bench $1 rc=$?"; tail -2 ${T}_bench_$1.err
  python - $1 <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2c11_bench_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("metric","value","ms_per_step","dtype","gpu_launches")}, d["parity"]["closed_form"], "c", d["engine"]["c"])
    for k in ("e2e","e2e_pageable","e2e_resident_bases","e2e_resident_tables","cpu_baseline"):
        v=d.get(k); print(" ",k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a not in ("sample","path","what","stages_ms","note","workload","host_memory")} if v else None)
except Exception as e:
    print("no line:", e)
PY
done
ls -la gpurun_out | grep r2c11
