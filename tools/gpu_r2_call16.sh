#!/bin/bash
# round 2, GPU call 16: final state after the bls12-377 multiplier fix (INV / modulus multiplicands through the constant bank) and
# the c = 20 window choices: full gpu suite, smoke, the driver's bench line, the affected configurations
mkdir -p gpurun_out
T=gpurun_out/r2c16
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5 ) 2>&1 | tee ${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
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
    print("%-12s 2^%s %-8s c=%d %.3f ms parity=%s pipe=%s stages %s" % (curve,logn," ".join(extra),d["engine"]["c"],d["ms_per_step"],d["parity"]["closed_form"],round(d["roofline"]["int_pipe"].get("multiplier_pipe",{}).get("frac",0),3),{k:round(v,2) for k,v in d["stages_ms"].items()}))
except Exception as e:
    print("%s 2^%s %s FAILED: %s" % (curve,logn,extra,(r.stderr or r.stdout)[-400:]))
PY
}
{
ab bls12377_g1 24; ab bls12377_g1 22; ab bls12377_g2 20; ab bls12377_g2 22
ab bls12381_g1 24; ab bn254_g1 25; ab bn254_g2 24; ab bw6761_g1 20
} | tee ${T}_sweeps.txt
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > ${T}_bench_n1.json 2> ${T}_bench_n1.err; echo "bench rc=$?"; tail -4 ${T}_bench_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2c16_bench_n1.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("metric","value","ms_per_step","n_gpus","gpu_launches")}, d["parity"]["closed_form"], d["engine"]["c"], "frac", round(d["roofline"]["frac"],4))
    for k in ("e2e","e2e_pageable","cpu_baseline","concurrent3"):
        v=d.get(k); print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a not in ("sample","path","what","stages_ms","note","workload","host_memory")} if v else None)
    for k,v in (d.get("configs") or {}).items():
        print(k, round(v["ms_per_step"],3), v["parity"]["closed_form"], "c=%d"%v["c"], {a:round(b,2) for a,b in v["stages_ms"].items()})
except Exception as e:
    print("no line:", e)
PY
