#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fft.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -8
timeout 300 python - <<'PY'
import importlib, time, numpy as np, torch, sys
sys.path.insert(0,'.')
import gnark_crypto_b200
fft = importlib.import_module("gnark-crypto_b200.fft")
for logn in (20, 24):
    n = 1 << logn
    d = fft.NewDomain("bn254", n)
    a = torch.randint(0, 2**59, (n*4,), dtype=torch.int64, device="cuda")
    for _ in range(3): d.fft_device(a, False, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): d.fft_device(a, False, 1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    print("bn254 fr FFT DIF n=2^%d: %.3f ms  (%.1f GB/s algorithmic at %d passes of 64 B/elt)" % (logn, ms, (logn-10+1)*n*64/ms/1e6, logn-10+1))
    d.close()
PY
