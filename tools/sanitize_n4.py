"""small MSMs of the N4 curves through every kernel path, for compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_n4.py
secp256k1 at c = 16 (17-bit last window), bw6-761 / bw6-633 (lane-parallel tail kernels by default, 48- / 40-byte scalars),
bls24-315 (80-byte points, 120-byte results: 8-byte load / store granules), the one-shot host call and resident bases."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("gnark_crypto_b200")
mx = importlib.import_module("gnark-crypto_b200.multiexp")
from oracle import cref  # noqa: E402
from tests.gpu_common import make_inputs  # noqa: E402

ok = True
for g, n, c in (("secp256k1_g1", 2000, 16), ("secp256k1_g1", 2000, 7), ("bw6761_g1", 500, 10), ("bw6761_g2", 300, 5),
                ("bls24315_g1", 1200, 11), ("bls24317_g1", 1200, 15), ("bw6633_g1", 600, 9), ("bw6633_g2", 400, 12)):
    pts, s = make_inputs(g, n, 3)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    eng = pkg.Engine(g, n, c=c)
    got = eng.msm_host_result(eng.to_device(pts), eng.to_device(s), n)
    eng.close()
    ok = ok and np.array_equal(got[: pts.shape[1]], want)
for g, n in (("bw6633_g1", 3000), ("bls24315_g1", 5000)):
    pts, s = make_inputs(g, n, 4, specials=False)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=8)
    A1, J1, A2, J2 = pkg.curve_package(g.split("_")[0])
    ok = ok and np.array_equal(A1().MultiExp(pts, s, pkg.MultiExpConfig()).limbs, want)
    rb = mx.ResidentBases(g, pts)
    ok = ok and np.array_equal(rb.MultiExp(s)[: pts.shape[1]], want)
    rb.Precompute(0)
    ok = ok and np.array_equal(rb.MultiExp(s)[: pts.shape[1]], want)
    rb.close()
    pkg.BatchScalarMultiplication(g, pts[0], s[:100])
print("SANITIZE_RUN_OK" if ok else "SANITIZE_RUN_MISMATCH")
