"""small MSMs through every kernel path, for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("gnark_crypto_b200")
from oracle import cref  # noqa: E402
from tests.gpu_common import make_inputs  # noqa: E402

ok = True
for mode in ("0", "1"):
    os.environ["GMSM_AFFINE"] = mode
    for g, n, c in (("bn254_g1", 3000, 8), ("bn254_g1", 3000, 13), ("bn254_g2", 700, 7), ("bls12381_g1", 900, 9)):
        pts, s = make_inputs(g, n, 3)
        want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
        eng = pkg.Engine(g, n, c=c)
        got = eng.msm_host_result(eng.to_device(pts), eng.to_device(s), n)
        eng.close()
        ok = ok and np.array_equal(got[: pts.shape[1]], want)
os.environ["GMSM_AFFINE"] = "0"
pts, s = make_inputs("bn254_g1", 300000, 4, specials=False)   # 2 pipelined batches (scratch buckets + merge)
want, _, _, _ = cref.msm("bn254_g1", pts, s, c=0, nthreads=8)
ok = ok and np.array_equal(pkg.G1Affine().MultiExp(pts, s, pkg.MultiExpConfig()).limbs, want)
G = pkg.BatchScalarMultiplication("bn254_g1", pts[0], s[:200])
print("SANITIZE_RUN_OK" if ok else "SANITIZE_RUN_MISMATCH")
