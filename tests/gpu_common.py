"""helpers shared by the -m gpu tests"""
import json
import os

import numpy as np

from oracle import cref
from oracle import oracle as O


def make_inputs(g, n, seed, specials=True, start=1):
    """on-curve points [start+i]G in Go layout + random Montgomery scalars, with the ingredients of the
    reference's cross test (multiexp_test.go:233-245): infinity points, duplicated (point, scalar)
    pairs (doubling branch), zero scalars, P / -P with equal scalars (cancellation branch)."""
    G = O.GROUPS[g]
    base = G.encode_affine([G.gen])[0]
    pts = cref.generate_multiples(g, base, start, n, nthreads=4)
    s = cref.random_scalars(g, n, seed)
    if specials and n >= 64:
        pts[5, :] = 0
        pts[17, :] = 0
        pts[n - 1, :] = 0
        s[9, :] = 0
        s[n - 2, :] = 0
        pts[20:26] = pts[30:36]
        s[20:26] = s[30:36]
        neg = G.decode_affine(pts[40:43])
        pts[43:46] = G.encode_affine([G.aff_neg(p) for p in neg])
        s[43:46] = s[40:43]
    return pts, s


def jac_to_affine_bytes(g, jac):
    """FromJacobian (g1.go:150-166) of the engine's output, via the oracle -> u64 affine limbs"""
    G = O.GROUPS[g]
    return G.encode_affine([G.jac_to_affine(G.decode_jac(jac))])[0]


_GOLD = None


def load_golden_msm(g):
    """tests/golden/msm_vectors.json (made by tests/golden/make_msm_golden.py): points, scalars, expected affine result
    as uint64 arrays in Go memory layout"""
    global _GOLD
    if _GOLD is None:
        _GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "msm_vectors.json")))["groups"]
    v = _GOLD[g]
    conv = lambda rows: np.array([[int(x, 16) for x in r] for r in rows], dtype=np.uint64)
    return conv(v["points"]), conv(v["scalars"]), conv([v["result_affine"]])[0]
