"""Generates tests/golden/msm_vectors.json: small MultiExp known answers for every group of oracle.GROUPS.

The reference holds no golden MSM output and cannot run here (Go), so these vectors are produced by the
pinned Python oracle (oracle/oracle.py: restatement of ecc/<curve>/multiexp.go, pinned to the reference's constants and
hash-to-curve known answers by tests/test_oracle.py) and cross-checked at generation time against the independent
double-and-add sum `msm_naive`.  They freeze the oracle: the C port, the Python oracle at other window widths and the
CUDA engine are all compared with the same committed bytes.

    python tests/golden/make_msm_golden.py        (about a minute)

Inputs follow the reference's cross test (multiexp_test.go:221-272): consecutive multiples of the generator with
infinity points, duplicated (point, scalar) pairs, P / -P with equal scalars, zero and extreme scalars."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402

N = 96


def build(g):
    G = O.GROUPS[g]
    pts = O.consecutive_multiples(G, N, start_k=2)
    sm = O.random_scalars_mont(G.fr, N, 0x601D + list(O.GROUPS).index(g))
    ks = [G.fr.from_mont(s) for s in sm]
    pts[5] = G.aff_inf()
    pts[N - 1] = G.aff_inf()
    ks[9] = 0
    ks[10] = 1
    ks[11] = G.fr.q - 1
    ks[12] = (1 << (G.fr.bits - 1)) + 12345
    pts[20:26] = pts[30:36]
    ks[20:26] = ks[30:36]
    pts[43:46] = [G.aff_neg(p) for p in pts[40:43]]
    ks[43:46] = ks[40:43]
    sm = [G.fr.to_mont(k) for k in ks]
    want = O.multi_exp(G, pts, sm, c=5)
    assert want == O.msm_naive(G, pts, ks), g
    assert want == O.multi_exp_tables(G, pts, sm, 7), g
    hexrow = lambda row: ["%016x" % int(x) for x in row]
    return {
        "points": [hexrow(r) for r in G.encode_affine(pts)],
        "scalars": [hexrow(r) for r in G.encode_scalars(ks)],
        "result_affine": hexrow(G.encode_affine([want])[0]),
    }


if __name__ == "__main__":
    out = {"note": "u64 little-endian limbs in Go memory layout (Montgomery form), hex; see make_msm_golden.py",
           "groups": {g: build(g) for g in O.GROUPS}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "msm_vectors.json")
    json.dump(out, open(path, "w"), indent=0)
    print("wrote", path, os.path.getsize(path), "bytes")
