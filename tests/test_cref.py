"""The C port of the oracle (oracle/msm_ref.c) against the pinned Python oracle.  CPU only."""
import random

import numpy as np
import pytest

from oracle import cref
from oracle import oracle as O

FIELD_IDS = {"bn254_fp": 0, "bn254_fr": 1, "bls12381_fp": 2, "bls12381_fr": 3, "bls12377_fp": 4, "bls12377_fr": 5,
             "secp256k1_fp": 6, "secp256k1_fr": 7, "bw6761_fp": 8, "bw6761_fr": 9, "bls24315_fp": 10, "bls24315_fr": 11,
             "bls24317_fp": 12, "bls24317_fr": 13, "bw6633_fp": 14, "bw6633_fr": 15}


def _limbs(f, vals):
    return np.array([f.to_limbs(v) for v in vals], dtype=np.uint64)


@pytest.mark.parametrize("name", list(FIELD_IDS))
def test_field_ops(name):
    f = O.FIELDS[name]
    rng = random.Random(3)
    vals = [0, 1, f.q - 1, f.Rmod, f.R2] + [rng.randrange(f.q) for _ in range(60)]
    a = vals
    b = list(reversed(vals))
    A, B = _limbs(f, a), _limbs(f, b)
    got = cref.field_op(FIELD_IDS[name], 0, A, B)
    assert [f.from_limbs(r) for r in got] == [x * y * f.Rinv % f.q for x, y in zip(a, b)]
    got = cref.field_op(FIELD_IDS[name], 1, A, B)
    assert [f.from_limbs(r) for r in got] == [(x + y) % f.q for x, y in zip(a, b)]
    got = cref.field_op(FIELD_IDS[name], 2, A, B)
    assert [f.from_limbs(r) for r in got] == [(x - y) % f.q for x, y in zip(a, b)]
    got = cref.field_op(FIELD_IDS[name], 3, A, B)
    assert [f.from_limbs(r) for r in got] == [(-x) % f.q for x in a]
    got = cref.field_op(FIELD_IDS[name], 4, A, B)
    # Montgomery inverse: inv(xR) = x^-1 R  ->  as integers: got * x == R^2 mod q
    for x, g in zip(a, got):
        g = f.from_limbs(g)
        assert (g == 0 and x == 0) or (g * x) % f.q == f.R2 % f.q


@pytest.mark.parametrize("g", list(O.GROUPS))
def test_random_scalars_and_digits_match_python(g):
    G = O.GROUPS[g]
    s = cref.random_scalars(g, 40, 0x5EED0000 + 7)
    py = O.random_scalars_mont(G.fr, 40, 0x5EED0000 + 7)
    assert [O.Field.from_limbs(r) for r in s] == py
    s[3, :] = 0
    py[3] = 0
    for c in (2, 5, 11, 13, 16, 21):
        d = cref.partition_scalars(g, s, c)
        assert np.array_equal(d, O.partition_scalars(G.fr, py, c)), c


@pytest.mark.parametrize("g", list(O.GROUPS))
def test_generate_multiples_and_scalar_mul(g):
    G = O.GROUPS[g]
    base = G.encode_affine([G.gen])[0]
    pts = cref.generate_multiples(g, base, 1, 700, nthreads=3)
    want = O.consecutive_multiples(G, 12)
    assert G.decode_affine(pts[:12]) == want
    # spot checks deep in the array, across thread / lane boundaries
    for idx in (255, 256, 257, 233, 466, 699):
        assert G.decode_affine(pts[idx : idx + 1])[0] == G.scalar_mul(G.gen, idx + 1), idx
    k = random.Random(1).randrange(G.fr.q)
    assert G.decode_affine(cref.scalar_mul(g, base, k).reshape(1, -1))[0] == G.scalar_mul(G.gen, k)
    # start = 0 -> first point is infinity; exercises the special-case path
    p0 = cref.generate_multiples(g, base, 0, 5)
    assert G.decode_affine(p0) == [G.aff_inf()] + O.consecutive_multiples(G, 4)


@pytest.mark.parametrize("g,n,cs", [("bn254_g1", 300, [2, 5, 8, 13, 16, 18]), ("bls12381_g1", 120, [4, 16]),
                                     ("bn254_g2", 100, [5, 16]), ("bls12381_g2", 40, [7]), ("bls12377_g1", 120, [5, 16]), ("bls12377_g2", 60, [6])])
def test_msm_matches_python_oracle(g, n, cs):
    G = O.GROUPS[g]
    base = G.encode_affine([G.gen])[0]
    pts = cref.generate_multiples(g, base, 1, n)
    s = cref.random_scalars(g, n, 42)
    # ingredients of the reference's cross test (multiexp_test.go:233-245): infinity points,
    # duplicated (point, scalar) pairs, zero scalars; plus P / -P with equal scalars
    pts[5, :] = 0
    pts[17, :] = 0
    s[9, :] = 0
    pts[20:26] = pts[30:36]
    s[20:26] = s[30:36]
    neg = G.decode_affine(pts[40:43])
    pts[43:46] = G.encode_affine([G.aff_neg(p) for p in neg])
    s[43:46] = s[40:43]
    py_pts = G.decode_affine(pts)
    py_s = [O.Field.from_limbs(r) for r in s]
    want = O.multi_exp(G, py_pts, py_s, c=8)
    want_enc = G.encode_affine([want])[0]
    for c in cs:
        aff, jac, uc, nl = cref.msm(g, pts, s, c=c, nthreads=2)
        assert np.array_equal(aff, want_enc), c
        assert uc == c and nl == 1
        assert G.jac_to_affine(G.decode_jac(jac)) == want
    # bestC + split recursion (multiexp.go:95-140) with many tasks
    aff, _, uc, nl = cref.msm(g, pts, s, c=0, nthreads=4, nb_tasks=128)
    assert np.array_equal(aff, want_enc)
    assert nl >= (2 if n >= 120 else 1)


def test_msm_closed_form_large():
    # n = 2^14 like the reference cross test; closed form [sum (i+1) s_i] G
    g = "bn254_g1"
    G = O.GROUPS[g]
    n = 1 << 14
    base = G.encode_affine([G.gen])[0]
    pts = cref.generate_multiples(g, base, 1, n, nthreads=4)
    s = cref.random_scalars(g, n, 99)
    aff, _, uc, _ = cref.msm(g, pts, s, c=0, nthreads=4, nb_tasks=1)
    assert uc == O.best_c(254, n)
    k = cref.dot_index(g, s, 1)
    ks = G.decode_scalars(s[:50])
    assert k == (sum((i + 1) * ks[i] for i in range(50)) + cref.dot_index(g, s[50:], 51)) % G.fr.q
    assert np.array_equal(aff, cref.scalar_mul(g, base, k))
    assert G.decode_affine(aff.reshape(1, -1))[0] == G.scalar_mul(G.gen, k)


def test_msm_edge_cases():
    g = "bn254_g1"
    G = O.GROUPS[g]
    base = G.encode_affine([G.gen])[0]
    pts = cref.generate_multiples(g, base, 1, 20)
    s = cref.random_scalars(g, 20, 1)
    aff, jac, _, _ = cref.msm(g, np.zeros_like(pts), s, c=4)
    assert not aff.any() and not jac.any()
    aff, jac, _, _ = cref.msm(g, pts, np.zeros_like(s), c=4)
    assert not aff.any() and not jac.any()
    aff, jac, _, _ = cref.msm(g, pts[:0], s[:0], c=0)
    assert not aff.any() and not jac.any()


@pytest.mark.parametrize("g,n", [("bn254_g1", 6000), ("bls12381_g1", 3000), ("bn254_g2", 2500), ("bls12377_g2", 2500)])
def test_msm_batch_affine_processor(g, n):
    """processChunkG1BatchAffine (multiexp_affine.go:24-231) restated in the port: same result as the extended-Jacobian
    processor and as the closed form, with the ingredients that hit its special cases (duplicates -> doubling in the
    fallback buckets, P / -P -> bucket back to infinity, conflicts -> queue)"""
    G = O.GROUPS[g]
    base = G.encode_affine([G.gen])[0]
    pts = cref.generate_multiples(g, base, 1, n, nthreads=4)
    s = cref.random_scalars(g, n, 4242)
    pts[7, :] = 0
    s[11, :] = 0
    pts[100:400] = pts[1000:1300]            # same (point, scalar) twice: same bucket, equal points
    s[100:400] = s[1000:1300]
    neg = G.decode_affine(pts[500:520])
    pts[520:540] = G.encode_affine([G.aff_neg(p) for p in neg])
    s[520:540] = s[500:520]
    s[600:900] = s[600]                      # one bucket per window hit 300 times: queue + fallback buckets
    try:
        cref.set_batch_affine(False)
        want, _, _, _ = cref.msm(g, pts, s, c=8, nthreads=4)
        for c in (10, 11, 12, 13, 14, 15, 16):
            cref.set_batch_affine(False)
            a0, _, _, _ = cref.msm(g, pts, s, c=c, nthreads=4)
            cref.set_batch_affine(True)
            a1, _, _, _ = cref.msm(g, pts, s, c=c, nthreads=4)
            assert np.array_equal(a0, want) and np.array_equal(a1, want), c
    finally:
        cref.set_batch_affine(True)
    # anchor on the Python oracle through a small prefix as well as the whole vector by closed form (G1 only: cheap)
    if g == "bn254_g1":
        py = O.multi_exp(G, G.decode_affine(pts[:400]), [O.Field.from_limbs(r) for r in s[:400]], c=8)
        a, _, _, _ = cref.msm(g, pts[:400], s[:400], c=10, nthreads=2)
        assert np.array_equal(a, G.encode_affine([py])[0])


@pytest.mark.parametrize("g", list(O.GROUPS))
def test_golden_msm_vectors(g):
    """the committed MultiExp known answers (tests/golden/msm_vectors.json) against the C port at the reference's own
    window choice and at forced widths (batch-affine and extended-Jacobian processors), and against the Python oracle
    at a width different from the one that generated them"""
    from tests.gpu_common import load_golden_msm

    pts, s, want = load_golden_msm(g)
    G = O.GROUPS[g]
    assert pts.shape == (96, G.aff_words) and s.shape == (96, G.fr.limbs)
    for c in (0, 4, 9, 12, 16):
        got, _, _, _ = cref.msm(g, pts, s, c=c, nthreads=2)
        assert np.array_equal(got, want), c
    py = O.multi_exp(G, G.decode_affine(pts), [O.Field.from_limbs(r) for r in s], c=11)
    assert np.array_equal(G.encode_affine([py])[0], want)
