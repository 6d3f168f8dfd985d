"""Parity of the CUDA MultiExp path against the oracle, through the C ABI (host mirror in
gnark-crypto_b200/multiexp.py).  Mirrors ecc/bn254/multiexp_test.go: all-c agreement (:95-126), infinity / zero
inputs (:128-182), closed form (:186-216), cross test with duplicates and infinities in affine
(:221-299), G2 twins (:437-709).  Bit-exactness criterion: the affine normal form's limbs."""
import numpy as np
import pytest

from oracle import cref
from oracle import oracle as O
from tests.gpu_common import jac_to_affine_bytes, load_golden_msm, make_inputs

pytestmark = pytest.mark.gpu


def _pkg():
    import gnark_crypto_b200 as pkg

    return pkg


@pytest.fixture(params=["affine", "xyzz"])
def accumulate_mode(request, monkeypatch):
    """both bucket-accumulation paths: the extended-Jacobian segmented reduction (the engine's default; multiexp_jacobian.go's
    counterpart) and the batch-affine tree (GMSM_AFFINE=1, off by default; multiexp_affine.go's counterpart)"""
    monkeypatch.setenv("GMSM_AFFINE", "1" if request.param == "affine" else "0")
    return request.param


def _engine_msm(g, pts, s, c):
    pkg = _pkg()
    n = pts.shape[0]
    eng = pkg.Engine(g, max(n, 1), c=c)
    try:
        dp, ds = eng.to_device(pts), eng.to_device(s)
        jac = eng.msm_host_result(dp, ds, n)
        return jac, eng.c, eng.last_launches
    finally:
        eng.close()


@pytest.mark.parametrize("g,n", [("bn254_g1", 2000), ("bls12381_g1", 1200), ("bn254_g2", 1000), ("bls12381_g2", 500),
                                 ("bls12377_g1", 1000), ("bls12377_g2", 400)])
def test_all_window_sizes_agree_with_oracle(g, n, accumulate_mode):
    """every c the reference implements (2..16) plus the wider windows the GPU model may pick"""
    pts, s = make_inputs(g, n, 1234)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    cs = list(range(2, 17)) + [18, 21] if g == "bn254_g1" else [2, 5, 8, 11, 13, 16, 19]
    for c in cs:
        jac, used_c, launches = _engine_msm(g, pts, s, c)
        assert used_c == c and launches > 0
        w = pts.shape[1] // 2
        # output convention: (X, Y, One) or (0, 0, 0)
        G = O.GROUPS[g]
        assert np.array_equal(jac[2 * w :], np.array(G.K.encode(G.K.one), dtype=np.uint64))
        assert np.array_equal(jac[: 2 * w], want), c
        assert np.array_equal(jac_to_affine_bytes(g, jac), want), c


@pytest.mark.parametrize("g", ["bn254_g1", "bn254_g2", "bls12381_g1", "bls12381_g2", "bls12377_g1", "bls12377_g2"])
def test_committed_golden_vectors(g):
    """tests/golden/msm_vectors.json: the committed known answers (Python oracle, cross-checked against independent
    double-and-add when generated; the C port and the oracle are held to the same bytes by tests/test_cref.py)"""
    pts, s, want = load_golden_msm(g)
    w = pts.shape[1] // 2
    for c in (8, 13):
        jac, _, _ = _engine_msm(g, pts, s, c)
        assert np.array_equal(jac[: 2 * w], want), c
    if g == "bn254_g1":
        pkg = _pkg()
        assert np.array_equal(pkg.G1Affine().MultiExp(pts, s, pkg.MultiExpConfig()).limbs, want)


def test_config1_n65536_bn254_g1(accumulate_mode):
    """BASELINE.json configs[0]: bn254 G1, n = 2^16, random scalars; oracle = the reference algorithm
    single-threaded with the reference's own window choice (c = 13)."""
    g = "bn254_g1"
    n = 1 << 16
    pts, s = make_inputs(g, n, 0x5EED0001)
    want, _, used_c, leaves = cref.msm(g, pts, s, c=0, nthreads=1, nb_tasks=1)
    assert used_c == 13 and leaves == 1
    pkg = _pkg()
    out = pkg.G1Affine().MultiExp(pts, s, pkg.MultiExpConfig(NbTasks=1))
    assert np.array_equal(out.limbs, want)
    # resident-base API gives the same bytes
    from importlib import import_module

    mx = import_module("gnark-crypto_b200.multiexp")
    rb = mx.ResidentBases(g, pts)
    jac = rb.MultiExp(s)
    assert np.array_equal(jac[:8], want)
    # sub-range (kzg.Commit passes pk.G1[:len(p)])
    want2, _, _, _ = cref.msm(g, pts[100:5100], s[:5000], c=0, nthreads=4)
    jac2 = rb.MultiExp(s[:5000], offset=100)
    assert np.array_equal(jac2[:8], want2)
    rb.close()


@pytest.mark.parametrize("g", ["bn254_g1", "bn254_g2", "bls12381_g1", "bls12377_g1"])
def test_infinity_zero_and_empty(g, accumulate_mode):
    pkg = _pkg()
    A1, J1, A2, J2 = pkg.curve_package(g.split("_")[0])
    Jac = J1 if g.endswith("g1") else J2
    pts, s = make_inputs(g, 300, 5, specials=False)
    j = Jac().MultiExp(np.zeros_like(pts), s, pkg.MultiExpConfig())
    assert j.IsInfinity() and not j.limbs.any()      # multiexp_test.go:128-154 (Z == 0)
    j = Jac().MultiExp(pts, np.zeros_like(s), pkg.MultiExpConfig())
    assert j.IsInfinity() and not j.limbs.any()      # :156-182
    j = Jac().MultiExp(pts[:0], s[:0], pkg.MultiExpConfig())
    assert j.IsInfinity()
    # a single point, scalar one -> the point itself
    G = O.GROUPS[g]
    one = G.encode_scalars([1])
    j = Jac().MultiExp(pts[7:8], one, pkg.MultiExpConfig())
    assert np.array_equal(j.limbs[: pts.shape[1]], pts[7])


def test_closed_form_sum_of_squares():
    # multiexp_test.go:186-216: points [i]G, scalars i*mixer, 30 terms -> [9455*mixer]G
    g = "bn254_g1"
    G = O.GROUPS[g]
    mixer = 0x1234567890ABCDEF1234567890ABCDEF % G.fr.q
    base = G.encode_affine([G.gen])[0]
    pts = cref.generate_multiples(g, base, 1, 30)
    s = G.encode_scalars([(i + 1) * mixer % G.fr.q for i in range(30)])
    pkg = _pkg()
    out = pkg.G1Affine().MultiExp(pts, s, pkg.MultiExpConfig())
    assert np.array_equal(out.limbs, cref.scalar_mul(g, base, 9455 * mixer % G.fr.q))


@pytest.mark.parametrize("kind", ["smallvalues", "redundancy", "one_bucket", "all_equal_points"])
def test_skewed_scalar_distributions(kind, accumulate_mode):
    """the reference's benchmark distributions (multiexp_test.go:319-334) and harder skews: the
    chunked segmented reduction + carry levels must stay correct when buckets span many chunks"""
    g = "bn254_g1"
    G = O.GROUPS[g]
    n = 40000
    pts, s = make_inputs(g, n, 77, specials=False)
    if kind == "smallvalues":
        s[::5] = np.array([1, 0, 0, 0], dtype=np.uint64)           # limbs {1,0,0,0} in Montgomery form
    elif kind == "redundancy":
        for i in range(0, n, 100):
            s[i : i + 100] = s[i]
    elif kind == "one_bucket":
        s[:] = s[0]                                                # every digit of every scalar equal
    else:
        pts[:] = pts[3]                                            # doubling branch everywhere
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    for c in (8, 13, 16):
        jac, _, _ = _engine_msm(g, pts, s, c)
        assert np.array_equal(jac[:8], want), (kind, c)


@pytest.mark.parametrize("g,n", [("bn254_g1", 1 << 20), ("bls12381_g1", 1 << 18), ("bn254_g2", 1 << 17), ("bls12377_g1", 1 << 18), ("bls12377_g2", 1 << 16)])
def test_large_closed_form_on_device_bases(g, n, accumulate_mode):
    """size-independent property at large n: bases [i+1]B generated on the device, result must equal
    [sum (i+1) s_i mod r] B (the KZG TestCommit identity, kzg_test.go:209-239); also pins the device
    base generator against the oracle on a sample."""
    pkg = _pkg()
    G = O.GROUPS[g]
    base_pt = G.scalar_mul(G.gen, 0xC0FFEE)
    base = G.encode_affine([base_pt])[0]
    eng = pkg.Engine(g, n, c=0)
    try:
        d_pts = eng.generate_multiples(base, 1, n)
        w = pts_w = base.size
        host = d_pts.cpu().numpy().view(np.uint64).reshape(n, w)
        idx = [0, 1, 15, 16, 17, 4095, n // 2, n - 2, n - 1]
        for i in idx:
            assert np.array_equal(host[i], cref.scalar_mul(g, base, i + 1)), i
        s = cref.random_scalars(g, n, 4242)
        jac = eng.msm_host_result(d_pts, eng.to_device(s), n)
        k = cref.dot_index(g, s, 1)
        assert np.array_equal(jac[:w], cref.scalar_mul(g, base, k))
    finally:
        eng.close()


def test_window_sums_and_finalize_compose(accumulate_mode):
    """the multi-GPU decomposition on one device: split the inputs in 3 shards, per-shard window
    partials, finalize over the 3 'ranks' == MSM of the whole"""
    import torch

    pkg = _pkg()
    g = "bn254_g1"
    n = 30000
    pts, s = make_inputs(g, n, 31)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    eng = pkg.Engine(g, n, c=12)
    try:
        cuts = [0, 9000, 21000, n]
        parts = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            out = torch.zeros(eng.partials_bytes // 8, dtype=torch.int64, device="cuda")
            eng.window_sums(eng.to_device(pts[a:b]), eng.to_device(s[a:b]), b - a, out=out)
            parts.append(out)
        jac = eng.finalize(torch.cat(parts), 3).cpu().numpy().view(np.uint64)
        assert np.array_equal(jac[:8], want)
    finally:
        eng.close()


@pytest.mark.parametrize("g", ["bn254_g1", "bn254_g2", "bls12381_g1"])
def test_batch_scalar_multiplication_fixed_base(g):
    """next-row N1: BatchScalarMultiplicationG1/G2 (g1.go:1039-1118) -- same base, n scalars, affine out"""
    pkg = _pkg()
    G = O.GROUPS[g]
    base_pt = G.scalar_mul(G.gen, 0xBEEF)
    base = G.encode_affine([base_pt])[0]
    n = 600
    s = cref.random_scalars(g, n, 9)
    specials = [0, 1, 2, G.fr.q - 1, G.fr.q - 2, (1 << 200) + 12345]
    s[: len(specials)] = G.encode_scalars(specials)
    got = pkg.BatchScalarMultiplication(g, base, s)
    ks = G.decode_scalars(s)
    for i in list(range(len(specials))) + [17, 100, 333, 599]:
        assert np.array_equal(got[i], cref.scalar_mul(g, base, ks[i])), i
    assert not got[0].any()                                   # [0]B = infinity = (0, 0)
    assert np.array_equal(got[1], base)
    # consistency with MultiExp: sum_i [s_i]B == MultiExp(got, ones) == [sum s_i]B
    tot = sum(ks) % G.fr.q
    A1, J1, A2, J2 = pkg.curve_package(g.split("_")[0])
    Aff = A1 if g.endswith("g1") else A2
    res = Aff().MultiExp(got, G.encode_scalars([1] * n), pkg.MultiExpConfig())
    assert np.array_equal(res.limbs, cref.scalar_mul(g, base, tot))


@pytest.mark.parametrize("g,n", [("bn254_g1", (1 << 19) + 3), ("bn254_g1", (1 << 21) + 17), ("bls12381_g1", (1 << 18) + 1),
                                 ("bn254_g2", (1 << 18) + 5)])
def test_pipelined_host_calls_closed_form(g, n):
    """the host entry points cut large inputs into batches that share one bucket array (H2D of batch k+1
    under the bucket pass of batch k; scratch buckets + merge): one-shot gmsm_multiexp and resident bases
    must both give [sum (i+1) s_i] B, and ragged n must work"""
    from importlib import import_module

    pkg = _pkg()
    mx = import_module("gnark-crypto_b200.multiexp")
    G = O.GROUPS[g]
    base = G.encode_affine([G.scalar_mul(G.gen, 0xABCDEF)])[0]
    eng = pkg.Engine(g, n, c=0)
    try:
        w = base.size
        pts = eng.generate_multiples(base, 1, n).cpu().numpy().view(np.uint64).reshape(n, w)
    finally:
        eng.close()
    s = cref.random_scalars(g, n, 2024)
    want = cref.scalar_mul(g, base, cref.dot_index(g, s, 1))
    A1, J1, A2, J2 = pkg.curve_package(g.split("_")[0])
    Aff = A1 if g.endswith("g1") else A2
    got = Aff().MultiExp(pts, s, pkg.MultiExpConfig())
    assert np.array_equal(got.limbs, want)
    rb = mx.ResidentBases(g, pts)
    try:
        assert np.array_equal(rb.MultiExp(s)[:w], want)
        # a second call on the same handle with fewer scalars (context reuse / shrink path)
        m = n // 3
        want2 = cref.scalar_mul(g, base, cref.dot_index(g, s[:m], 1))
        assert np.array_equal(rb.MultiExp(s[:m])[:w], want2)
    finally:
        rb.close()


@pytest.mark.parametrize("n", [1, 2, 3, 31, 33, 257])
def test_tiny_and_ragged_sizes(n):
    g = "bn254_g1"
    pkg = _pkg()
    pts, s = make_inputs(g, max(n, 64), 5, specials=False)
    pts, s = pts[:n], s[:n]
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=1)
    got = pkg.G1Affine().MultiExp(pts, s, pkg.MultiExpConfig())
    assert np.array_equal(got.limbs, want)


def test_kzg_commit_over_generated_srs_and_dump_roundtrip(tmp_path):
    """next-row N2 on top of N1: SRS = [alpha^i]G by BatchScalarMultiplication (kzg.NewSRS, kzg.go:129),
    raw dump (marker + unsafe.WriteSlice image) -> resident bases, Commit(f) == [f(alpha)]G
    (TestCommit, ecc/bn254/kzg/kzg_test.go:209-239)"""
    from importlib import import_module

    kzg = import_module("gnark-crypto_b200.kzg")
    g = "bn254_g1"
    G = O.GROUPS[g]
    r = G.fr.q
    size, alpha = 3000, 0x1234567890ABCDEF1234567
    gen = G.encode_affine([G.gen])[0]
    srs = kzg.new_srs_g1("bn254", size, alpha, gen, r, G.encode_scalars)
    assert np.array_equal(srs[0], gen) and np.array_equal(srs[1], cref.scalar_mul(g, gen, alpha))
    path = tmp_path / "srs.dump"
    with open(path, "wb") as f:
        f.write(b"\x00" * 40)                      # stands in for the verifying-key prefix
        kzg.write_marker(f)
        kzg.write_slice(f, srs)
    with open(path, "rb") as f:
        f.seek(40)
        pk = kzg.ProvingKey.from_dump("bn254", f, max_pk_points=2048)
    assert pk.G1.shape == (2048, 8) and np.array_equal(pk.G1, srs[:2048])
    rng = np.random.default_rng(3)
    coeffs = [int(x) for x in rng.integers(0, 2**62, size=2000)]
    f_alpha = sum(c * pow(alpha, i, r) for i, c in enumerate(coeffs)) % r
    digest = kzg.Commit(G.encode_scalars(coeffs), pk)
    assert np.array_equal(digest, cref.scalar_mul(g, gen, f_alpha))
    # kzg.Open (kzg.go:180-204): ClaimedValue = f(a), H = [h(alpha)]G with h = (f - f(a)) / (X - a)
    a = 0xDEADBEEF12345
    proof = kzg.Open(G.encode_scalars(coeffs), G.encode_scalars([a])[0], pk)
    f_a = sum(c * pow(a, i, r) for i, c in enumerate(coeffs)) % r
    assert np.array_equal(proof.ClaimedValue, G.encode_scalars([f_a])[0])
    h_alpha = (f_alpha - f_a) * pow(alpha - a, -1, r) % r
    assert np.array_equal(proof.H, cref.scalar_mul(g, gen, h_alpha))
    with pytest.raises(kzg.ErrInvalidPolynomialSize):
        kzg.Open(G.encode_scalars([5]), G.encode_scalars([a])[0], pk)      # constant polynomial: empty quotient
    # the same key with window tables (static SRS): identical digest
    pk2 = kzg.ProvingKey("bn254", srs[:2048], window_tables=True)
    assert np.array_equal(kzg.Commit(G.encode_scalars(coeffs), pk2), digest)
    pk2.close()
    with pytest.raises(kzg.ErrInvalidPolynomialSize):
        kzg.Commit(G.encode_scalars([1] * 2049), pk)
    with pytest.raises(kzg.ErrInvalidPolynomialSize):
        kzg.Commit(np.zeros((0, 4), dtype=np.uint64), pk)
    pk.close()


def test_concurrent_callers_are_safe():
    """BenchmarkManyMultiExpG1Reference (multiexp_test.go:385-415) launches 3 MSMs from 3 goroutines; the C ABI
    must be callable from several OS threads at once (ctypes drops the GIL during the call)"""
    import threading

    pkg = _pkg()
    cases = []
    for i, (g, n) in enumerate([("bn254_g1", 30000), ("bn254_g1", 70000), ("bn254_g2", 9000), ("bls12381_g1", 20000)]):
        pts, s = make_inputs(g, n, 100 + i)
        want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
        cases.append((g, pts, s, want))
    results = [None] * (3 * len(cases))

    def work(k):
        g, pts, s, want = cases[k % len(cases)]
        A1, J1, A2, J2 = pkg.curve_package(g.split("_")[0])
        Aff = A1 if g.endswith("g1") else A2
        ok = True
        for _ in range(3):
            ok = ok and np.array_equal(Aff().MultiExp(pts, s, pkg.MultiExpConfig()).limbs, want)
        results[k] = ok

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(results))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert all(results), results
