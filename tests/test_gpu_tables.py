"""Window-table mode for resident bases (gmsm_bases_precompute / gmsm_ctx_create_tables): row j of the table is
2^(c*j) * bases and all windows share one bucket set.  The results must be bit-identical to the reference's
MultiExp (oracle), whatever the table width; the table rows themselves are pinned against the oracle's scalar
multiplication.  Mirrors the cases of ecc/bn254/multiexp_test.go the plain path is tested with (all-c agreement
:95-126, infinities / zero scalars :128-182, duplicates and cancellation :221-299, skewed digits :319-334)."""
from importlib import import_module

import numpy as np
import pytest

from oracle import cref
from oracle import oracle as O
from tests.gpu_common import make_inputs

pytestmark = pytest.mark.gpu


def _pkg():
    import gnark_crypto_b200 as pkg

    return pkg


def _tables_msm(g, pts, s, c, offset=0, m=None):
    """device-level table path: build the table of all of pts, MSM of s[:m] over pts[offset:offset+m]"""
    pkg = _pkg()
    n = pts.shape[0]
    m = n - offset if m is None else m
    eng = pkg.Engine(g, max(n, 1), c=c, tables=True)
    try:
        dp, ds = eng.to_device(pts), eng.to_device(s[:m])
        tab = eng.build_tables(dp, n)
        jac = eng.msm_tables(tab, n, ds, m, offset=offset).cpu().numpy().view(np.uint64).copy()
        return jac, eng.c, eng.nwin, tab.cpu().numpy().view(np.uint64).reshape(eng.nwin, n, -1), eng.last_launches
    finally:
        eng.close()


@pytest.mark.parametrize("g,n,cs", [("bn254_g1", 2000, [2, 5, 8, 11, 13, 16, 19, 22]), ("bls12381_g1", 1200, [5, 15, 22]),
                                    ("bn254_g2", 1000, [8, 17]), ("bls12381_g2", 500, [13]), ("bls12377_g1", 1000, [16, 21]),
                                    ("bls12377_g2", 400, [12])])
def test_tables_every_width_agrees_with_oracle(g, n, cs):
    pts, s = make_inputs(g, n, 4321)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    G = O.GROUPS[g]
    w = pts.shape[1] // 2
    for c in cs:
        jac, used_c, nwin, tab, launches = _tables_msm(g, pts, s, c)
        assert used_c == c and launches > 0
        assert np.array_equal(jac[: 2 * w], want), c
        assert np.array_equal(jac[2 * w :], np.array(G.K.encode(G.K.one), dtype=np.uint64))
        # the table itself: row 0 = the bases, row j = 2^(c*j) * bases (infinity stays (0, 0))
        assert np.array_equal(tab[0], pts)
        for j in sorted({1, nwin // 2, nwin - 1}):
            if j == 0 or j >= nwin:
                continue
            for i in (0, 5, 77, n - 2):
                if not pts[i].any():
                    assert not tab[j, i].any()
                else:
                    assert np.array_equal(tab[j, i], cref.scalar_mul(g, pts[i], (1 << (c * j)) % G.fr.q)), (c, j, i)


def test_tables_subrange_and_ragged_sizes():
    """kzg.Commit passes pk.G1[:len(p)] (kzg.go:159-176): MSMs over a prefix / an inner range of the table"""
    g = "bn254_g1"
    pts, s = make_inputs(g, 5000, 99)
    for off, m in [(0, 5000), (0, 1), (100, 3333), (4999, 1), (17, 0)]:
        jac, _, _, _, _ = _tables_msm(g, pts, s, 12, offset=off, m=m)
        if m == 0:
            assert not jac.any()
            continue
        want, _, _, _ = cref.msm(g, pts[off : off + m], s[:m], c=0, nthreads=4)
        assert np.array_equal(jac[:8], want), (off, m)


@pytest.mark.parametrize("kind", ["smallvalues", "redundancy", "one_bucket", "all_equal_points", "all_infinity", "zero_scalars"])
def test_tables_skewed_distributions(kind):
    g = "bn254_g1"
    n = 40000
    pts, s = make_inputs(g, n, 78, specials=False)
    if kind == "smallvalues":
        s[::5] = np.array([1, 0, 0, 0], dtype=np.uint64)
    elif kind == "redundancy":
        for i in range(0, n, 100):
            s[i : i + 100] = s[i]
    elif kind == "one_bucket":
        s[:] = s[0]
    elif kind == "all_equal_points":
        pts[:] = pts[3]
    elif kind == "all_infinity":
        pts[:] = 0
    else:
        s[:] = 0
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    for c in (9, 16, 20):
        jac, _, _, _, _ = _tables_msm(g, pts, s, c)
        assert np.array_equal(jac[:8], want), (kind, c)
        if kind in ("all_infinity", "zero_scalars"):
            assert not jac.any()


def test_plain_entry_points_refuse_a_table_context():
    pkg = _pkg()
    mx = import_module("gnark-crypto_b200.multiexp")
    g = "bn254_g1"
    pts, s = make_inputs(g, 300, 1)
    eng = pkg.Engine(g, 300, c=10, tables=True)
    try:
        with pytest.raises(mx.MultiExpError):
            eng.msm(eng.to_device(pts), eng.to_device(s), 300)
    finally:
        eng.close()


@pytest.mark.parametrize("g,n,c", [("bn254_g1", (1 << 16) + 5, 0), ("bn254_g1", (1 << 19) + 3, 0), ("bn254_g1", (1 << 18) + 1, 20),
                                   ("bls12381_g1", (1 << 18) + 7, 0), ("bn254_g2", (1 << 17) + 2, 0)])
def test_resident_bases_precompute_matches_plain_and_closed_form(g, n, c):
    """gmsm_bases_precompute through the host API: same bytes as before the precomputation and as the closed form
    [sum (i+1) s_i] B; large n runs the pipelined batches (scratch buckets + merge) over the shared bucket set"""
    pkg = _pkg()
    mx = import_module("gnark-crypto_b200.multiexp")
    G = O.GROUPS[g]
    base = G.encode_affine([G.scalar_mul(G.gen, 0xFACE)])[0]
    eng = pkg.Engine(g, n, c=0)
    try:
        w = base.size
        pts = eng.generate_multiples(base, 1, n).cpu().numpy().view(np.uint64).reshape(n, w)
    finally:
        eng.close()
    pts[11, :] = 0
    s = cref.random_scalars(g, n, 777)
    s[13, :] = 0
    k = (cref.dot_index(g, s, 1) - 12 * G.decode_scalars(s[11:12])[0]) % G.fr.q
    want = cref.scalar_mul(g, base, k)
    rb = mx.ResidentBases(g, pts)
    try:
        before = rb.MultiExp(s)
        assert np.array_equal(before[:w], want)
        used = rb.Precompute(c)
        assert used == c or (c == 0 and 6 <= used <= 24)
        after = rb.MultiExp(s)
        assert np.array_equal(after, before)
        # sub-range and a smaller call on the same handle
        m = n // 3
        want2, _, _, _ = cref.msm(g, pts[100 : 100 + m], s[:m], c=0, nthreads=8)
        assert np.array_equal(rb.MultiExp(s[:m], offset=100)[:w], want2)
        with pytest.raises(mx.MultiExpError):
            rb.Precompute(c)          # already built
    finally:
        rb.close()


def test_sharded_resident_bases_with_tables(monkeypatch):
    """bases sharded over the devices of GMSM_DEVICES (device = -1), here two shards on device 0 so that the path runs on a
    one-GPU box: one host thread per shard, every shard returns ONE partial in window-table mode, joined on the first"""
    mx = import_module("gnark-crypto_b200.multiexp")
    g = "bn254_g1"
    n = (1 << 17) + 5
    pts, s = make_inputs(g, n, 12)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=8)
    monkeypatch.setenv("GMSM_DEVICES", "0,0")
    rb = mx.ResidentBases(g, pts, device=-1)
    try:
        w = pts.shape[1]
        assert np.array_equal(rb.MultiExp(s)[:w], want)
        assert rb.Precompute(0) >= 6
        assert np.array_equal(rb.MultiExp(s)[:w], want)
        lo, m = n // 5, n // 2                      # straddles the shard boundary
        want_sub, _, _, _ = cref.msm(g, pts[lo : lo + m], s[:m], c=0, nthreads=8)
        assert np.array_equal(rb.MultiExp(s[:m], offset=lo)[:w], want_sub)
        m2 = n // 4                                 # inside the first shard only (single-job path)
        want_one, _, _, _ = cref.msm(g, pts[3 : 3 + m2], s[:m2], c=0, nthreads=8)
        assert np.array_equal(rb.MultiExp(s[:m2], offset=3)[:w], want_one)
    finally:
        rb.close()
