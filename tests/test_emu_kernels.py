"""The product's CUDA kernels (gnark-crypto_b200/csrc/kernels.cuh, the file nvcc compiles) run on the CPU, one emulated
thread at a time, in the engine's launch order (tests/emu/emu_engine.cpp over the stand-in tests/emu/cuda_runtime.h), and
compared with the oracle.  This checks the kernel-level logic without a GPU: digit recoding + histogram, the per-window
and the shared (window-table) scatter, the chunked segmented reduction with its owner / carry rule and the two-part
launch, the carry levels, the segment reduction, the finalize and the table level -- at chunk lengths, carry run
lengths, segment lengths and pass counts the engine's own heuristics would not pick.  The device arithmetic itself
(PTX carry chains) is covered by tests/test_hostcheck.py; the real launches by the -m gpu tests.  Kernels with barriers or
warp shuffles (the K1b scans, k_finalize, the product scans of the batch-affine pass) run under a cooperative launcher:
one ucontext fiber per CUDA thread.  One thread runs at a time, so this finds logic errors, not data races (compute-sanitizer racecheck on the GPU does that:
profiles/r01_compute_sanitizer_racecheck.log).  CPU only; a test artefact (build/libgmsm_emu.so), never part of libgmsm.so."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import cref
from oracle import oracle as O
from tests.gpu_common import load_golden_msm, make_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gnark-crypto_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")
OUT = os.path.join(ROOT, "gnark-crypto_b200", "build", "libgmsm_emu.so")
GROUPS = list(O.GROUPS)
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        bdir = os.path.dirname(OUT)
        os.makedirs(bdir, exist_ok=True)
        deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [
            os.path.join(EMU, f) for f in os.listdir(EMU)]
        if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps):
            objs, procs = [], []
            for k in range(len(GROUPS)):
                o = os.path.join(bdir, "emu_%d.o" % k)
                objs.append(o)
                # tests/emu FIRST: its cuda_runtime.h stands in for the real one
                procs.append(subprocess.Popen(["g++", "-std=c++17", "-O1", "-fPIC", "-DEMU_GROUP=%d" % k, "-I", EMU, "-I", CSRC, "-c",
                                               os.path.join(EMU, "emu_engine.cpp"), "-o", o]))
            o = os.path.join(bdir, "emu_fft.o")
            objs.append(o)
            procs.append(subprocess.Popen(["g++", "-std=c++17", "-O1", "-fPIC", "-I", EMU, "-I", CSRC, "-c", os.path.join(EMU, "emu_fft.cpp"), "-o", o]))
            assert all(p.wait() == 0 for p in procs)
            subprocess.run(["g++", "-shared", "-o", OUT, *objs], check=True)
        _LIB = ctypes.CDLL(OUT)
    return _LIB


def emu_msm(g, pts, s, c, tables=0, K=16, K2_first=4, K2=16, L=32, passes=4, split=0, batches=1, order=0, mode=0):
    """order: block execution order of every emulated launch (0 ascending, 1 descending, >= 2 pseudo-random)
    mode : bit 0 = the real K1b scan kernels (cooperative launch: fibers with barriers and warp shuffles) instead of a
           host scan; bit 1 = batch-affine bucket accumulation (affine_kernels.cuh) instead of k_accumulate;
           bit 2 = every launch through the cooperative launcher"""
    getattr(_lib(), "emu_set_block_order_%d" % GROUPS.index(g))(order)
    pts = np.ascontiguousarray(pts, dtype=np.uint64)
    s = np.ascontiguousarray(s, dtype=np.uint64)
    w = pts.shape[1] // 2
    out = np.zeros(3 * w, dtype=np.uint64)
    fn = getattr(_lib(), "emu_msm_%d" % GROUPS.index(g))
    rc = fn(pts.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(pts.shape[0]), c, tables,
            K, K2_first, K2, L, passes, split, batches, mode, out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    return out


def _check(g, jac, want):
    G = O.GROUPS[g]
    w = want.size // 2
    assert np.array_equal(jac[: 2 * w], want)
    one = np.array(G.K.encode(G.K.one), dtype=np.uint64)
    assert np.array_equal(jac[2 * w :], one if want.any() else np.zeros_like(one))


@pytest.mark.parametrize("g", GROUPS)
def test_emulated_kernels_match_golden_vectors(g):
    """committed known answers through the emulated kernel pipeline: plain and window-table mode, several widths"""
    pts, s, want = load_golden_msm(g)
    for c, tables in ((4, 0), (9, 0), (16, 0), (5, 1), (11, 1)):
        _check(g, emu_msm(g, pts, s, c, tables=tables, K=8, split=2 if c < 16 else 0), want)


@pytest.mark.parametrize("tables", [0, 1])
def test_emulated_kernels_shapes_bn254_g1(tables):
    """chunk lengths from 1 (every entry its own chunk) to longer than any bucket, short and long carry runs, segment
    lengths, pass counts, with and without the two-part accumulate"""
    g = "bn254_g1"
    pts, s = make_inputs(g, 1500, 321)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    for c in (3, 8, 13):
        for K, K2f, K2, L, passes, split in ((1, 2, 2, 1, 1, 0), (3, 4, 16, 7, 3, 1), (16, 4, 16, 32, 4, 2), (256, 16, 16, 64, 12, 5),
                                             (5000, 3, 5, 1000, 2, 1)):
            _check(g, emu_msm(g, pts, s, c, tables=tables, K=K, K2_first=K2f, K2=K2, L=L, passes=passes, split=split), want)


@pytest.mark.parametrize("kind", ["smallvalues", "redundancy", "one_bucket", "all_equal_points", "all_infinity", "zero_scalars", "empty", "single"])
@pytest.mark.parametrize("tables", [0, 1])
def test_emulated_kernels_skewed_inputs(kind, tables):
    """the distributions of multiexp_test.go:319-334 and harder ones: buckets that span many chunks go through several
    carry levels; all-equal points take the doubling branch; empty and one-element inputs"""
    g = "bn254_g1"
    n = 900
    pts, s = make_inputs(g, n, 55, specials=False)
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
    elif kind == "zero_scalars":
        s[:] = 0
    elif kind == "empty":
        pts, s = pts[:0], s[:0]
    else:
        pts, s = pts[7:8], s[7:8]
    if pts.shape[0]:
        want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    else:
        want = np.zeros(8, dtype=np.uint64)
    for c, K, split in ((6, 4, 0), (12, 16, 2)):
        _check(g, emu_msm(g, pts, s, c, tables=tables, K=K, K2_first=4, K2=4, split=split), want)


@pytest.mark.parametrize("g,n", [("bn254_g2", 300), ("bls12381_g1", 400), ("bls12381_g2", 150), ("bls12377_g1", 300), ("bls12377_g2", 120),
                                 ("secp256k1_g1", 400), ("bw6761_g1", 150), ("bw6761_g2", 120), ("bls24315_g1", 300), ("bls24317_g1", 300),
                                 ("bw6633_g1", 150), ("bw6633_g2", 120)])
def test_emulated_kernels_other_groups(g, n):
    pts, s = make_inputs(g, n, 99)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    _check(g, emu_msm(g, pts, s, 7, tables=0, K=8, split=3), want)
    _check(g, emu_msm(g, pts, s, 10, tables=1, K=32, passes=3, split=1), want)


@pytest.mark.parametrize("tables", [0, 1])
def test_emulated_pipelined_batches(tables):
    """the host calls cut large inputs into batches that share one bucket array (gmsm.cu pipeline_run): later batches go
    to scratch buckets and k_merge_buckets adds them on top; in window-table mode the batch's points are a sub-range of
    every table row"""
    g = "bn254_g1"
    pts, s = make_inputs(g, 1100, 77)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    for c, batches in ((7, 2), (11, 5)):
        _check(g, emu_msm(g, pts, s, c, tables=tables, K=8, batches=batches, split=1), want)


@pytest.mark.parametrize("g", ["bn254_g1", "bn254_g2", "bls12381_g1"])
def test_emulated_fixed_base_kernels(g):
    """k_generate_multiples and k_batch_scalar_mul (next-row N1: BatchScalarMultiplicationG1/G2, g1.go:1039-1118)"""
    G = O.GROUPS[g]
    lib = _lib()
    k = GROUPS.index(g)
    base = G.encode_affine([G.scalar_mul(G.gen, 0xBEEF)])[0]
    w = base.size
    n = 77
    out = np.zeros((n, w), dtype=np.uint64)
    vp = ctypes.c_void_p
    assert getattr(lib, "emu_generate_%d" % k)(base.ctypes.data_as(vp), ctypes.c_uint64(5), ctypes.c_size_t(n), out.ctypes.data_as(vp)) == 0
    for i in (0, 1, 15, 16, 17, 76):
        assert np.array_equal(out[i], cref.scalar_mul(g, base, 5 + i)), i
    s = cref.random_scalars(g, 40, 9)
    specials = [0, 1, 2, G.fr.q - 1, (1 << 200) + 12345]
    s[: len(specials)] = G.encode_scalars(specials)
    ks = G.decode_scalars(s)
    for c in (4, 8):
        got = np.zeros((40, w), dtype=np.uint64)
        assert getattr(lib, "emu_batch_scalar_mul_%d" % k)(base.ctypes.data_as(vp), s.ctypes.data_as(vp), ctypes.c_size_t(40), c,
                                                           got.ctypes.data_as(vp)) == 0
        for i in range(40):
            assert np.array_equal(got[i], cref.scalar_mul(g, base, ks[i])), (c, i)
        assert not got[0].any() and np.array_equal(got[1], base)


def test_emulated_kernels_randomised_shapes():
    """seeded sweep over sizes, widths, chunk / run / segment lengths, pass counts, launch splits, batch counts, both modes
    and input mixes (infinity points, zero / tiny / repeated scalars, duplicated and negated points): whatever the shape,
    the kernel pipeline must reproduce the reference algorithm's result"""
    import random

    g = "bn254_g1"
    G = O.GROUPS[g]
    base_pts, base_s = make_inputs(g, 320, 4242, specials=False)
    rng = random.Random(20260923)
    for it in range(160):
        n = rng.choice([1, 2, 3, 5, 17, 31, 32, 33, 64, 100, 127, 200, 257, 320])
        pts, s = base_pts[:n].copy(), base_s[:n].copy()
        for _ in range(rng.randrange(0, 4)):                       # input mix
            kind = rng.randrange(6)
            i, j = rng.randrange(n), rng.randrange(n)
            if kind == 0:
                pts[i] = 0
            elif kind == 1:
                s[i] = 0
            elif kind == 2:
                s[i] = np.array([rng.randrange(1, 9), 0, 0, 0], dtype=np.uint64)
            elif kind == 3:
                s[min(i, j) : max(i, j) + 1] = s[i]
            elif kind == 4:
                pts[i], s[i] = pts[j], s[j]
            else:
                pts[i] = G.encode_affine([G.aff_neg(G.decode_affine(pts[j : j + 1])[0])])[0]
                s[i] = s[j]
        c = rng.randrange(2, 17)
        opt = dict(tables=rng.randrange(2), K=rng.choice([1, 2, 3, 4, 7, 8, 16, 33, 64, 1000]), K2_first=rng.randrange(2, 9),
                   K2=rng.randrange(2, 17), L=rng.choice([1, 2, 3, 8, 32, 64, 100]), passes=rng.randrange(1, 9),
                   split=rng.randrange(0, 7), batches=rng.choice([1, 1, 2, 3]), order=rng.choice([0, 1, 2, 3, 7, 12345]))
        want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=2)
        jac = emu_msm(g, pts, s, c, **opt)
        assert np.array_equal(jac[:8], want), (it, n, c, opt)


@pytest.mark.parametrize("g,n", [("bn254_g2", 90), ("bls12381_g1", 120)])
def test_emulated_kernels_block_order_independent(g, n):
    """a GPU runs the blocks of a launch in no particular order, which changes the order of the entries inside a bucket
    (atomic scatter) and with it the chain of additions: the affine normal form of the result must not depend on it"""
    pts, s = make_inputs(g, n, 31)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    for order in (0, 1, 2, 5, 99):
        for tables in (0, 1):
            _check(g, emu_msm(g, pts, s, 6, tables=tables, K=4, passes=2, split=1, order=order), want)


@pytest.mark.parametrize("tables", [0, 1])
def test_emulated_real_scan_kernels(tables):
    """K1b: k_scan_block_sums / k_scan_top / k_scan_final (block scans built on warp shuffles and barriers) under the
    cooperative launcher, on histograms shorter and longer than one scan tile (2048 counters)"""
    g = "bn254_g1"
    pts, s = make_inputs(g, 700, 8)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    for c in (3, 9, 13):     # nb_total + 1 = 341 / 7169 (plain) counters ... up to 4097 shared
        _check(g, emu_msm(g, pts, s, c, tables=tables, K=8, mode=1), want)
    _check(g, emu_msm(g, pts, s, 7, tables=tables, K=8, split=2, mode=1 | 4), want)   # everything cooperative


@pytest.mark.parametrize("g,n", [("bn254_g1", 900), ("bn254_g2", 200), ("bls12381_g1", 300)])
def test_emulated_batch_affine_accumulation(g, n):
    """the batch-affine bucket pass (GMSM_AFFINE=1; the GPU restatement of processChunkG1BatchAffine + batchAddG1Affine,
    multiexp_affine.go:24-231, g1.go:1122-1182): tree levels over the bucket-ordered entries, forward / backward passes
    around the hierarchical product scans with ONE inversion per level -- with duplicated points (doubling), P / -P
    (cancellation), infinity points and zero scalars in the input"""
    pts, s = make_inputs(g, n, 2718)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    for c, B, mode in ((4, 8, 2), (10, 3, 2 | 1), (13, 128, 2)):
        _check(g, emu_msm(g, pts, s, c, K=B, mode=mode), want)


@pytest.mark.parametrize("kind", ["one_bucket", "all_equal_points", "all_infinity", "single"])
def test_emulated_batch_affine_skewed(kind):
    g = "bn254_g1"
    n = 600
    pts, s = make_inputs(g, n, 56, specials=False)
    if kind == "one_bucket":
        s[:] = s[0]
    elif kind == "all_equal_points":
        pts[:] = pts[3]
    elif kind == "all_infinity":
        pts[:] = 0
    else:
        pts, s = pts[7:8], s[7:8]
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    for c in (5, 12):
        _check(g, emu_msm(g, pts, s, c, K=8, mode=2 | 1), want)


# ---- next-row N3: the Fr FFT kernels (fft_kernels.cuh) ----
FR_FIELDS = ["bn254_fr", "bls12381_fr", "bls12377_fr"]


def _emu_fft(frname, vals, inverse, decimation, coset, shift=None, bit_reverse_only=False):
    f = O.FIELDS[frname]
    n = len(vals)
    logn = n.bit_length() - 1
    od = O.FFTDomain(frname, n, shift=shift)
    enc = lambda xs: np.array([f.to_limbs(f.to_mont(v)) for v in xs], dtype=np.uint64)
    a = enc(vals)
    consts = enc([od.generator, od.generator_inv, od.cardinality_inv, od.shift, od.shift_inv])
    rc = _lib().emu_fft_run(FR_FIELDS.index(frname), a.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(n), logn, int(inverse), int(decimation),
                            int(coset), consts.ctypes.data_as(ctypes.c_void_p), int(bit_reverse_only))
    assert rc == 0
    return [f.from_mont(O.Field.from_limbs([int(x) for x in r])) for r in a], od


@pytest.mark.parametrize("frname", FR_FIELDS)
@pytest.mark.parametrize("logn", [0, 1, 3, 6, 10, 11, 12])
def test_emulated_fft_kernels(frname, logn):
    """Domain.FFT / FFTInverse (ecc/bn254/fr/fft/fft.go:31-190), both decimations, plain and on the coset: sizes below, at
    and above the shared-memory tile (2^10), where the strided stages (k_fft_dif_stage / k_fft_dit_stage) join the tile
    kernel (barriers: cooperative launcher)"""
    import random

    f = O.FIELDS[frname]
    n = 1 << logn
    rng = random.Random(100 + logn)
    vals = [rng.randrange(f.q) for _ in range(n)]
    cases = [(dec, coset) for dec in (O.DIT, O.DIF) for coset in (False, True)]
    if logn >= 11 and frname != "bn254_fr":
        cases = cases[1:3]                       # keep the big sizes cheap for the other fields
    for dec, coset in cases:
        got, od = _emu_fft(frname, vals, False, dec, coset)
        assert got == od.fft(vals, dec, coset), (dec, coset)
        got, od = _emu_fft(frname, vals, True, dec, coset)
        assert got == od.fft_inverse(vals, dec, coset), (dec, coset)


def test_emulated_fft_custom_shift_and_bit_reverse():
    f = O.FIELDS["bn254_fr"]
    n = 256
    vals = [(7 * i * i + 3) % f.q for i in range(n)]
    got, od = _emu_fft("bn254_fr", vals, False, O.DIF, True, shift=987654321)
    assert got == od.fft(vals, O.DIF, True)
    got, _ = _emu_fft("bn254_fr", vals, False, O.DIF, False, bit_reverse_only=True)      # fft.BitReverse (bitreverse.go:17-42)
    assert got == O.bit_reverse(list(vals))
    # DIF then DIT-inverse without any reordering is the identity (the composition gnark's provers use)
    fwd, _ = _emu_fft("bn254_fr", vals, False, O.DIF, False)
    back, _ = _emu_fft("bn254_fr", fwd, True, O.DIT, False)
    assert back == vals


@pytest.mark.parametrize("g,tables", [("bn254_g1", 0), ("bn254_g1", 1), ("bn254_g2", 0)])
def test_emulated_sharded_window_sums_and_finalize(g, tables):
    """the multi-GPU decomposition (gnark-crypto_b200/dist.py): contiguous shards, W window partials per rank
    (one in window-table mode), gathered rank-major, k_finalize sums them per window over the ranks before the Horner --
    the reference's analogue is the recursive split joined by AddAssign (multiexp.go:128-140)"""
    n = 600 if g == "bn254_g1" else 150
    pts, s = make_inputs(g, n, 4)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    lib = _lib()
    k = GROUPS.index(g)
    w = pts.shape[1] // 2
    vp = ctypes.c_void_p
    for c in (6, 11):
        W = 1 if tables else O.compute_nb_chunks(O.GROUPS[g].fr.bits, c)
        cuts = [0, n // 7, n // 2, n - 1, n]              # four ranks, one of them with a single point
        parts = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            part = np.zeros(W * 4 * w, dtype=np.uint64)
            P, S = np.ascontiguousarray(pts[a:b]), np.ascontiguousarray(s[a:b])
            assert getattr(lib, "emu_window_sums_%d" % k)(P.ctypes.data_as(vp), S.ctypes.data_as(vp), ctypes.c_size_t(b - a), c, tables, 8,
                                                          part.ctypes.data_as(vp)) == 0
            parts.append(part)
        gathered = np.concatenate(parts)
        out = np.zeros(3 * w, dtype=np.uint64)
        assert getattr(lib, "emu_finalize_%d" % k)(gathered.ctypes.data_as(vp), len(parts), c, tables, out.ctypes.data_as(vp)) == 0
        _check(g, out, want)


@pytest.mark.parametrize("g,n", [("bn254_g1", 700), ("bls12381_g1", 260), ("bls12377_g1", 150), ("bn254_g2", 180), ("bls12381_g2", 100), ("bls12377_g2", 80),
                                 # the groups whose contexts take the lane-parallel tail BY DEFAULT (gmsm.cu: bw6-761, bw6-633), and
                                 # the full-width-modulus field through the same lane-parallel formulas
                                 ("bw6761_g1", 70), ("bw6633_g2", 80), ("secp256k1_g1", 200)])
def test_emulated_quad_tail(g, n):
    """the lane-parallel (quad) form of the tail kernels (csrc/quad.cuh: one point operation per four lanes, products of a
    formula step spread over the lanes and broadcast with masked shuffles) -- carry levels, segment reduction with its
    double-and-add weights, group sums -- and k_finalize's quad Horner, against the oracle: cross-test ingredients
    (infinity points, duplicates -> doubling branch, P / -P -> cancellation), plain and window-table mode, odd segment and
    run lengths, several block orders"""
    pts, s = make_inputs(g, n, 21)
    want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=4)
    for c, K, K2f, K2, L, tables, order in ((5, 4, 2, 3, 3, 0, 0), (8, 16, 4, 16, 32, 0, 1), (6, 7, 3, 5, 5, 1, 3)):
        _check(g, emu_msm(g, pts, s, c, tables=tables, K=K, K2_first=K2f, K2=K2, L=L, passes=2, order=order, mode=8), want)


def test_emulated_quad_tail_skewed_and_all_equal():
    """quads whose chains take different branches inside one warp: all scalars equal (one bucket per window spans every chunk:
    the carry join does all the work), all points equal (doubling branch everywhere), a single non-zero bucket"""
    g = "bn254_g1"
    G = O.GROUPS[g]
    pts, s = make_inputs(g, 400, 5, specials=False)
    s2 = s.copy()
    s2[:] = s[0]
    for (pp, ss) in ((pts, s2), (np.repeat(pts[:1], 400, axis=0), s), (np.repeat(pts[:1], 400, axis=0), s2)):
        want, _, _, _ = cref.msm(g, pp, ss, c=0, nthreads=4)
        _check(g, emu_msm(g, pp, ss, 7, K=4, K2_first=2, K2=2, L=4, mode=8), want)
        _check(g, emu_msm(g, pp, ss, 7, K=4, K2_first=2, K2=2, L=4, mode=0), want)


def test_emulated_k1_rank_numbering_paths():
    """the two modes of the counting sort (kernels.cuh K1 / K1c): plain -- RED histogram, positions from a returning atomicSub in
    the scatter -- and rank mode -- K1's warp-aggregated returning atomicAdd numbers the entries, the scatter has no atomics.
    Mode bit 4 forces rank mode, bit 5 plain mode, neither lets the sampling pass decide: it must flag runs of equal scalars
    ("redundancy", multiexp_test.go:327-334) and a global hot value ("smallvalues", :316-325: every 5th scalar equal).  The
    emulated engine checks the ranks of every bucket are a permutation of 0 .. count-1, or that the plain scatter consumed
    every counter exactly (rc 10 otherwise)."""
    g = "bn254_g1"
    pts, s = make_inputs(g, 3000, 13, specials=False)
    small = s.copy()
    small[::5] = np.array([1, 0, 0, 0], dtype=np.uint64)
    red = s.copy()
    for i in range(0, 3000, 100):
        red[i : i + 100] = red[i]
    for scal in (s, small, red):
        want, _, _, _ = cref.msm(g, pts, scal, c=0, nthreads=4)
        for mode in (0, 16, 32):
            _check(g, emu_msm(g, pts, scal, 9, K=8, mode=mode), want)
            _check(g, emu_msm(g, pts, scal, 6, tables=1, K=8, passes=2, mode=mode), want)
