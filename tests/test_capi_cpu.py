"""CPU-side checks of the product boundary: libgmsm.so loads, exports every symbol include/gmsm.h
declares, reports sizes, and REFUSES to compute without a GPU (no CPU fallback).  No compute calls."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pkg():
    import gnark_crypto_b200 as pkg

    return pkg


def test_header_symbols_exported():
    pkg = _pkg()
    L = pkg._native.lib()
    hdr = open(os.path.join(ROOT, "include", "gmsm.h")).read()
    declared = set(re.findall(r"\b(gmsm_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"gmsm_curve_t"}
    assert declared == set(pkg._native.SYMBOLS), declared ^ set(pkg._native.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s


def test_sizes_match_go_layout():
    pkg = _pkg()
    L = pkg._native.lib()
    # affine 64 / 128 / 96 / 192 B, scalars 32 B, Jacobian 96 / 192 / 144 / 288 B
    assert [L.gmsm_affine_bytes(c) for c in range(13)] == [64, 128, 96, 192, 96, 192, 64, 192, 192, 80, 80, 160, 160]
    assert [L.gmsm_scalar_bytes(c) for c in range(13)] == [32] * 7 + [48, 48, 32, 32, 40, 40]     # bw6-761: fr.Limbs = 6, bw6-633: 5
    assert L.gmsm_affine_bytes(13) == 0 and L.gmsm_scalar_bytes(13) == 0
    assert [L.gmsm_jac_bytes(c) for c in range(4)] == [96, 192, 144, 288]
    assert [L.gmsm_xyzz_bytes(c) for c in range(4)] == [128, 256, 192, 384]
    assert b"sm_100a" in L.gmsm_version()


def test_reference_error_strings():
    pkg = _pkg()
    pts = np.zeros((3, 8), dtype=np.uint64)
    sc = np.zeros((2, 4), dtype=np.uint64)
    with pytest.raises(pkg.MultiExpError, match=r"len\(points\) != len\(scalars\)"):
        pkg.G1Jac().MultiExp(pts, sc, pkg.MultiExpConfig())
    with pytest.raises(pkg.MultiExpError, match=r"invalid config: config.NbTasks > 1024"):
        pkg.G1Jac().MultiExp(pts, np.zeros((3, 4), dtype=np.uint64), pkg.MultiExpConfig(NbTasks=1025))


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device refusal is exercised on the CPU-only container")
    pkg = _pkg()
    pts = np.zeros((3, 8), dtype=np.uint64)
    sc = np.ones((3, 4), dtype=np.uint64)
    with pytest.raises(pkg.MultiExpError, match="no CUDA device|no CPU fallback|CUDA"):
        pkg.G1Jac().MultiExp(pts, sc, pkg.MultiExpConfig())
    with pytest.raises(pkg.MultiExpError):
        pkg.G1Jac().MultiExp(pts[:0], sc[:0], pkg.MultiExpConfig())  # even n = 0 needs the device


def test_product_does_not_import_oracle():
    # the product path must never route through oracle/
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gnark-crypto_b200")):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\"[^\"]*oracle/|libmsmref", src, re.M), f


def test_header_is_plain_c99_and_links(tmp_path):
    """include/gmsm.h is what a cgo shim compiles: it must be valid C (no C++-isms) and a C program must
    link against libgmsm.so and get the reference layout sizes"""
    import subprocess

    src = tmp_path / "abi.c"
    src.write_text('#include "gmsm.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%zu %zu %zu %zu %s\\n", gmsm_affine_bytes(GMSM_BN254_G1), gmsm_affine_bytes(GMSM_BLS12381_G2),\n'
                   '  gmsm_jac_bytes(GMSM_BN254_G2), gmsm_scalar_bytes(GMSM_BLS12377_G1), gmsm_version()); return 0; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "gnark-crypto_b200")
    cuda = "/usr/local/cuda/lib64"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lgmsm", "-L", cuda, "-lcudart", "-Wl,-rpath," + libdir, "-Wl,-rpath," + cuda], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out[:4] == ["64", "192", "192", "32"] and out[4].startswith("gmsm-b200")


def test_window_model_choices_are_pinned():
    """gmsm_choose_window_bits (the fitted width model of gmsm.cu) is pure host arithmetic: pin its choices at
    the configurations whose sweeps were measured on a B200 (profiles/r02_c_sweep_call3.txt, r02_window_model_validation.md) --
    every one of them is the measured optimum -- and its monotone behaviour in n"""
    import importlib

    native = importlib.import_module("gnark-crypto_b200._native")
    L = native.lib()
    BN_G1, BN_G2, BLS_G1, BLS_G2, B377_G1 = 0, 1, 2, 3, 4
    want = {(BN_G1, 16): 15, (BN_G1, 18): 15, (BN_G1, 20): 17, (BN_G1, 22): 17, (BN_G1, 23): 17, (BN_G1, 24): 17, (BN_G1, 26): 20,
            (BLS_G1, 20): 16, (BLS_G1, 22): 17, (BLS_G1, 23): 17, (BLS_G1, 24): 20, (BN_G1, 25): 20, (BN_G2, 24): 20, (B377_G1, 24): 20, (BN_G2, 20): 17, (BN_G2, 22): 17, (BLS_G2, 20): 17, (B377_G1, 22): 17}
    # N4 curves (profiles/r02_n4_new_curves_call11.txt, _call12.txt): secp256k1's 256-bit scalars make 16 | 256 and 20 (13 windows,
    # a full-width last one) the good widths -- 15, 17, 18, 19 leave a last window of 1..9 bits whose few buckets serialise K1
    SECP, BW6_G1, BW6_G2 = 6, 7, 8
    want.update({(SECP, 16): 14, (SECP, 20): 16, (SECP, 22): 16, (SECP, 24): 20, (BW6_G1, 18): 14, (BW6_G1, 20): 16, (BW6_G1, 22): 18,
                 (BW6_G1, 24): 19, (BW6_G2, 20): 16})
    # bls24-315 / bls24-317 G1, bw6-633 (profiles/r02_n4_more_curves_call13.txt, r02_n4_model_checks_call14.txt)
    B315, B317, BW633_G1, BW633_G2 = 9, 10, 11, 12
    want.update({(B315, 20): 17, (B315, 24): 20, (B317, 22): 17, (BW633_G1, 18): 15, (BW633_G1, 22): 18, (BW633_G2, 20): 16})
    # (c = 20 entries: profiles/r02_c20_checks_call15.txt -- 13 windows of 20 bits end on a full-width last window; bls12-381 G1's
    # earlier optimum c = 19 has a 9-bit last window whose 256 buckets serialise K1: 105.0 ms against 96.8)
    for (curve, logn), c in want.items():
        assert L.gmsm_choose_window_bits(curve, 1 << logn) == c, (curve, logn)
    for curve in (BN_G1, BN_G2, BLS_G1, BLS_G2, BW6_G1):
        cs = [L.gmsm_choose_window_bits(curve, 1 << k) for k in range(4, 28)]
        assert all(2 <= c <= 24 for c in cs) and all(b >= a for a, b in zip(cs, cs[1:])), cs     # wider windows for larger n
