"""CPU build (g++, portable arithmetic path) of the product's field / curve headers against the
oracle: checks the formulas and the window plan without a GPU.  The device (PTX) arithmetic path is
checked by the same vectors in tests/test_gpu_ops.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from tests import opcases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gnark-crypto_b200", "csrc")
OUT = os.path.join(ROOT, "gnark-crypto_b200", "build", "libgmsm_hostcheck.so")


@pytest.fixture(scope="module")
def hc():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", OUT, os.path.join(CSRC, "hostcheck.cpp")], check=True)
    return ctypes.CDLL(OUT)


def _runner(hc, g):
    cid = list(O.GROUPS).index(g)
    assert list(O.GROUPS) == ["bn254_g1", "bn254_g2", "bls12381_g1", "bls12381_g2", "bls12377_g1", "bls12377_g2"]

    def run(op, a, b, out_words):
        a = np.ascontiguousarray(a, dtype=np.uint32)
        n = a.shape[0]
        b = np.ascontiguousarray(b, dtype=np.uint32) if b is not None else np.zeros((n, 1), dtype=np.uint32)
        out = np.zeros((n, out_words), dtype=np.uint32)
        rc = hc.hostcheck_op(cid, op, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                             out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n))
        assert rc == 0
        return out

    return run


@pytest.mark.parametrize("g", list(O.GROUPS))
def test_field_ops_host(hc, g):
    opcases.check_field_ops(O.GROUPS[g], _runner(hc, g))
    opcases.check_fr_from_mont(O.GROUPS[g], _runner(hc, g))


@pytest.mark.parametrize("g", list(O.GROUPS))
def test_point_ops_host(hc, g):
    opcases.check_point_ops(O.GROUPS[g], _runner(hc, g))


def test_window_plan(hc):
    buf = (ctypes.c_int * 6)()
    for bits in (253, 254, 255):
        for c in range(2, 25):
            hc.hostcheck_plan(bits, c, buf)
            W = O.compute_nb_chunks(bits, c)
            lc = O.last_c(bits, c)
            assert list(buf) == [c, W, lc, 1 << (c - 1), 1 << (lc - 1), (W - 1) * (1 << (c - 1)) + (1 << (lc - 1))]
