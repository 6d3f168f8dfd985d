"""ctypes binding of the CPU oracle's C port (oracle/msm_ref.c) -- TEST INFRASTRUCTURE ONLY.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libmsmref.so")

CURVES = {"bn254_g1": 0, "bn254_g2": 1, "bls12381_g1": 2, "bls12381_g2": 3, "bls12377_g1": 4, "bls12377_g2": 5,
          "secp256k1_g1": 6, "bw6761_g1": 7, "bw6761_g2": 8, "bls24315_g1": 9, "bls24317_g1": 10, "bw6633_g1": 11, "bw6633_g2": 12}
AFF_WORDS = {0: 8, 1: 16, 2: 12, 3: 24, 4: 12, 5: 24, 6: 8, 7: 24, 8: 24, 9: 10, 10: 10, 11: 20, 12: 20}  # u64 words per affine point
SCALAR_WORDS = {0: 4, 1: 4, 2: 4, 3: 4, 4: 4, 5: 4, 6: 4, 7: 6, 8: 6, 9: 4, 10: 4, 11: 5, 12: 5}      # u64 words per scalar (fr.Limbs)
SCALAR_BITS = {0: 254, 1: 254, 2: 255, 3: 255, 4: 253, 5: 253, 6: 256, 7: 377, 8: 377, 9: 253, 10: 255, 11: 315, 12: 315}   # fr.Bits

_lib = None


def build():
    subprocess.run(["make", "-C", HERE], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        _lib.ref_msm.argtypes = [ctypes.c_int, u64p, u64p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, u64p, u64p,
                                 ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        _lib.ref_partition_scalars.argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
        _lib.ref_generate_multiples.argtypes = [ctypes.c_int, u64p, ctypes.c_uint64, ctypes.c_size_t, u64p, ctypes.c_int]
        _lib.ref_scalar_mul.argtypes = [ctypes.c_int, u64p, u64p, u64p]
        _lib.ref_dot_index.argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, ctypes.c_uint64, u64p]
        _lib.ref_random_scalars.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_uint64, u64p]
        _lib.ref_field_op.argtypes = [ctypes.c_int, ctypes.c_int, u64p, u64p, u64p, ctypes.c_size_t]
    return _lib


def set_batch_affine(on: bool):
    """processor selection of the port: True (default) = the reference's getChunkProcessorG1 (batch-affine buckets
    for 10 <= c <= 16), False = extended-Jacobian buckets everywhere"""
    lib().ref_set_batch_affine(1 if on else 0)


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def _cid(curve):
    return CURVES[curve] if isinstance(curve, str) else int(curve)


def msm(curve, points: np.ndarray, scalars: np.ndarray, c: int = 0, nthreads: int = 1, nb_tasks: int = 0):
    """-> (affine u64 array, jac u64 array, used_c, leaves)"""
    cid = _cid(curve)
    w = AFF_WORDS[cid]
    points = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, w)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, SCALAR_WORDS[cid])
    n = points.shape[0]
    assert scalars.shape[0] == n
    aff = np.zeros(w, dtype=np.uint64)
    jac = np.zeros(w // 2 * 3, dtype=np.uint64)
    uc, ul = ctypes.c_int(0), ctypes.c_int(0)
    rc = lib().ref_msm(cid, _p(points), _p(scalars), n, c, nthreads, nb_tasks, _p(aff), _p(jac), ctypes.byref(uc), ctypes.byref(ul))
    if rc != 0:
        raise RuntimeError("ref_msm rc=%d" % rc)
    return aff, jac, uc.value, ul.value


def partition_scalars(curve, scalars: np.ndarray, c: int) -> np.ndarray:
    cid = _cid(curve)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, SCALAR_WORDS[cid])
    n = scalars.shape[0]
    bits = SCALAR_BITS[cid]
    W = (bits + c - 1) // c
    out = np.zeros((W, n), dtype=np.uint32)
    rc = lib().ref_partition_scalars(cid, _p(scalars), n, c, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    if rc != 0:
        raise RuntimeError("ref_partition_scalars rc=%d" % rc)
    return out


def generate_multiples(curve, base_aff: np.ndarray, start: int, n: int, nthreads: int = 1) -> np.ndarray:
    cid = _cid(curve)
    w = AFF_WORDS[cid]
    base_aff = np.ascontiguousarray(base_aff, dtype=np.uint64).reshape(w)
    out = np.zeros((n, w), dtype=np.uint64)
    rc = lib().ref_generate_multiples(cid, _p(base_aff), start, n, _p(out), nthreads)
    if rc != 0:
        raise RuntimeError("ref_generate_multiples rc=%d" % rc)
    return out


def scalar_mul(curve, base_aff: np.ndarray, k: int) -> np.ndarray:
    cid = _cid(curve)
    w = AFF_WORDS[cid]
    base_aff = np.ascontiguousarray(base_aff, dtype=np.uint64).reshape(w)
    kl = np.array([(k >> (64 * i)) & (2**64 - 1) for i in range(SCALAR_WORDS[cid])], dtype=np.uint64)
    out = np.zeros(w, dtype=np.uint64)
    rc = lib().ref_scalar_mul(cid, _p(base_aff), _p(kl), _p(out))
    if rc != 0:
        raise RuntimeError("ref_scalar_mul rc=%d" % rc)
    return out


def dot_index(curve, scalars: np.ndarray, start: int) -> int:
    """sum_i (start+i) * s_i mod r as a Python int"""
    cid = _cid(curve)
    sw = SCALAR_WORDS[cid]
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, sw)
    out = np.zeros(sw, dtype=np.uint64)
    rc = lib().ref_dot_index(cid, _p(scalars), scalars.shape[0], start, _p(out))
    if rc != 0:
        raise RuntimeError("ref_dot_index rc=%d" % rc)
    return sum(int(out[i]) << (64 * i) for i in range(sw))


def random_scalars(curve, n: int, seed: int) -> np.ndarray:
    cid = _cid(curve)
    out = np.zeros((n, SCALAR_WORDS[cid]), dtype=np.uint64)
    rc = lib().ref_random_scalars(cid, n, seed, _p(out))
    if rc != 0:
        raise RuntimeError("ref_random_scalars rc=%d" % rc)
    return out


def field_op(field: int, op: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.zeros_like(a)
    L = {2: 6, 4: 6, 8: 12, 9: 6, 10: 5, 12: 5, 14: 10, 15: 5}.get(field, 4)
    rc = lib().ref_field_op(field, op, _p(a), _p(b), _p(out), a.size // L)
    if rc != 0:
        raise RuntimeError("ref_field_op rc=%d" % rc)
    return out
