"""Next-row N2: the caller on top of MultiExp and its raw SRS format.

  * kzg.Commit(p, pk, nbTasks...)            ecc/bn254/kzg/kzg.go:159-176  -> MultiExp over pk.G1[:len(p)]
  * kzg.NewSRS's G1 side (powers of alpha)   ecc/bn254/kzg/kzg.go:100-135  -> BatchScalarMultiplicationG1
  * unsafe.WriteSlice / ReadSlice            utils/unsafe/dump_slice.go:16-76 (uint64 LE length + raw
    []G1Affine memory) and the 0xdeadbeef marker that precedes it in SRS.WriteDump,
    ecc/bn254/kzg/marshal.go:70-115 -- the raw image IS the layout the device wants, so a dump streams
    straight into resident bases.
Only the G1 proving-key side is handled (the verifying key / pairing side is out of scope)."""
from __future__ import annotations

import io
import struct

import numpy as np

from .multiexp import CURVES, BatchScalarMultiplication, MultiExpConfig, MultiExpError, ResidentBases, _words

MARKER = 0xDEADBEEF  # utils/unsafe/dump_slice.go:78


class ErrInvalidPolynomialSize(MultiExpError):
    """kzg.ErrInvalidPolynomialSize (kzg.go:24)"""


def write_slice(w, points: np.ndarray) -> None:
    """unsafe.WriteSlice: uint64 little-endian length, then the raw element memory"""
    points = np.ascontiguousarray(points, dtype=np.uint64)
    w.write(struct.pack("<Q", points.shape[0]))
    if points.shape[0]:
        w.write(points.tobytes())


def read_slice(r, words_per_element: int, max_elements: int = 0) -> np.ndarray:
    """unsafe.ReadSlice (dump_slice.go:36-76): reads min(length, max_elements) elements, skips the rest"""
    hdr = r.read(8)
    if len(hdr) != 8:
        raise EOFError("unexpected EOF")
    (length,) = struct.unpack("<Q", hdr)
    limit = length
    if max_elements > 0 and length > max_elements:
        limit = max_elements
    size = 8 * words_per_element
    data = r.read(size * limit)
    if len(data) != size * limit:
        raise EOFError("unexpected EOF")
    if length > limit:
        r.seek((length - limit) * size, io.SEEK_CUR)
    return np.frombuffer(data, dtype=np.uint64).reshape(limit, words_per_element).copy()


def write_marker(w) -> None:
    w.write(struct.pack("<Q", MARKER))


def read_marker(r) -> None:
    b = r.read(8)
    if len(b) != 8 or struct.unpack("<Q", b)[0] != MARKER:
        raise ValueError("marker mismatch")  # dump_slice.go:92-99


class ProvingKey:
    """kzg.ProvingKey{G1 []G1Affine} (kzg.go:38-41) with the bases resident in HBM"""

    def __init__(self, curve: str, g1_points: np.ndarray, device: int = 0):
        self.curve = curve + "_g1" if not curve.endswith("_g1") else curve
        self.words = 2 * _words(CURVES[self.curve])
        self.G1 = np.ascontiguousarray(g1_points, dtype=np.uint64).reshape(-1, self.words)
        self._bases = ResidentBases(self.curve, self.G1, device)

    @classmethod
    def from_dump(cls, curve: str, r, max_pk_points: int = 0, device: int = 0):
        """the marker + slice tail of SRS.ReadDump (marshal.go:98-115); `r` positioned at the marker"""
        read_marker(r)
        cname = curve + "_g1" if not curve.endswith("_g1") else curve
        pts = read_slice(r, 2 * _words(CURVES[cname]), max_pk_points)
        return cls(curve, pts, device)

    def close(self):
        self._bases.close()


def new_srs_g1(curve: str, size: int, alpha: int, generator: np.ndarray, r_modulus: int, encode_scalars) -> np.ndarray:
    """G1 side of kzg.NewSRS (kzg.go:100-135): [1, alpha, alpha^2, ...] * G via BatchScalarMultiplicationG1.
    `encode_scalars` turns Python ints into Montgomery fr limbs (the caller's fr.Element constructor)."""
    alphas, a = [], 1
    for _ in range(size):
        alphas.append(a)
        a = a * alpha % r_modulus
    cname = curve + "_g1" if not curve.endswith("_g1") else curve
    return BatchScalarMultiplication(cname, generator, encode_scalars(alphas))


def Commit(p: np.ndarray, pk: ProvingKey, *nbTasks: int) -> np.ndarray:
    """kzg.Commit (kzg.go:159-176): Digest = MultiExp(pk.G1[:len(p)], p) as an affine point"""
    p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 4)
    if p.shape[0] == 0 or p.shape[0] > pk.G1.shape[0]:
        raise ErrInvalidPolynomialSize("invalid polynomial size (larger than SRS or == 0)")
    cfg = MultiExpConfig(NbTasks=nbTasks[0] if nbTasks else 0)
    jac = pk._bases.MultiExp(p, cfg)
    w = pk.words
    return jac[:w].copy() if jac[w:].any() else np.zeros(w, dtype=np.uint64)
