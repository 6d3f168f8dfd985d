"""Next-row N2: the caller on top of MultiExp and its raw SRS format.

  * kzg.Commit(p, pk, nbTasks...)            ecc/bn254/kzg/kzg.go:159-176  -> MultiExp over pk.G1[:len(p)]
  * kzg.NewSRS's G1 side (powers of alpha)   ecc/bn254/kzg/kzg.go:100-135  -> BatchScalarMultiplicationG1
  * unsafe.WriteSlice / ReadSlice            utils/unsafe/dump_slice.go:16-76 (uint64 LE length + raw
    []G1Affine memory) and the 0xdeadbeef marker that precedes it in SRS.WriteDump,
    ecc/bn254/kzg/marshal.go:70-115 -- the raw image IS the layout the device wants, so a dump streams
    straight into resident bases.
  * kzg.Open(p, point, pk)                   ecc/bn254/kzg/kzg.go:180-204  -> eval + dividePolyByXminusA on the host
    (Fr Horner loops, as in the reference) and one MultiExp for the quotient commitment
Only the G1 proving-key side is handled (the verifying key / pairing side is out of scope)."""
from __future__ import annotations

import io
import struct
from dataclasses import dataclass

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

    def __init__(self, curve: str, g1_points: np.ndarray, device: int = 0, window_tables: bool = False):
        self.curve = curve + "_g1" if not curve.endswith("_g1") else curve
        self.words = 2 * _words(CURVES[self.curve])
        self.G1 = np.ascontiguousarray(g1_points, dtype=np.uint64).reshape(-1, self.words)
        self._bases = ResidentBases(self.curve, self.G1, device)
        if window_tables:          # the SRS is static: trade W x the device memory for ~20 % faster commitments
            self._bases.Precompute()

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


# scalar-field moduli r (fr/element.go:44-49 `q` of ecc/bn254/fr, ecc/bls12-381/fr, ecc/bls12-377/fr); fr.Element holds
# v * 2^256 mod r (Montgomery form, 4 little-endian u64 limbs)
FR_MODULUS = {
    "bn254": 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
    "bls12381": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    "bls12377": 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
}


def _fr_decode(limbs: np.ndarray, r: int) -> list:
    """Montgomery limbs -> regular integers"""
    rinv = pow(1 << 256, -1, r)
    a = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4)
    return [(int(x[0]) | int(x[1]) << 64 | int(x[2]) << 128 | int(x[3]) << 192) * rinv % r for x in a]


def _fr_encode(vals, r: int) -> np.ndarray:
    out = np.empty((len(vals), 4), dtype=np.uint64)
    m64 = (1 << 64) - 1
    for i, v in enumerate(vals):
        m = (v << 256) % r
        out[i] = [m & m64, (m >> 64) & m64, (m >> 128) & m64, m >> 192]
    return out


def _eval(p: list, point: int, r: int) -> int:
    """eval (kzg.go:55-63): Horner from the top coefficient"""
    res = p[-1]
    for i in range(len(p) - 2, -1, -1):
        res = (res * point + p[i]) % r
    return res


def _divide_by_x_minus_a(f: list, fa: int, a: int, r: int) -> list:
    """dividePolyByXminusA (kzg.go:567-584): (f - f(a)) / (X - a) by synthetic division, result of degree deg(f) - 1"""
    f = list(f)
    f[0] = (f[0] - fa) % r
    for i in range(len(f) - 2, -1, -1):
        f[i] = (f[i] + f[i + 1] * a) % r
    return f[1:]


@dataclass
class OpeningProof:
    """kzg.OpeningProof{H G1Affine, ClaimedValue fr.Element} (kzg.go:43-51), both in Go memory layout"""

    H: np.ndarray
    ClaimedValue: np.ndarray


def Open(p: np.ndarray, point: np.ndarray, pk: ProvingKey) -> OpeningProof:
    """kzg.Open (kzg.go:180-204): ClaimedValue = p(point); H = Commit((p - p(point)) / (X - point))."""
    p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 4)
    if p.shape[0] == 0 or p.shape[0] > pk.G1.shape[0]:
        raise ErrInvalidPolynomialSize("invalid polynomial size (larger than SRS or == 0)")
    r = FR_MODULUS[pk.curve.split("_")[0]]
    coeffs = _fr_decode(p, r)
    a = _fr_decode(point, r)[0]
    fa = _eval(coeffs, a, r)
    h = _divide_by_x_minus_a(coeffs, fa, a, r)
    w = pk.words
    # Commit(h, pk) errors on an empty h in the reference (kzg.go:160-162): a constant polynomial cannot be opened
    H = Commit(_fr_encode(h, r), pk) if h else None
    if H is None:
        raise ErrInvalidPolynomialSize("invalid polynomial size (larger than SRS or == 0)")
    return OpeningProof(H=H.reshape(w), ClaimedValue=_fr_encode([fa], r)[0])


def Commit(p: np.ndarray, pk: ProvingKey, *nbTasks: int) -> np.ndarray:
    """kzg.Commit (kzg.go:159-176): Digest = MultiExp(pk.G1[:len(p)], p) as an affine point"""
    p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 4)
    if p.shape[0] == 0 or p.shape[0] > pk.G1.shape[0]:
        raise ErrInvalidPolynomialSize("invalid polynomial size (larger than SRS or == 0)")
    cfg = MultiExpConfig(NbTasks=nbTasks[0] if nbTasks else 0)
    jac = pk._bases.MultiExp(p, cfg)
    w = pk.words
    return jac[:w].copy() if jac[w:].any() else np.zeros(w, dtype=np.uint64)
