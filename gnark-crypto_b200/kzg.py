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

    @classmethod
    def from_bytes(cls, curve: str, data: bytes, n: int, raw: bool = False, check_on_curve: bool = True, device: int = 0):
        """n G1 points in the standard encoding (Encoder.Encode of a []G1Affine without its length prefix: Bytes() each, or
        RawBytes() each with RawEncoding, marshal.go:560-640) decoded on the device into resident bases"""
        return cls(curve, decode_g1_points(curve, data, n, raw, check_on_curve), device)

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


# ----------------------------------------------------------------------------------------------------------------
# point (de)serialisation on the host -- G1Affine.RawBytes / Bytes / SetBytes (ecc/bn254/marshal.go:801-950,
# ecc/bls12-381/marshal.go:830-1000).  Used for the few points a Fiat-Shamir transcript binds (G1Affine.Marshal is
# RawBytes, marshal.go:779-782); bulk SRS decoding runs on the device (gmsm_g1_decode, csrc/decode.cu).
# ----------------------------------------------------------------------------------------------------------------
FP_MODULUS = {
    "bn254": 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
    "bls12381": 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    "bls12377": 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001,
}
CURVE_B = {"bn254": 3, "bls12381": 4, "bls12377": 1}     # y^2 = x^3 + b (bn254.go:12, bls12-381.go:9, bls12-377.go)
# flag bits of the most significant byte (marshal.go:25-31 bn254: two bits; bls12-381 / bls12-377: three bits)
_FLAGS = {
    "bn254": dict(mask=0b11 << 6, unc=0b00 << 6, unc_inf=None, small=0b10 << 6, large=0b11 << 6, inf=0b01 << 6),
    "bls12381": dict(mask=0b111 << 5, unc=0b000 << 5, unc_inf=0b010 << 5, small=0b100 << 5, large=0b101 << 5, inf=0b110 << 5),
    "bls12377": dict(mask=0b111 << 5, unc=0b000 << 5, unc_inf=0b010 << 5, small=0b100 << 5, large=0b101 << 5, inf=0b110 << 5),
}


def _fp_words(curve: str) -> int:
    return 4 if curve == "bn254" else 6


def _fp_decode(limbs, curve: str) -> int:
    L = _fp_words(curve)
    p = FP_MODULUS[curve]
    v = sum(int(limbs[i]) << (64 * i) for i in range(L))
    return v * pow(1 << (64 * L), -1, p) % p


def _fp_encode(v: int, curve: str) -> np.ndarray:
    L = _fp_words(curve)
    p = FP_MODULUS[curve]
    m = (v << (64 * L)) % p
    return np.array([(m >> (64 * i)) & (2**64 - 1) for i in range(L)], dtype=np.uint64)


def g1_raw_bytes(point: np.ndarray, curve: str) -> bytes:
    """G1Affine.RawBytes (marshal.go:826-846): big-endian X || Y, canonical; infinity = flag + zeroes"""
    L = _fp_words(curve)
    point = np.ascontiguousarray(point, dtype=np.uint64).reshape(2 * L)
    nb = 8 * L
    if not point.any():
        f = _FLAGS[curve]
        out = bytearray(2 * nb)
        out[0] = f["unc"] if f["unc_inf"] is None else f["unc_inf"]
        return bytes(out)
    x, y = _fp_decode(point[:L], curve), _fp_decode(point[L:], curve)
    return x.to_bytes(nb, "big") + y.to_bytes(nb, "big")      # mUncompressed = 0: no bits to set


def g1_bytes(point: np.ndarray, curve: str) -> bytes:
    """G1Affine.Bytes (marshal.go:801-823): compressed -- big-endian X with the flag bits in the top byte"""
    L = _fp_words(curve)
    point = np.ascontiguousarray(point, dtype=np.uint64).reshape(2 * L)
    nb = 8 * L
    f = _FLAGS[curve]
    if not point.any():
        out = bytearray(nb)
        out[0] = f["inf"]
        return bytes(out)
    p = FP_MODULUS[curve]
    x, y = _fp_decode(point[:L], curve), _fp_decode(point[L:], curve)
    out = bytearray(x.to_bytes(nb, "big"))
    out[0] |= f["large"] if y > (p - 1) // 2 else f["small"]        # LexicographicallyLargest, fp/element.go:282-296
    return bytes(out)


def g1_set_bytes(buf: bytes, curve: str):
    """G1Affine.SetBytes without the subgroup check (marshal.go:858-950) for ONE point on the host -> (point limbs, consumed).
    Raises ValueError with the reference's messages on invalid encodings."""
    L = _fp_words(curve)
    nb = 8 * L
    f = _FLAGS[curve]
    p = FP_MODULUS[curve]
    if len(buf) < nb:
        raise EOFError("short buffer")
    m = buf[0] & f["mask"]
    if m == f["inf"]:
        if (buf[0] & ~f["mask"] & 0xFF) or any(buf[1:nb]):
            raise ValueError("invalid infinity point encoding")
        return np.zeros(2 * L, dtype=np.uint64), nb
    if f["unc_inf"] is not None and m == f["unc_inf"]:
        if len(buf) < 2 * nb:
            raise EOFError("short buffer")
        if (buf[0] & ~f["mask"] & 0xFF) or any(buf[1:2 * nb]):
            raise ValueError("invalid infinity point encoding")
        return np.zeros(2 * L, dtype=np.uint64), 2 * nb
    xb = bytearray(buf[:nb])
    xb[0] &= ~f["mask"] & 0xFF
    x = int.from_bytes(xb, "big")
    if x >= p:
        raise ValueError("invalid fp.Element encoding")
    if m == f["unc"]:
        if len(buf) < 2 * nb:
            raise EOFError("short buffer")
        y = int.from_bytes(buf[nb:2 * nb], "big")
        if y >= p:
            raise ValueError("invalid fp.Element encoding")
        return np.concatenate([_fp_encode(x, curve), _fp_encode(y, curve)]), 2 * nb
    if m not in (f["small"], f["large"]):
        raise ValueError("invalid point encoding")
    y2 = (x * x * x + CURVE_B[curve]) % p
    if p % 4 == 3:
        y = pow(y2, (p + 1) // 4, p)
    else:                                   # Tonelli-Shanks (bls12-377: q = 1 mod 4, fp/element.go Sqrt)
        y = _tonelli(y2, p)
    if y is None or y * y % p != y2:
        raise ValueError("invalid compressed coordinate: square root doesn't exist")
    if (y > (p - 1) // 2) != (m == f["large"]):
        y = p - y
    return np.concatenate([_fp_encode(x, curve), _fp_encode(y, curve)]), nb


def _tonelli(a: int, p: int):
    if a == 0:
        return 0
    if pow(a, (p - 1) // 2, p) != 1:
        return None
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, t, r = s, pow(z, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c, t, r = i, b * b % p, t * b * b % p, r * b % p
    return r


# ----------------------------------------------------------------------------------------------------------------
# batched openings at one point (kzg.go:246-420): the prover side; verification (pairings) is out of scope
# ----------------------------------------------------------------------------------------------------------------
class ErrInvalidNbDigests(MultiExpError):
    """kzg.ErrInvalidNbDigests (kzg.go:23)"""


@dataclass
class BatchOpeningProof:
    """kzg.BatchOpeningProof{H G1Affine, ClaimedValues []fr.Element} (kzg.go:66-77)"""

    H: np.ndarray
    ClaimedValues: np.ndarray


def _fr_marshal(limbs, r: int) -> bytes:
    """fr.Element.Marshal (fr/element.go:868-871): 32 bytes big-endian, canonical value"""
    return _fr_decode(np.asarray(limbs, dtype=np.uint64), r)[0].to_bytes(32, "big")


def derive_gamma(point, digests, claimed_values, hf, curve: str, *data_transcript: bytes) -> int:
    """deriveGamma (kzg.go:531-563) over fiatshamir.Transcript (fiat-shamir/transcript.go:61-131) with the single challenge
    "gamma": H("gamma" || point || digests (RawBytes) || claimed values || extra data), read big-endian and reduced mod r
    (fr.SetBytes, fr/element.go:880-903).  `hf` is a hashlib constructor (e.g. hashlib.sha256)."""
    c = curve.split("_")[0]
    r = FR_MODULUS[c]
    h = hf()
    h.update(b"gamma")
    h.update(_fr_marshal(point, r))
    for d in digests:
        h.update(g1_raw_bytes(d, c))
    for v in np.ascontiguousarray(claimed_values, dtype=np.uint64).reshape(-1, 4):
        h.update(_fr_marshal(v, r))
    for b in data_transcript:
        h.update(b)
    return int.from_bytes(h.digest(), "big") % r


def BatchOpenSinglePoint(polynomials, digests, point: np.ndarray, hf, pk: ProvingKey, *data_transcript: bytes) -> BatchOpeningProof:
    """kzg.BatchOpenSinglePoint (kzg.go:246-331): ClaimedValues[i] = f_i(point); gamma by Fiat-Shamir; the folded polynomial
    sum_i gamma^i f_i is divided by (X - point) on the host (Fr loops, as in the reference) and committed with ONE MultiExp
    over the resident bases."""
    if len(digests) != len(polynomials):
        raise ErrInvalidNbDigests("number of digests is not the same as the number of polynomials")
    c = pk.curve.split("_")[0]
    r = FR_MODULUS[c]
    polys = []
    for p in polynomials:
        p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 4)
        if p.shape[0] == 0 or p.shape[0] > pk.G1.shape[0]:
            raise ErrInvalidPolynomialSize("invalid polynomial size (larger than SRS or == 0)")
        polys.append(_fr_decode(p, r))
    a = _fr_decode(point, r)[0]
    claimed = [_eval(f, a, r) for f in polys]
    claimed_limbs = _fr_encode(claimed, r)
    gamma = derive_gamma(point, digests, claimed_limbs, hf, c, *data_transcript)
    folded_eval = claimed[-1]
    for v in reversed(claimed[:-1]):
        folded_eval = (folded_eval * gamma + v) % r
    largest = max(len(f) for f in polys)
    folded = list(polys[0]) + [0] * (largest - len(polys[0]))
    g = 1
    for f in polys[1:]:
        g = g * gamma % r
        for j, v in enumerate(f):
            folded[j] = (folded[j] + v * g) % r
    h = _divide_by_x_minus_a(folded, folded_eval, a, r)
    if not h:
        raise ErrInvalidPolynomialSize("invalid polynomial size (larger than SRS or == 0)")
    H = Commit(_fr_encode(h, r), pk)
    return BatchOpeningProof(H=H.reshape(pk.words), ClaimedValues=claimed_limbs)


def FoldProof(digests, proof: BatchOpeningProof, point: np.ndarray, hf, curve: str, *data_transcript: bytes):
    """kzg.FoldProof (kzg.go:341-380) -> (OpeningProof, folded digest): the claimed values are folded with [1, gamma, ...] on the
    host, the digests with one MultiExp (`fold`, kzg.go:506-528 -- the reference calls MultiExp for it too)."""
    from .multiexp import curve_package

    claimed = np.ascontiguousarray(proof.ClaimedValues, dtype=np.uint64).reshape(-1, 4)
    if len(digests) != claimed.shape[0]:
        raise ErrInvalidNbDigests("number of digests is not the same as the number of polynomials")
    c = curve.split("_")[0]
    r = FR_MODULUS[c]
    gamma = derive_gamma(point, digests, claimed, hf, c, *data_transcript)
    gam = [1]
    for _ in range(1, len(digests)):
        gam.append(gam[-1] * gamma % r)
    vals = _fr_decode(claimed, r)
    folded_eval = sum(v * g for v, g in zip(vals, gam)) % r
    aff_cls = curve_package(c)[0]
    pts = np.ascontiguousarray(np.stack([np.asarray(d, dtype=np.uint64).reshape(-1) for d in digests]))
    folded_digest = aff_cls().MultiExp(pts, _fr_encode(gam, r), MultiExpConfig()).limbs
    return OpeningProof(H=np.array(proof.H, dtype=np.uint64), ClaimedValue=_fr_encode([folded_eval], r)[0]), folded_digest


def decode_g1_points(curve: str, data: bytes, n: int, raw: bool = False, check_on_curve: bool = True) -> np.ndarray:
    """bulk G1Affine.SetBytes on the GPU (gmsm_g1_decode, csrc/decode.cu): n points of a homogeneous stream -> (n, words) uint64
    in Go memory layout.  Raises MultiExpError with the reference's message and the index of the first invalid point."""
    from . import _native

    cname = curve + "_g1" if not curve.endswith("_g1") else curve
    words = 2 * _words(CURVES[cname])
    per = 8 * words if raw else 4 * words
    if len(data) < n * per:
        raise EOFError("short buffer")      # io.ErrShortBuffer
    buf = np.frombuffer(data, dtype=np.uint8, count=n * per)
    out = np.zeros((n, words), dtype=np.uint64)
    rc = _native.lib().gmsm_g1_decode(CURVES[cname], buf.ctypes.data, n, 1 if raw else 0, 1 if check_on_curve else 0, out.ctypes.data)
    if rc != 0:
        raise MultiExpError(_native.last_error())
    return out


def CommitLagrange(evals, pk: ProvingKey, domain) -> np.ndarray:
    """Digest of the polynomial given by its values on `domain` (fft.Domain of this package): the evaluations go to the device
    once, FFTInverse(DIF) + BitReverse (fft.go:111-190, bitreverse.go:17-42) run there and their output -- the coefficients,
    still in device memory, Montgomery form -- feeds the MultiExp directly (gmsm_bases_multiexp_device): the canonical-form
    coefficients never visit the host.  Equals Commit(FFTInverse(evals), pk)."""
    import torch

    from .fft import DIF

    ev = np.ascontiguousarray(evals, dtype=np.uint64).reshape(-1, 4)
    if ev.shape[0] != domain.Cardinality:
        raise MultiExpError("len(a) must equal the domain cardinality")
    if ev.shape[0] == 0 or ev.shape[0] > pk.G1.shape[0]:
        raise ErrInvalidPolynomialSize("invalid polynomial size (larger than SRS or == 0)")
    dev = torch.device("cuda", domain.device)
    d = torch.from_numpy(ev.view(np.int64).reshape(-1).copy()).to(dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    domain.fft_device(d, True, DIF, False, st)          # natural in, bit-reversed out
    domain.bit_reverse_device(d, st)
    jac = pk._bases.MultiExpDevice(d, ev.shape[0], stream=st)
    w = pk.words
    return jac[:w].copy() if jac[w:].any() else np.zeros(w, dtype=np.uint64)
