"""unsafe.WriteSlice / ReadSlice image (utils/unsafe/dump_slice.go:16-76) -- CPU-only format checks."""
import importlib
import io
import struct

import numpy as np
import pytest


def test_slice_roundtrip_and_limits():
    kzg = importlib.import_module("gnark-crypto_b200.kzg")
    pts = np.arange(5 * 8, dtype=np.uint64).reshape(5, 8)
    buf = io.BytesIO()
    kzg.write_marker(buf)
    kzg.write_slice(buf, pts)
    buf.write(b"tail")
    raw = buf.getvalue()
    assert raw[:8] == struct.pack("<Q", 0xDEADBEEF) and raw[8:16] == struct.pack("<Q", 5)
    assert raw[16 : 16 + 64] == pts[0].tobytes()              # raw little-endian memory of the first element
    buf.seek(0)
    kzg.read_marker(buf)
    got = kzg.read_slice(buf, 8, max_elements=3)
    assert np.array_equal(got, pts[:3]) and buf.read() == b"tail"   # the unread elements are skipped
    buf.seek(8)
    assert np.array_equal(kzg.read_slice(buf, 8), pts)
    with pytest.raises(ValueError):
        kzg.read_marker(io.BytesIO(b"\x00" * 8))
    with pytest.raises(EOFError):
        kzg.read_slice(io.BytesIO(struct.pack("<Q", 4) + b"\x00" * 10), 8)
    e = io.BytesIO()
    kzg.write_slice(e, pts[:0])
    e.seek(0)
    assert kzg.read_slice(e, 8).shape == (0, 8)


def test_open_host_polynomial_arithmetic():
    """eval / dividePolyByXminusA (kzg.go:55-63, 567-584) as kzg.Open uses them, and the Montgomery codec"""
    import random

    kzg = importlib.import_module("gnark-crypto_b200.kzg")
    rng = random.Random(7)
    for name, r in kzg.FR_MODULUS.items():
        f = [rng.randrange(r) for _ in range(37)]
        a = rng.randrange(r)
        fa = kzg._eval(f, a, r)
        assert fa == sum(c * pow(a, i, r) for i, c in enumerate(f)) % r
        h = kzg._divide_by_x_minus_a(f, fa, a, r)
        assert len(h) == len(f) - 1
        for x in (0, 1, rng.randrange(r)):
            assert (kzg._eval(h, x, r) * (x - a) + fa) % r == kzg._eval(f, x, r)
        enc = kzg._fr_encode(f, r)
        assert kzg._fr_decode(enc, r) == f
    # fr.One of bn254 in Montgomery form (ecc/bn254/fr/element.go:227)
    one = kzg._fr_encode([1], kzg.FR_MODULUS["bn254"])[0]
    assert [int(x) for x in one] == [12436184717236109307, 3962172157175319849, 7381016538464732718, 1011752739694698287]
