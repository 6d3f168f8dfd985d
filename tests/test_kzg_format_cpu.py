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
