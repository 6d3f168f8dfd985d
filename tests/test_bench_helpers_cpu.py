"""bench.py's host-side helpers (no GPU): synthetic scalars are uniform-looking reduced field elements like fr.SetRandom's
(fr/element.go:302-343), deterministic per seed; the multiplier-pipe accounting matches the operation counts of DESIGN.md."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_synth_scalars_reduced_and_deterministic():
    b = _bench()
    for bits in (253, 254, 255):
        q = b.FR_MOD[bits]
        s = b.synth_scalars(5000, bits, 42)
        vals = [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in s]
        assert all(v < q for v in vals) and len(set(vals)) == len(vals)
        assert max(vals).bit_length() >= bits - 2           # the whole range is used
        assert np.array_equal(s, b.synth_scalars(5000, bits, 42)) and not np.array_equal(s, b.synth_scalars(5000, bits, 43))
    # the moduli are the reference's (fr/element.go:44-49 of the three curves): they match the oracle's pinned fields
    from oracle import oracle as O

    assert b.FR_MOD[254] == O.FIELDS["bn254_fr"].q and b.FR_MOD[255] == O.FIELDS["bls12381_fr"].q and b.FR_MOD[253] == O.FIELDS["bls12377_fr"].q
    # every group's scalar modulus, bit length, point size and scalar width as the oracle has them (bls24-315 / bls24-317 share
    # their bit lengths with bls12-377 / bls12-381: looked up by group)
    assert set(b.CURVE_BITS) == set(O.GROUPS)
    for g, G in O.GROUPS.items():
        assert b.fr_mod(g) == G.fr.q and b.CURVE_BITS[g] == G.fr.bits and b.scalar_words(b.CURVE_BITS[g]) == G.fr.limbs, g
        assert b.AFF_BYTES[g] == 8 * G.aff_words, g


def test_multiplier_pipe_accounting():
    b = _bench()
    r = b.int_pipe_fraction("bn254_g1", 6.0e9, 1965.0)
    assert r["wide_mads_per_mixed_add"] == 10 * (2 * 8 * 8 + 8) == 1360
    assert abs(r["peak"] - 148 * 32 * 1965e6) < 1 and abs(r["frac"] - 6.0e9 * 1360 / (148 * 32 * 1965e6)) < 1e-12
    assert b.int_pipe_fraction("bls12381_g1", 1.0, 1000.0)["wide_mads_per_mixed_add"] == 10 * (2 * 12 * 12 + 12)
    assert b.int_pipe_fraction("bn254_g2", 1.0, 1000.0)["wide_mads_per_mixed_add"] == 28 * (2 * 8 * 8 + 8)   # Fp2: M = 3, S = 2 Fp products
    assert b.int_pipe_fraction("bn254_g1", 1.0, None) is None
