"""sm_100a device arithmetic (PTX Montgomery multiplication, carry chains, point formulas, digit
recoding) against the oracle, element-wise through the C ABI's test hooks."""
import numpy as np
import pytest

from oracle import cref
from oracle import oracle as O
from tests import opcases

pytestmark = pytest.mark.gpu


def _runner(g):
    import gnark_crypto_b200 as pkg
    from importlib import import_module

    mx = import_module("gnark-crypto_b200.multiexp")

    def run(op, a, b, out_words):
        return mx.test_op(g, op, a, b, out_words)

    return run


@pytest.mark.parametrize("g", list(O.GROUPS))
def test_field_ops_device(g):
    opcases.check_field_ops(O.GROUPS[g], _runner(g))
    opcases.check_fr_from_mont(O.GROUPS[g], _runner(g))


@pytest.mark.parametrize("g", list(O.GROUPS))
def test_point_ops_device(g):
    opcases.check_point_ops(O.GROUPS[g], _runner(g))


@pytest.mark.parametrize("g", ["bn254_g1", "bls12381_g1", "bls12377_g1"])
def test_field_mul_bulk_random(g):
    """20k random products per base field against the C oracle (the Python one is pinned to it)"""
    G = O.GROUPS[g]
    f = G.K.f
    n = 20000
    rng = np.random.default_rng(7)
    L = f.limbs
    a = rng.integers(0, 2**63, size=(n, L), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, L), dtype=np.uint64)
    b = rng.integers(0, 2**63, size=(n, L), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, L), dtype=np.uint64)
    top = (1 << (f.bits - 64 * (L - 1) - 1)) - 1  # < q for sure
    a[:, -1] &= np.uint64(top)
    b[:, -1] &= np.uint64(top)
    want = cref.field_op({"bn254_g1": 0, "bls12381_g1": 2, "bls12377_g1": 4}[g], 0, a, b)
    got = _runner(g)(opcases.OPS["FMUL"], opcases.u32(a), opcases.u32(b), 2 * L)
    assert np.array_equal(got.view(np.uint64), want)


@pytest.mark.parametrize("g", ["bn254_g1", "bls12381_g1", "bls12377_g1"])
@pytest.mark.parametrize("c", [2, 3, 5, 8, 11, 13, 15, 16, 17, 20, 23, 24])
def test_digits_match_partition_scalars(g, c):
    from importlib import import_module

    mx = import_module("gnark-crypto_b200.multiexp")
    G = O.GROUPS[g]
    fr = G.fr
    s = cref.random_scalars(g, 3000, 1000 + c)
    specials = [0, 1, fr.q - 1, fr.q - 2, 1 << (fr.bits - 1), (1 << c) - 1, 1 << (c - 1), (1 << (c - 1)) - 1]
    s[: len(specials)] = G.encode_scalars(specials)
    got = mx.test_digits(g, c, s)
    want = cref.partition_scalars(g, s, c)
    assert np.array_equal(got, want)
