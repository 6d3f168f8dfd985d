"""Pins the oracle's Fr FFT restatement (next-row N3) with the identities the reference's own tests use
(ecc/bn254/fr/fft/fft_test.go: FFT/FFTInverse round trips, DIF+DIT compositions, coset variants, evaluation
on the domain) plus the O(n^2) definition.  CPU only."""
import random

import pytest

from oracle import oracle as O


@pytest.mark.parametrize("frname", ["bn254_fr", "bls12381_fr", "bls12377_fr"])
def test_domain_constants(frname):
    q = O.FIELDS[frname].q
    for lg in (1, 4, 10, 20):
        d = O.FFTDomain(frname, (1 << lg) - (1 if lg > 1 else 0))
        assert d.cardinality == 1 << lg
        assert pow(d.generator, 1 << lg, q) == 1 and pow(d.generator, 1 << (lg - 1), q) == q - 1
        assert d.generator * d.generator_inv % q == 1 and d.cardinality * d.cardinality_inv % q == 1
    with pytest.raises(ValueError):
        O.FFTDomain(frname, 1 << 48)
    # maximal 2-adic order really is max_order: root^(2^(max-1)) = -1
    P = O.FFT_PARAMS[frname]
    assert pow(P["root"], 1 << (P["max_order"] - 1), q) == q - 1


@pytest.mark.parametrize("frname", ["bn254_fr", "bls12381_fr", "bls12377_fr"])
def test_fft_matches_definition_and_roundtrips(frname):
    q = O.FIELDS[frname].q
    rng = random.Random(4)
    n = 32
    d = O.FFTDomain(frname, n)
    a = [rng.randrange(q) for _ in range(n)]
    evals = [sum(a[i] * pow(d.generator, i * k, q) for i in range(n)) % q for k in range(n)]
    # DIF: natural in, bit-reversed out
    out = d.fft(list(a), O.DIF)
    assert O.bit_reverse(list(out)) == evals
    # DIT: bit-reversed in, natural out
    assert d.fft(O.bit_reverse(list(a)), O.DIT) == evals
    # the compositions gnark uses to avoid explicit bit reversals (fft_test.go)
    assert d.fft_inverse(d.fft(list(a), O.DIF), O.DIT) == a
    assert d.fft_inverse(d.fft(O.bit_reverse(list(a)), O.DIT), O.DIF) == O.bit_reverse(list(a))
    # coset: evaluations on u * <w>
    cos = [sum(a[i] * pow(d.shift * pow(d.generator, k, q) % q, i, q) for i in range(n)) % q for k in range(n)]
    assert O.bit_reverse(d.fft(list(a), O.DIF, coset=True)) == cos
    assert d.fft(O.bit_reverse(list(a)), O.DIT, coset=True) == cos
    assert d.fft_inverse(d.fft(list(a), O.DIF, coset=True), O.DIT, coset=True) == a
    assert d.fft_inverse(d.fft(O.bit_reverse(list(a)), O.DIT, coset=True), O.DIF, coset=True) == O.bit_reverse(list(a))
    # custom shift (WithShift option)
    d2 = O.FFTDomain(frname, n, shift=12345)
    cos2 = [sum(a[i] * pow(12345 * pow(d2.generator, k, q) % q, i, q) for i in range(n)) % q for k in range(n)]
    assert d2.fft(O.bit_reverse(list(a)), O.DIT, coset=True) == cos2
