"""Limb-exact simulation of the even/odd-accumulator Montgomery multiplication used by
gnark-crypto_b200/csrc/field.cuh (mad.lo.cc/madc.hi.cc chains, 32-bit limbs).  Validates the
algorithm (and the no-carry-out claims, as asserts) against big-int arithmetic on the CPU,
because the PTX itself can only run on the GPU box."""
import random
import sys

M32 = 0xFFFFFFFF


def chain_mad(acc, lo_idx0, xs, y, addend, addend_idx0, cin):
    """for k, x in enumerate(xs): pair (acc[lo_idx0+2k], acc[lo_idx0+2k+1]) = x*y + (addend pair) + carry.
    Returns carry out."""
    c = cin
    for k, x in enumerate(xs):
        prod = x * y
        lo, hi = prod & M32, prod >> 32
        i = lo_idx0 + 2 * k
        ai = addend_idx0 + 2 * k
        s = lo + addend[ai] + c
        acc_lo = s & M32
        c = s >> 32
        s = hi + addend[ai + 1] + c
        acc_hi = s & M32
        c = s >> 32
        acc[i], acc[i + 1] = acc_lo, acc_hi
    return c


def montmul(a, b, p, N, qinv):
    al = [(a >> (32 * i)) & M32 for i in range(N)]
    bl = [(b >> (32 * i)) & M32 for i in range(N)]
    pl = [(p >> (32 * i)) & M32 for i in range(N)]
    a_even, a_odd = al[0::2], al[1::2]
    p_even, p_odd = pl[0::2], pl[1::2]
    # accumulators have N+2 slots: [0..N-1] limbs, [N] carry limb, [N+1] always zero
    Ev = [0] * (N + 2)
    Od = [0] * (N + 2)
    d = 0
    k = 0
    for i in range(N):
        # step 1: Ev += a_even*b_i (carry in k), carry -> Ev[N]
        c = chain_mad(Ev, 0, a_even, bl[i], Ev, 0, k)
        assert Ev[N] == 0
        Ev[N] = c
        # step 2: Od already rshifted by construction (see swap below): Od += a_odd*b_i
        c = chain_mad(Od, 0, a_odd, bl[i], Od, 0, 0)
        assert c == 0
        # step 3
        m = (((Ev[0] + d) & M32) * qinv) & M32
        # step 4
        c = chain_mad(Ev, 0, p_even, m, Ev, 0, 0)
        Ev[N] += c
        assert Ev[N] <= M32
        # step 5
        c = chain_mad(Od, 0, p_odd, m, Od, 0, 0)
        assert c == 0
        # step 6: frame shift
        s = Ev[0] + d
        assert s & M32 == 0
        k = s >> 32
        d = Ev[1]
        newOd = Ev[2 : N + 1] + [0, 0, 0]  # Ev[2..N] -> N-1 limbs, rest zero
        newEv = Od[:N] + [0, 0]
        assert Od[N] == 0 and Od[N + 1] == 0
        Ev, Od = newEv, newOd[: N + 2]
    # merge: R = Ev + d + k + 2^32*Od
    R = sum(Ev[i] << (32 * i) for i in range(N)) + d + k + (sum(Od[i] << (32 * i) for i in range(N)) << 32)
    assert R < (1 << (32 * N)), "merge overflow"
    if R >= p:
        R -= p
    assert R < p
    return R


def main():
    fields = {
        "bn254_fp": 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
        "bn254_fr": 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
        "bls12381_fp": 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
        "bls12381_fr": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    }
    rng = random.Random(1)
    for name, p in fields.items():
        N = (p.bit_length() + 31) // 32
        N += N & 1
        Rinv = pow(1 << (32 * N), -1, p)
        qinv = (-pow(p, -1, 1 << 32)) % (1 << 32)
        vals = [0, 1, p - 1, p - 2, (1 << (32 * N)) % p, M32, (p - 1) // 2] + [rng.randrange(p) for _ in range(300)]
        for a in vals:
            for b in vals[:20]:
                assert montmul(a, b, p, N, qinv) == a * b * Rinv % p, (name, hex(a), hex(b))
        print(name, "N=%d ok" % N)


if __name__ == "__main__":
    main()
