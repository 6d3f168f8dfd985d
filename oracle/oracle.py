"""CPU oracle for the gnark-crypto MultiExp (MSM) hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-Python (big-int) restatement of the reference algorithm for
`ecc/<curve>.G1Jac.MultiExp` / `G2Jac.MultiExp`.  It is imported ONLY by `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg, as the checker.  The
product path (gnark-crypto_b200/) never imports it.

Parity pinning: the reference is Go and cannot be built here (no Go toolchain).  The
oracle is pinned against everything the reference's own tests hold for this path:
  * the generated field constants (qInvNeg, rSquare, One) -- see FIELDS[..]['pin'],
    cited per field below;
  * the curve generators (ecc/bn254/bn254.go:111-119, ecc/bls12-381/bls12-381.go:107-116);
  * the hash-to-curve vectors with explicit affine points, identity P = Q0 + Q1 on
    bn254 (ecc/bn254/hash_vectors_test.go:29-83, G2 :84+) -> tests/golden/;
  * the relational MSM properties of ecc/bn254/multiexp_test.go:63-299 restated in
    tests/test_oracle.py.
There is no stored golden MSM output in the reference.

All `file:line` citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# Fields.  Everything else is derived, then pinned against the
# generated constants quoted from the reference.
# --------------------------------------------------------------------------------------


class Field:
    """Prime field with gnark-crypto's Montgomery representation.

    ecc/bn254/fp/element.go:24-36 : an Element is [Limbs]uint64, little-endian, value*R mod q.
    """

    def __init__(self, name, q, limbs, pin):
        self.name = name
        self.q = q
        self.limbs = limbs
        self.bits = q.bit_length()
        self.R = 1 << (64 * limbs)
        self.Rmod = self.R % q  # "One" (SetOne, fp/element.go:194-200)
        self.R2 = (self.R * self.R) % q  # rSquare (fp/element.go:773)
        self.Rinv = pow(self.R, -1, q)
        # field/generator/config/field_config.go:140-183 : qInvNeg = -q^{-1} mod 2^64
        self.qinvneg = (-pow(q, -1, 1 << 64)) % (1 << 64)
        self.pin = pin

    # ---- Montgomery edges (fp/element.go:593-642 fromMont, :782-784 toMont) ----
    def to_mont(self, a: int) -> int:
        return (a * self.R) % self.q

    def from_mont(self, m: int) -> int:
        return (m * self.Rinv) % self.q

    def to_limbs(self, v: int):
        return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(self.limbs)]

    @staticmethod
    def from_limbs(l) -> int:
        v = 0
        for i, w in enumerate(l):
            v |= int(w) << (64 * i)
        return v

    # ---- textbook CIOS on 64-bit limbs: _mulGeneric, fp/element.go:470-591 ----
    def mont_mul_cios(self, x: int, y: int) -> int:
        """x, y Montgomery-form integers < q -> x*y*R^-1 mod q, computed limb by limb
        exactly as the reference's generic CIOS (used to pin the device CIOS)."""
        N = self.limbs
        M = 0xFFFFFFFFFFFFFFFF
        xl, yl, ql = self.to_limbs(x), self.to_limbs(y), self.to_limbs(self.q)
        t = [0] * (N + 2)
        for i in range(N):
            c = 0
            for j in range(N):
                s = t[j] + xl[j] * yl[i] + c
                t[j] = s & M
                c = s >> 64
            s = t[N] + c
            t[N] = s & M
            t[N + 1] = s >> 64
            m = (t[0] * self.qinvneg) & M
            s = t[0] + m * ql[0]
            c = s >> 64
            for j in range(1, N):
                s = t[j] + m * ql[j] + c
                t[j - 1] = s & M
                c = s >> 64
            s = t[N] + c
            t[N - 1] = s & M
            t[N] = t[N + 1] + (s >> 64)
        z = self.from_limbs(t[:N]) + (t[N] << (64 * N))
        if z >= self.q:  # fp/element.go:583-590
            z -= self.q
        return z


FIELDS = {
    # ecc/bn254/fp/element.go:31,39-50,71,195-198,773
    "bn254_fp": Field(
        "bn254_fp",
        0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
        4,
        dict(
            qinvneg=9786893198990664585,
            rsquare=[17522657719365597833, 13107472804851548667, 5164255478447964150, 493319470278259999],
        ),
    ),
    # ecc/bn254/fr/element.go:31,39-49,71,773-778
    "bn254_fr": Field(
        "bn254_fr",
        0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
        4,
        dict(
            qinvneg=14042775128853446655,
            rsquare=[1997599621687373223, 6052339484930628067, 10108755138030829701, 150537098327114917],
        ),
    ),
    # ecc/bls12-381/fp/element.go:31,38-51,75,928-935
    "bls12381_fp": Field(
        "bls12381_fp",
        0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
        6,
        dict(
            qinvneg=9940570264628428797,
            rsquare=[
                17644856173732828998,
                754043588434789617,
                10224657059481499349,
                7488229067341005760,
                11130996698012816685,
                1267921511277847466,
            ],
        ),
    ),
    # ecc/bls12-381/fr/element.go:31,39-49,71,773-778
    "bls12381_fr": Field(
        "bls12381_fr",
        0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
        4,
        dict(
            qinvneg=18446744069414584319,
            rsquare=[14526898881837571181, 3129137299524312099, 419701826671360399, 524908885293268753],
        ),
    ),
    # next-row N4: ecc/bls12-377/fp/element.go:28-31,75,928-935 ; fr/element.go:28-31,71,773-778
    "bls12377_fp": Field(
        "bls12377_fp",
        0x1AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001,
        6,
        dict(qinvneg=9586122913090633727,
             rsquare=[13224372171368877346, 227991066186625457, 2496666625421784173, 13825906835078366124, 9475172226622360569,
                      30958721782860680]),
    ),
    "bls12377_fr": Field(
        "bls12377_fr",
        0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
        4,
        dict(qinvneg=725501752471715839, rsquare=[2726216793283724667, 14712177743343147295, 12091039717619697043, 81024008013859129]),
    ),
    # N4 remainder: ecc/secp256k1/fp/element.go:31,71,801-806 ; fr/element.go:31,71,801-806 (moduli without a spare top bit)
    "secp256k1_fp": Field(
        "secp256k1_fp",
        0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2F,
        4,
        dict(qinvneg=15580212934572586289, rsquare=[8392367050913, 1, 0, 0]),
    ),
    "secp256k1_fr": Field(
        "secp256k1_fr",
        0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141,
        4,
        dict(qinvneg=5408259542528602431, rsquare=[9902555850136342848, 8364476168144746616, 16616019711348246470, 11342065889886772165]),
    ),
    # ecc/bw6-761/fp/element.go:31,87,1546-1559 (12 words) ; fr/element.go:31,75,928-935 (6 words, = the bls12-377 base field)
    "bw6761_fp": Field(
        "bw6761_fp",
        0x122E824FB83CE0AD187C94004FAFF3EB926186A81D14688528275EF8087BE41707BA638E584E91903CEBAFF25B423048689C8ED12F9FD9071DCD3DC73EBFF2E98A116C25667A8F8160CF8AEEAF0A437E6913E6870000082F49D00000000008B,
        12,
        dict(qinvneg=744663313386281181,
             rsquare=[14305184132582319705, 8868935336694416555, 9196887162930508889, 15486798265448570248, 5402985275949444416,
                      10893197322525159598, 3204916688966998390, 12417238192559061753, 12426306557607898622, 1305582522441154384,
                      10311846026977660324, 48736111365249031]),
    ),
    "bw6761_fr": Field(
        "bw6761_fr",
        0x1AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001,
        6,
        dict(qinvneg=9586122913090633727,
             rsquare=[13224372171368877346, 227991066186625457, 2496666625421784173, 13825906835078366124, 9475172226622360569,
                      30958721782860680]),
    ),
    # N4: ecc/bls24-315/fp/element.go:31,73 fr:31,71 ; ecc/bls24-317/fp:31,73 fr:31,71 ; ecc/bw6-633/fp:31,83 fr:31,73 (qInvNeg, rSquare pins)
    "bls24315_fp": Field("bls24315_fp", 0x4C23A02B586D650D3F7498BE97C5EAFDEC1D01AA27A1AE0421EE5DA52BDE5026FE802FF40300001, 5, dict(qinvneg=8083954730842193919, rsquare=[7746605402484284438, 6457291528853138485, 14067144135019420374, 14705958577488011058, 150264569250089173])),
    "bls24315_fr": Field("bls24315_fr", 0x196DEAC24A9DA12B25FC7EC9CF927A98C8C480ECE644E36419D0C5FD00C00001, 4, dict(qinvneg=2184305180030271487, rsquare=[6242551132904523857, 16951295617263545407, 10923821274252739203, 584663452775307866])),
    "bls24317_fp": Field("bls24317_fp", 0x1058CA226F60892CF28FC5A0B7F9D039169A61E684C73446D6F339E43424BF7E8D512E565DAB2AAB, 5, dict(qinvneg=6176088765535387645, rsquare=[8184925746953654484, 11847028797714522427, 6382817893761672566, 4341726315782040335, 1146553493836047074])),
    "bls24317_fr": Field("bls24317_fr", 0x443F917EA68DAFC2D0B097F28D83CD491CD1E79196BF0E7AF000000000000001, 4, dict(qinvneg=17293822569102704639, rsquare=[14966889745918050766, 10836803306611491707, 10398613988537905008, 4216292045776253362])),
    "bw6633_fp": Field("bw6633_fp", 0x126633CC0F35F63FC1A174F01D72AB5A8FCD8C75D79D2C74E59769AD9BBDA2F8152A6C0FADEA490B8DA9F5E83F57C497E0E8850EDBDA407D7B5CE7AB839C2253D369BD31147F73CD74916EA4570000D, 10, dict(qinvneg=13046692460116554043, rsquare=[7358459907925294924, 14414180951914241931, 16619482658146888203, 760736596725344926, 12753071240931896792, 13425190760400245818, 12591714441439252728, 15325516497554583360, 5301152003049442834, 35368377961363834])),
    "bw6633_fr": Field("bw6633_fr", 0x4C23A02B586D650D3F7498BE97C5EAFDEC1D01AA27A1AE0421EE5DA52BDE5026FE802FF40300001, 5, dict(qinvneg=8083954730842193919, rsquare=[7746605402484284438, 6457291528853138485, 14067144135019420374, 14705958577488011058, 150264569250089173])),
}


# --------------------------------------------------------------------------------------
# Coordinate-field adapters: Fp (ints mod q) and Fp2 = Fp[u]/(u^2+1) (tuples).
# Values here are PLAIN integers (not Montgomery); SURVEY.md A.1 -- Montgomery form is
# closed under the ops, so we encode only at the edges and the encoded result is unique.
# --------------------------------------------------------------------------------------


class FpOps:
    ext = 1

    def __init__(self, field: Field):
        self.f = field
        self.q = field.q
        self.zero = 0
        self.one = 1

    def add(self, a, b):
        return (a + b) % self.q

    def sub(self, a, b):
        return (a - b) % self.q

    def neg(self, a):
        return (-a) % self.q

    def dbl(self, a):
        return (2 * a) % self.q

    def mul(self, a, b):
        return (a * b) % self.q

    def sqr(self, a):
        return (a * a) % self.q

    def inv(self, a):
        # fp.Inverse(0) = 0 (fp/element.go:1170-1172)
        return 0 if a == 0 else pow(a, -1, self.q)

    def is_zero(self, a):
        return a == 0

    def from_int(self, v):
        return v % self.q

    # memory edges: L u64 limbs of the Montgomery representation
    def encode(self, a):
        return self.f.to_limbs(self.f.to_mont(a))

    def decode(self, limbs):
        return self.f.from_mont(Field.from_limbs(limbs))

    @property
    def words(self):
        return self.f.limbs


class Fp2Ops:
    """E2 = A0 + A1*u, u^2 = -1 (ecc/bn254/internal/fptower/e2.go:14-16, e2_bn254.go:28-73;
    same formulas ecc/bls12-381/internal/fptower/e2_bls381.go:16-74)."""

    ext = 2

    def __init__(self, field: Field, nonres: int = -1):
        # u^2 = nonres: -1 for bn254 / bls12-381, -5 for bls12-377 (e2_bls377.go:12-80)
        self.f = field
        self.q = field.q
        self.beta = nonres
        self.zero = (0, 0)
        self.one = (1, 0)

    def add(self, a, b):
        return ((a[0] + b[0]) % self.q, (a[1] + b[1]) % self.q)

    def sub(self, a, b):
        return ((a[0] - b[0]) % self.q, (a[1] - b[1]) % self.q)

    def neg(self, a):
        return ((-a[0]) % self.q, (-a[1]) % self.q)

    def dbl(self, a):
        return ((2 * a[0]) % self.q, (2 * a[1]) % self.q)

    def mul(self, a, b):
        # e2_bn254.go:28-38 (Karatsuba)
        q = self.q
        ac = a[0] * b[0] % q
        bd = a[1] * b[1] % q
        t = (a[0] + a[1]) * (b[0] + b[1]) % q
        return ((ac + self.beta * bd) % q, (t - ac - bd) % q)

    def sqr(self, a):
        # e2_bn254.go:41-51
        q = self.q
        return ((a[0] * a[0] + self.beta * a[1] * a[1]) % q, (2 * a[0] * a[1]) % q)

    def inv(self, a):
        # e2_bn254.go:61-73
        q = self.q
        n = (a[0] * a[0] - self.beta * a[1] * a[1]) % q
        ni = 0 if n == 0 else pow(n, -1, q)
        return (a[0] * ni % q, (-a[1] * ni) % q)

    def is_zero(self, a):
        return a[0] == 0 and a[1] == 0

    def from_int(self, v):
        return (v % self.q, 0)

    def encode(self, a):
        return self.f.to_limbs(self.f.to_mont(a[0])) + self.f.to_limbs(self.f.to_mont(a[1]))

    def decode(self, limbs):
        L = self.f.limbs
        return (self.f.from_mont(Field.from_limbs(limbs[:L])), self.f.from_mont(Field.from_limbs(limbs[L:])))

    @property
    def words(self):
        return 2 * self.f.limbs


# --------------------------------------------------------------------------------------
# Groups
# --------------------------------------------------------------------------------------

INF_AFF = None  # python-side marker; in memory the affine infinity is (0,0) (g1.go:41-47,178-180)


class Group:
    """Short-Weierstrass group y^2 = x^3 + b (a = 0; g1.go:804) over FpOps / Fp2Ops.

    Affine points are (x, y) tuples of coordinate-field values, or (zero, zero) for infinity
    exactly as in Go memory.  xyzz points are [X, Y, ZZ, ZZZ] lists; infinity iff ZZ == 0
    (g1.go:688-699).
    """

    def __init__(self, name, K, fr: Field, b, gen):
        self.name = name
        self.K = K
        self.fr = fr
        self.b = b
        self.gen = gen

    # ---- predicates ----
    def aff_is_inf(self, a):
        return self.K.is_zero(a[0]) and self.K.is_zero(a[1])

    def aff_inf(self):
        return (self.K.zero, self.K.zero)

    def is_on_curve(self, a):
        K = self.K
        if self.aff_is_inf(a):
            return True
        return K.sqr(a[1]) == K.add(K.mul(K.sqr(a[0]), a[0]), self.b)

    def aff_neg(self, a):
        return (a[0], self.K.neg(a[1]))

    # ---- xyzz formulas, SURVEY.md A.4 / A.5 ----
    def xyzz_inf(self):
        K = self.K
        return [K.one, K.one, K.zero, K.zero]  # SetInfinity g1.go:688-694

    def _double_mixed(self, a, negate):
        # doubleMixed g1.go:962-985 ; doubleNegMixed :933-957
        K = self.K
        y = K.neg(a[1]) if negate else a[1]
        U = K.dbl(y)
        V = K.sqr(U)
        W = K.mul(U, V)
        S = K.mul(a[0], V)
        XX = K.sqr(a[0])
        M = K.add(K.dbl(XX), XX)
        S2 = K.dbl(S)
        L = K.mul(W, y)
        X3 = K.sub(K.sqr(M), S2)
        Y3 = K.sub(K.mul(K.sub(S, X3), M), L)
        return [X3, Y3, V, W]

    def add_mixed(self, p, a, negate=False):
        """p (xyzz, mutated copy returned) += a (affine), or -= a if negate.
        addMixed g1.go:822-873, subMixed :878-930."""
        K = self.K
        if self.aff_is_inf(a):
            return p
        ay = K.neg(a[1]) if negate else a[1]
        if K.is_zero(p[2]):
            return [a[0], ay, K.one, K.one]
        P = K.sub(K.mul(a[0], p[2]), p[0])
        R = K.sub(K.mul(ay, p[3]), p[1])
        if K.is_zero(P):
            if K.is_zero(R):
                return self._double_mixed(a, negate)
            return [p[0], p[1], K.zero, K.zero]
        PP = K.sqr(P)
        PPP = K.mul(P, PP)
        Q = K.mul(p[0], PP)
        RR = K.sqr(R)
        X3 = K.sub(K.sub(RR, PPP), K.dbl(Q))
        Y3 = K.sub(K.mul(K.sub(Q, X3), R), K.mul(p[1], PPP))
        return [X3, Y3, K.mul(p[2], PP), K.mul(p[3], PPP)]

    def xyzz_double(self, q):
        # double g1.go:795-817 (valid for infinity)
        K = self.K
        U = K.dbl(q[1])
        V = K.sqr(U)
        W = K.mul(U, V)
        S = K.mul(q[0], V)
        XX = K.sqr(q[0])
        M = K.add(K.dbl(XX), XX)
        U2 = K.mul(W, q[1])
        X3 = K.sub(K.sub(K.sqr(M), S), S)
        Y3 = K.sub(K.mul(K.sub(S, X3), M), U2)
        return [X3, Y3, K.mul(V, q[2]), K.mul(W, q[3])]

    def xyzz_add(self, p, q):
        # add g1.go:736-788
        K = self.K
        if K.is_zero(q[2]):
            return p
        if K.is_zero(p[2]):
            return list(q)
        U2 = K.mul(q[0], p[2])
        U1 = K.mul(p[0], q[2])
        S2 = K.mul(q[1], p[3])
        S1 = K.mul(p[1], q[3])
        P = K.sub(U2, U1)
        R = K.sub(S2, S1)
        if K.is_zero(P):
            if K.is_zero(R):
                return self.xyzz_double(q)
            return [p[0], p[1], K.zero, K.zero]
        PP = K.sqr(P)
        PPP = K.mul(P, PP)
        Q = K.mul(U1, PP)
        V = K.mul(S1, PPP)
        X3 = K.sub(K.sub(K.sub(K.sqr(R), PPP), Q), Q)
        Y3 = K.sub(K.mul(K.sub(Q, X3), R), V)
        return [X3, Y3, K.mul(K.mul(p[2], q[2]), PP), K.mul(K.mul(p[3], q[3]), PPP)]

    # ---- conversions, SURVEY.md A.7 ----
    def xyzz_to_jac(self, p):
        """unsafeFromJacExtended g1.go:726-731 : (ZZ^2 X, ZZZ^2 Y, ZZZ); infinity -> (0,0,0)."""
        K = self.K
        if K.is_zero(p[2]):
            # ZZ = ZZZ = 0 -> the products are all zero
            return (K.zero, K.zero, K.zero)
        return (K.mul(K.sqr(p[2]), p[0]), K.mul(K.sqr(p[3]), p[1]), p[3])

    def jac_to_affine(self, j):
        """FromJacobian g1.go:150-166"""
        K = self.K
        if K.is_zero(j[2]):
            return self.aff_inf()
        a = K.inv(j[2])
        b = K.sqr(a)
        return (K.mul(j[0], b), K.mul(K.mul(j[1], b), a))

    def xyzz_to_affine(self, p):
        return self.jac_to_affine(self.xyzz_to_jac(p))

    # ---- plain affine arithmetic, used only to build inputs / closed forms in tests ----
    def aff_add(self, p, q):
        K = self.K
        if self.aff_is_inf(p):
            return q
        if self.aff_is_inf(q):
            return p
        if p[0] == q[0]:
            if p[1] == q[1]:
                lam = K.mul(K.mul(K.from_int(3), K.sqr(p[0])), K.inv(K.dbl(p[1])))
            else:
                return self.aff_inf()
        else:
            lam = K.mul(K.sub(q[1], p[1]), K.inv(K.sub(q[0], p[0])))
        x3 = K.sub(K.sub(K.sqr(lam), p[0]), q[0])
        y3 = K.sub(K.mul(lam, K.sub(p[0], x3)), p[1])
        return (x3, y3)

    def scalar_mul(self, p, k: int):
        """double-and-add on xyzz; k any non-negative integer."""
        acc = self.xyzz_inf()
        if k < 0:
            p, k = self.aff_neg(p), -k
        for bit in bin(k)[2:] if k else "":
            acc = self.xyzz_double(acc)
            if bit == "1":
                acc = self.add_mixed(acc, p)
        return self.xyzz_to_affine(acc)

    # ---- memory layout: n x {X, Y} u64 LE Montgomery; infinity = zeros ----
    @property
    def aff_words(self):
        return 2 * self.K.words

    def encode_affine(self, pts) -> np.ndarray:
        out = np.zeros((len(pts), self.aff_words), dtype=np.uint64)
        for i, a in enumerate(pts):
            if self.aff_is_inf(a):
                continue
            out[i, :] = np.array(self.K.encode(a[0]) + self.K.encode(a[1]), dtype=np.uint64)
        return out

    def decode_affine(self, arr) -> list:
        w = self.K.words
        arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 2 * w)
        return [(self.K.decode([int(v) for v in r[:w]]), self.K.decode([int(v) for v in r[w:]])) for r in arr]

    def encode_jac(self, j) -> np.ndarray:
        return np.array(self.K.encode(j[0]) + self.K.encode(j[1]) + self.K.encode(j[2]), dtype=np.uint64)

    def decode_jac(self, arr):
        w = self.K.words
        l = [int(v) for v in np.asarray(arr, dtype=np.uint64).reshape(-1)]
        return (self.K.decode(l[:w]), self.K.decode(l[w : 2 * w]), self.K.decode(l[2 * w : 3 * w]))

    def encode_scalars(self, ks) -> np.ndarray:
        """integers in [0, r) -> n x 4 u64, Montgomery form (what Go passes)."""
        out = np.zeros((len(ks), self.fr.limbs), dtype=np.uint64)
        for i, k in enumerate(ks):
            out[i, :] = np.array(self.fr.to_limbs(self.fr.to_mont(k % self.fr.q)), dtype=np.uint64)
        return out

    def decode_scalars(self, arr) -> list:
        arr = np.asarray(arr, dtype=np.uint64).reshape(-1, self.fr.limbs)
        return [self.fr.from_mont(Field.from_limbs([int(v) for v in r])) for r in arr]


def _mk_groups():
    bn_fp, bn_fr = FIELDS["bn254_fp"], FIELDS["bn254_fr"]
    bl_fp, bl_fr = FIELDS["bls12381_fp"], FIELDS["bls12381_fr"]
    g = {}
    # ecc/bn254/bn254.go:12-13,105-119 : Y^2 = X^3 + 3, generator (1, 2)
    K = FpOps(bn_fp)
    g["bn254_g1"] = Group("bn254_g1", K, bn_fr, 3, (1, 2))
    # twist b' = 3/(9+u)  (bn254.go:106-109); generator bn254.go:115-118
    K2 = Fp2Ops(bn_fp)
    bt = K2.mul(K2.inv((9, 1)), (3, 0))
    g["bn254_g2"] = Group(
        "bn254_g2",
        K2,
        bn_fr,
        bt,
        (
            (
                10857046999023057135944570762232829481370756359578518086990519993285655852781,
                11559732032986387107991004021392285783925812861821192530917403151452391805634,
            ),
            (
                8495653923123431417604973247489272438418190587263600148770280649306958101930,
                4082367875863433681332203403145435568316851327593401208105741076214120093531,
            ),
        ),
    )
    # ecc/bls12-381/bls12-381.go:9-10,100-116 : Y^2 = X^3 + 4 ; twist b' = 4(1+u)
    K = FpOps(bl_fp)
    g["bls12381_g1"] = Group(
        "bls12381_g1",
        K,
        bl_fr,
        4,
        (
            3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
            1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569,
        ),
    )
    K2 = Fp2Ops(bl_fp)
    g["bls12381_g2"] = Group(
        "bls12381_g2",
        K2,
        bl_fr,
        K2.mul((1, 1), (4, 0)),
        (
            (
                352701069587466618187139116011060144890029952792775240219908644239793785735715026873347600343865175952761926303160,
                3059144344244213709971259814753781636986470325476647558659373206291635324768958432433509563104347017837885763365758,
            ),
            (
                1985150602287291935568054521177171638300868978215655730859378665066344726373823718423869104263333984641494340347905,
                927553665492332455747201965776037880757740193453592970025027978793976877002675564980949289727957565575433344219582,
            ),
        ),
    )
    # next-row N4: ecc/bls12-377/bls12-377.go:9,102,107-108 : Y^2 = X^3 + 1
    g["bls12377_g1"] = Group(
        "bls12377_g1",
        FpOps(FIELDS["bls12377_fp"]),
        FIELDS["bls12377_fr"],
        1,
        (
            81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,
            241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030,
        ),
    )
    # ecc/bls12-377/bls12-377.go:10,103-105,111-114 : twist Y^2 = X^3 + 1/u, u^2 = -5
    K2 = Fp2Ops(FIELDS["bls12377_fp"], nonres=-5)
    g["bls12377_g2"] = Group(
        "bls12377_g2",
        K2,
        FIELDS["bls12377_fr"],
        K2.inv((0, 1)),
        (
            (
                233578398248691099356572568220835526895379068987715365179118596935057653620464273615301663571204657964920925606294,
                140913150380207355837477652521042157274541796891053068589147167627541651775299824604154852141315666357241556069118,
            ),
            (
                63160294768292073209381361943935198908131692476676907196754037919244929611450776219210369229519898517858833747423,
                149157405641012693445398062341192467754805999074082136895788947234480009303640899064710353187729182149407503257491,
            ),
        ),
    )
    # N4 remainder.  ecc/secp256k1/secp256k1.go:48-55 : Y^2 = X^3 + 7, the SEC2 generator
    g["secp256k1_g1"] = Group(
        "secp256k1_g1",
        FpOps(FIELDS["secp256k1_fp"]),
        FIELDS["secp256k1_fr"],
        7,
        (
            55066263022277343669578718895168534326250603453777594175500187360389116729240,
            32670510020758816978083085130507043184471273380659243275938904335757337482424,
        ),
    )
    # ecc/bw6-761/bw6-761.go:92-103 : G1 Y^2 = X^3 - 1 and the M-twist G2 Y^2 = X^3 + 4, BOTH over Fp (12 words); fr 377 bits
    K = FpOps(FIELDS["bw6761_fp"])
    g["bw6761_g1"] = Group(
        "bw6761_g1",
        K,
        FIELDS["bw6761_fr"],
        K.neg(1),
        (
            6238772257594679368032145693622812838779005809760824733138787810501188623461307351759238099287535516224314149266511977132140828635950940021790489507611754366317801811090811367945064510304504157188661901055903167026722666149426237,
            2101735126520897423911504562215834951148127555913367997162789335052900271653517958562461315794228241561913734371411178226936527683203879553093934185950470971848972085321797958124416462268292467002957525517188485984766314758624099,
        ),
    )
    g["bw6761_g2"] = Group(
        "bw6761_g2",
        K,
        FIELDS["bw6761_fr"],
        4,
        (
            6445332910596979336035888152774071626898886139774101364933948236926875073754470830732273879639675437155036544153105017729592600560631678554299562762294743927912429096636156401171909259073181112518725201388196280039960074422214428,
            562923658089539719386922163444547387757586534741080263946953401595155211934630598999300396317104182598044793758153214972605680357108252243146746187917218885078195819486220416605630144001533548163105316661692978285266378674355041,
        ),
    )
    # ecc/bls24-315/bls24-315.go:105-113 (Y^2 = X^3 + 1), ecc/bls24-317/bls24-317.go:91-99 (Y^2 = X^3 + 4): G1 only -- G2 of the
    # bls24 curves is defined over Fp4 and is not on this path
    g["bls24315_g1"] = Group("bls24315_g1", FpOps(FIELDS["bls24315_fp"]), FIELDS["bls24315_fr"], 1, (
        34223510504517033132712852754388476272837911830964394866541204856091481856889569724484362330263,
        24215295174889464585413596429561903295150472552154479431771837786124301185073987899223459122783))
    g["bls24317_g1"] = Group("bls24317_g1", FpOps(FIELDS["bls24317_fp"]), FIELDS["bls24317_fr"], 4, (
        26261810162995192444253184251590159762050205376519976412461726336843100448942248976252388876791,
        26146603602820658047261036676090398397874822703333117264049387703172159980214065566219085800243))
    # ecc/bw6-633/bw6-633.go:82-93 : G1 Y^2 = X^3 + 4, M-twist G2 Y^2 = X^3 + 8, both over the 10-word Fp; fr = 5 words, 315 bits
    K = FpOps(FIELDS["bw6633_fp"])
    g["bw6633_g1"] = Group("bw6633_g1", K, FIELDS["bw6633_fr"], 4, (
        14087405796052437206213362229855313116771222912153372774869400386285407949123477431442535997951698710614498307938219633856996133201713506830167161540335446217605918678317160130862890417553415,
        5208886161111258314476333487866604447704068601830026647530443033297117148121067806438008469463787158470000157308702133756065259580313172904438248825389121766442385979570644351664733475122746))
    g["bw6633_g2"] = Group("bw6633_g2", K, FIELDS["bw6633_fr"], 8, (
        13658793733252505713431834233072715040674666715141692574468286839081203251180283741830175712695426047062165811313478642863696265647598838732554425602399576125615559121457137320131899043374497,
        599560264833409786573595720823495699033661029721475252751314180543773745554433461106678360045466656230822473390866244089461950086268801746497554519984580043036179195728559548424763890207250))
    return g


GROUPS = _mk_groups()

# implementedCs of each curve's MultiExp (multiexp.go:77): what bestC may return
IMPLEMENTED_CS = {"secp256k1_g1": tuple(range(4, 16)), "bw6761_g1": (4, 5, 8, 10, 16), "bw6761_g2": (4, 5, 8, 10, 16),
                  "bls24315_g1": tuple(range(4, 17)), "bls24317_g1": tuple(range(4, 17)), "bw6633_g1": (4, 5, 6, 8, 12, 16),
                  "bw6633_g2": (4, 5, 6, 8, 12, 16)}


# --------------------------------------------------------------------------------------
# MSM
# --------------------------------------------------------------------------------------


def compute_nb_chunks(bits: int, c: int) -> int:
    """computeNbChunks multiexp.go:681-683"""
    return (bits + c - 1) // c


def last_c(bits: int, c: int) -> int:
    """lastC multiexp.go:690-693"""
    nb_available = compute_nb_chunks(bits, c) * c - bits
    return c + 1 - nb_available


def best_c(bits: int, n: int, cs=range(4, 17)) -> int:
    """bestC multiexp.go:75-93"""
    best, bc = None, None
    for c in cs:
        cost = float((bits + 1) * (n + (1 << c))) / float(c)
        if best is None or cost < best:
            best, bc = cost, c
    return bc


def partition_scalars(fr: Field, scalars_mont, c: int) -> np.ndarray:
    """partitionScalars multiexp.go:709-803.

    scalars_mont: iterable of Montgomery-form integers (< r) as Go holds them.
    Returns digits[chunk][i] as uint32 (the reference stores uint16 since its c <= 16;
    the encoding is the same: 0 skip, 2d for d>0, 2(-d-1)+1 for d<0; last chunk 2d).
    """
    scalars_mont = list(scalars_mont)
    n = len(scalars_mont)
    W = compute_nb_chunks(fr.bits, c)
    digits = np.zeros((W, n), dtype=np.uint32)
    mask = (1 << c) - 1
    mx = (1 << (c - 1)) - 1
    limbs = fr.limbs
    for i, s in enumerate(scalars_mont):
        if s == 0:  # IsZero() on the Montgomery limbs, multiexp.go:743
            continue
        k = fr.from_mont(s)  # Bits(), fr/element.go:855-859
        kl = fr.to_limbs(k)
        carry = 0
        for ch in range(W):
            jc = ch * c
            idx = jc // 64
            shift = jc - idx * 64
            # selector logic multiexp.go:729-737 (bits beyond the last limb read as absent)
            d = carry + (((kl[idx] & ((mask << shift) & 0xFFFFFFFFFFFFFFFF)) >> shift) if idx < limbs else 0)
            multi = (64 % c != 0) and shift > (64 - c) and idx < limbs - 1
            if multi:
                nb_hi = shift - (64 - c)
                d += (kl[idx + 1] & ((1 << nb_hi) - 1)) << (c - nb_hi)
            if ch < W - 1:
                carry = 0
                if d > mx:
                    d -= 1 << c
                    carry = 1
                if d == 0:
                    continue
                digits[ch, i] = (d << 1) if d > 0 else (((-d - 1) << 1) + 1)
            else:
                digits[ch, i] = d << 1  # multiexp.go:788-800
    return digits


def process_chunk(G: Group, c_eff: int, points, digits_row):
    """processChunkG1Jacobian multiexp_jacobian.go:8-61 -> window total (xyzz)."""
    nb = 1 << (c_eff - 1)
    buckets = [G.xyzz_inf() for _ in range(nb)]
    for i, e in enumerate(digits_row):
        e = int(e)
        if e == 0:
            continue
        if e & 1 == 0:
            buckets[(e >> 1) - 1] = G.add_mixed(buckets[(e >> 1) - 1], points[i])
        else:
            buckets[e >> 1] = G.add_mixed(buckets[e >> 1], points[i], negate=True)
    run, tot = G.xyzz_inf(), G.xyzz_inf()
    for k in range(nb - 1, -1, -1):
        if not G.K.is_zero(buckets[k][2]):
            run = G.xyzz_add(run, buckets[k])
        tot = G.xyzz_add(tot, run)
    return tot


def reduce_chunks(G: Group, c: int, totals):
    """msmReduceChunkG1Affine multiexp.go:302-315 -> Jacobian triple."""
    acc = list(totals[-1])
    for j in range(len(totals) - 2, -1, -1):
        for _ in range(c):
            acc = G.xyzz_double(acc)
        acc = G.xyzz_add(acc, totals[j])
    return G.xyzz_to_jac(acc)


def inner_msm(G: Group, c: int, points, scalars_mont):
    """_innerMsmG1 multiexp.go:148-209 (sequential restatement). Returns Jacobian triple."""
    fr = G.fr
    W = compute_nb_chunks(fr.bits, c)
    digits = partition_scalars(fr, scalars_mont, c)
    totals = []
    for j in range(W):
        ce = last_c(fr.bits, c) if j == W - 1 else c
        totals.append(process_chunk(G, ce, points, digits[j]))
    return reduce_chunks(G, c, totals)


def multi_exp(G: Group, points, scalars_mont, c=None):
    """(*G1Jac).MultiExp multiexp.go:32-146 without the (result-preserving) split. -> affine."""
    if len(points) != len(scalars_mont):
        raise ValueError("len(points) != len(scalars)")
    if c is None:
        c = best_c(G.fr.bits, len(points))
    return G.jac_to_affine(inner_msm(G, c, points, scalars_mont))


def multi_exp_tables(G: Group, points, scalars_mont, c: int):
    """The window-table restatement the GPU engine uses for resident bases (gmsm_bases_precompute; no counterpart in
    the reference, which re-reads its bases per call): table row j = 2^(c*j) * points, so that
    sum_i s_i P_i = sum_j sum_i d_ij (2^(c*j) P_i) is ONE bucket problem over the digits of partitionScalars
    (multiexp.go:709-803) with max(2^(c-1), 2^(lastC-1)) shared buckets -- no Horner (multiexp.go:302-315).
    Its agreement with multi_exp() is what tests/test_oracle.py checks. -> affine."""
    fr = G.fr
    W = compute_nb_chunks(fr.bits, c)
    digits = partition_scalars(fr, scalars_mont, c)
    rows = [list(points)]
    for j in range(1, W):
        rows.append([G.aff_inf() if G.aff_is_inf(p) else G.scalar_mul(p, 1 << c) for p in rows[-1]])
    nb = max(1 << (c - 1), 1 << (last_c(fr.bits, c) - 1))
    buckets = [G.xyzz_inf() for _ in range(nb)]
    for j in range(W):
        for i, e in enumerate(digits[j]):
            e = int(e)
            if e == 0:
                continue
            if e & 1 == 0:
                buckets[(e >> 1) - 1] = G.add_mixed(buckets[(e >> 1) - 1], rows[j][i])
            else:
                buckets[e >> 1] = G.add_mixed(buckets[e >> 1], rows[j][i], negate=True)
    run, tot = G.xyzz_inf(), G.xyzz_inf()
    for k in range(nb - 1, -1, -1):
        if not G.K.is_zero(buckets[k][2]):
            run = G.xyzz_add(run, buckets[k])
        tot = G.xyzz_add(tot, run)
    return G.jac_to_affine(G.xyzz_to_jac(tot))


def msm_naive(G: Group, points, ks):
    """sum k_i * P_i by independent double-and-add (independent of the bucket method)."""
    acc = G.aff_inf()
    for p, k in zip(points, ks):
        acc = G.aff_add(acc, G.scalar_mul(p, k))
    return acc


# --------------------------------------------------------------------------------------
# deterministic synthetic inputs
# --------------------------------------------------------------------------------------

M64 = 0xFFFFFFFFFFFFFFFF


def splitmix64(state: int):
    state = (state + 0x9E3779B97F4A7C15) & M64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return state, z ^ (z >> 31)


def random_scalars_mont(fr: Field, n: int, seed: int):
    """n uniform values < r, 4 x u64 with top bits masked to fr.bits and rejection-sampled
    (analogue of SetRandom fr/element.go:302-343); the limbs are read AS the Montgomery
    representation.  Scalar i depends only on (seed, i): state = seed + i*2^32 stream."""
    out = []
    topmask = (1 << (fr.bits - 64 * (fr.limbs - 1))) - 1
    for i in range(n):
        st = (seed ^ (i * 0xD1342543DE82EF95)) & M64
        while True:
            limbs = []
            for _ in range(fr.limbs):
                st, v = splitmix64(st)
                limbs.append(v)
            limbs[-1] &= topmask
            v = Field.from_limbs(limbs)
            if v < fr.q:
                out.append(v)
                break
    return out


def consecutive_multiples(G: Group, n: int, start_k: int = 1, base=None):
    """[start_k]B, [start_k+1]B, ... as in TestCrossMultiExpG1 (multiexp_test.go:224-230)."""
    base = base if base is not None else G.gen
    p = G.scalar_mul(base, start_k)
    out = []
    for _ in range(n):
        out.append(p)
        p = G.aff_add(p, base)
    return out


# --------------------------------------------------------------------------------------
# Next-row N3: Fr FFT (ecc/bn254/fr/fft/domain.go:67-110, fft.go:31-190, bitreverse.go:17-42;
# ecc/bls12-381/fr/fft/ identical with its own constants).  Plain integers mod r.
# --------------------------------------------------------------------------------------

# fr.Generator (ecc/bn254/fr/generator.go:18-36, ecc/bls12-381/fr/generator.go) and
# GeneratorFullMultiplicativeGroup (fft/domain.go:55-63)
FFT_PARAMS = {
    "bn254_fr": dict(root=19103219067921713944291392827692070036145651957329286315305642004821462161904, max_order=28, mult_gen=5),
    "bls12381_fr": dict(root=10238227357739495823651030575849232062558860180284477541189508159991286009131, max_order=32, mult_gen=7),
    # ecc/bls12-377/fr/generator.go:23-24, fr/fft/domain.go:59
    "bls12377_fr": dict(root=8065159656716812877374967518403273466521432693661810619979959746626482506078, max_order=47, mult_gen=22),
}
DIT, DIF = 0, 1  # fft.Decimation (fft.go:18-23)


def bit_reverse(v):
    """BitReverse (bitreverse.go:17-42): in-place bit-reversal permutation, returns the list"""
    n = len(v)
    assert n & (n - 1) == 0
    lg = n.bit_length() - 1
    for i in range(n):
        r = int(bin(i)[2:].zfill(lg)[::-1], 2) if lg else 0
        if r > i:
            v[i], v[r] = v[r], v[i]
    return v


class FFTDomain:
    """fft.Domain / NewDomain (domain.go:24-110)"""

    def __init__(self, frname: str, m: int, shift: int = None):
        f = FIELDS[frname]
        P = FFT_PARAMS[frname]
        self.q = f.q
        x = 1
        while x < m:
            x <<= 1
        self.cardinality = x
        logx = x.bit_length() - 1
        if logx > P["max_order"]:
            raise ValueError("m (%d) is too big: the required root of unity does not exist" % m)
        self.generator = pow(P["root"], 1 << (P["max_order"] - logx), self.q)
        self.generator_inv = pow(self.generator, -1, self.q)
        self.cardinality_inv = pow(x, -1, self.q)
        self.shift = P["mult_gen"] if shift is None else shift % self.q
        self.shift_inv = pow(self.shift, -1, self.q)

    def _dif(self, a, w):
        # difFFT (fft.go:195-260): natural order in, bit-reversed order out
        q = self.q
        n = len(a)
        m = n >> 1
        while m >= 1:
            wm = pow(w, n // (2 * m), q)
            for start in range(0, n, 2 * m):
                t = 1
                for j in range(m):
                    x, y = a[start + j], a[start + j + m]
                    a[start + j] = (x + y) % q
                    a[start + j + m] = (x - y) * t % q
                    t = t * wm % q
            m >>= 1
        return a

    def _dit(self, a, w):
        # ditFFT (fft.go:262+): bit-reversed order in, natural order out
        q = self.q
        n = len(a)
        m = 1
        while m < n:
            wm = pow(w, n // (2 * m), q)
            for start in range(0, n, 2 * m):
                t = 1
                for j in range(m):
                    x, y = a[start + j], a[start + j + m] * t % q
                    a[start + j] = (x + y) % q
                    a[start + j + m] = (x - y) % q
                    t = t * wm % q
            m <<= 1
        return a

    def _rev(self, i):
        lg = self.cardinality.bit_length() - 1
        return int(bin(i)[2:].zfill(lg)[::-1], 2) if lg else 0

    def fft(self, a, decimation, coset=False):
        """Domain.FFT (fft.go:31-110)"""
        a = [v % self.q for v in a]
        assert len(a) == self.cardinality
        if coset:
            for i in range(len(a)):
                e = self._rev(i) if decimation == DIT else i
                a[i] = a[i] * pow(self.shift, e, self.q) % self.q
        return self._dif(a, self.generator) if decimation == DIF else self._dit(a, self.generator)

    def fft_inverse(self, a, decimation, coset=False):
        """Domain.FFTInverse (fft.go:117-190)"""
        a = [v % self.q for v in a]
        assert len(a) == self.cardinality
        a = self._dif(a, self.generator_inv) if decimation == DIF else self._dit(a, self.generator_inv)
        for i in range(len(a)):
            s = self.cardinality_inv
            if coset:
                e = i if decimation == DIT else self._rev(i)
                s = s * pow(self.shift_inv, e, self.q) % self.q
            a[i] = a[i] * s % self.q
        return a
