// Short-Weierstrass (a = 0) point arithmetic over a coordinate field F (Fp<P> for G1, Fp2<P> for G2).
//
// Replaces (reference, /root/reference/ecc/bn254, same in ecc/bls12-381):
//   G1Affine{X,Y}, infinity = (0,0)                 g1.go:18-20,41-47,178-180
//   g1JacExtended{X,Y,ZZ,ZZZ}, infinity iff ZZ==0   g1.go:28-30,688-699
//   addMixed / subMixed (madd-2008-s)               g1.go:822-873, 878-930
//   doubleMixed / doubleNegMixed                    g1.go:962-985, 933-957
//   add (add-2008-s), double (dbl-2008-s-1)         g1.go:736-788, 795-817
//   unsafeFromJacExtended, FromJacobian             g1.go:726-731, 150-166
//   G2 twins over E2                                g2.go:663-970
#pragma once
#include "field.cuh"
#include "fp2.cuh"

namespace gmsm {

template <class F>
struct Affine {
  F x, y;
  GMSM_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  GMSM_HD static Affine inf() { return Affine{F::zero(), F::zero()}; }
};

template <class F>
struct XYZZ {
  F x, y, zz, zzz;
  GMSM_HD bool is_inf() const { return zz.is_zero(); }
  // any (X, Y, 0, 0) is infinity; we use all-zero so that cudaMemset(0) initialises bucket arrays
  GMSM_HD static XYZZ inf() { return XYZZ{F::zero(), F::zero(), F::zero(), F::zero()}; }
};

template <class F>
struct Jac {
  F x, y, z;
};

template <class F>
GMSM_HD XYZZ<F> xyzz_from_affine(const Affine<F>& a) {
  if (a.is_inf()) return XYZZ<F>::inf();
  return XYZZ<F>{a.x, a.y, F::one(), F::one()};
}

// [2](x, y) with y already sign-adjusted: doubleMixed g1.go:962-985 (doubleNegMixed :933-957 is the
// same computation on (x, -y)).
template <class F>
GMSM_HD XYZZ<F> xyzz_double_affine(const F& x, const F& y) {
  F U = f_dbl(y);
  F V = f_sqr(U);
  F W = f_mul(U, V);
  F S = f_mul(x, V);
  F XX = f_sqr(x);
  F M = f_add(f_dbl(XX), XX);
  F S2 = f_dbl(S);
#if defined(GMSM_DOT2)
  XYZZ<F> r;
  r.x = f_sub(f_sqr(M), S2);
  r.y = f_dot2(f_sub(S, r.x), M, W, f_neg(y));   // (S - X3) M - W y with one reduction (experimental, field.cuh)
#else
  F L = f_mul(W, y);
  XYZZ<F> r;
  r.x = f_sub(f_sqr(M), S2);
  r.y = f_sub(f_mul(f_sub(S, r.x), M), L);
#endif
  r.zz = V;
  r.zzz = W;
  return r;
}

// p += a  (negate == false)  or  p -= a  (negate == true)
template <class F>
GMSM_HD void xyzz_add_mixed(XYZZ<F>& p, const Affine<F>& a, bool negate) {
  if (a.is_inf()) return;  // g1.go:825
  F ay = negate ? f_neg(a.y) : a.y;
  if (p.zz.is_zero()) {  // g1.go:829-835
    p.x = a.x;
    p.y = ay;
    p.zz = F::one();
    p.zzz = F::one();
    return;
  }
  F P = f_sub(f_mul(a.x, p.zz), p.x);
  F R = f_sub(f_mul(ay, p.zzz), p.y);
  if (P.is_zero()) {  // g1.go:846-854
    if (R.is_zero()) {
      p = xyzz_double_affine(a.x, ay);
    } else {
      p.zz = F::zero();
      p.zzz = F::zero();
    }
    return;
  }
  F PP = f_sqr(P);
  F PPP = f_mul(P, PP);
  F Q = f_mul(p.x, PP);
  F RR = f_sqr(R);
  F X3 = f_sub(f_sub(RR, PPP), f_dbl(Q));
#if defined(GMSM_DOT2)
  F Y3 = f_dot2(f_sub(Q, X3), R, f_neg(p.y), PPP);   // (Q - X3) R - Y1 PPP with one reduction (experimental, field.cuh)
#else
  F Y3 = f_sub(f_mul(f_sub(Q, X3), R), f_mul(p.y, PPP));
#endif
  p.x = X3;
  p.y = Y3;
  p.zz = f_mul(p.zz, PP);
  p.zzz = f_mul(p.zzz, PPP);
}

// double g1.go:795-817 (valid for infinity)
template <class F>
GMSM_HD XYZZ<F> xyzz_double(const XYZZ<F>& q) {
  F U = f_dbl(q.y);
  F V = f_sqr(U);
  F W = f_mul(U, V);
  F S = f_mul(q.x, V);
  F XX = f_sqr(q.x);
  F M = f_add(f_dbl(XX), XX);
  F U2 = f_mul(W, q.y);
  XYZZ<F> r;
  r.x = f_sub(f_sub(f_sqr(M), S), S);
  r.y = f_sub(f_mul(f_sub(S, r.x), M), U2);
  r.zz = f_mul(V, q.zz);
  r.zzz = f_mul(W, q.zzz);
  return r;
}

// p += q   g1.go:736-788
template <class F>
GMSM_HD void xyzz_add(XYZZ<F>& p, const XYZZ<F>& q) {
  if (q.zz.is_zero()) return;
  if (p.zz.is_zero()) {
    p = q;
    return;
  }
  F U2 = f_mul(q.x, p.zz);
  F U1 = f_mul(p.x, q.zz);
  F S2 = f_mul(q.y, p.zzz);
  F S1 = f_mul(p.y, q.zzz);
  F P = f_sub(U2, U1);
  F R = f_sub(S2, S1);
  if (P.is_zero()) {
    if (R.is_zero()) {
      p = xyzz_double(q);
    } else {
      p.zz = F::zero();
      p.zzz = F::zero();
    }
    return;
  }
  F PP = f_sqr(P);
  F PPP = f_mul(P, PP);
  F Q = f_mul(U1, PP);
#if defined(GMSM_DOT2)
  F X3 = f_sub(f_sub(f_sub(f_sqr(R), PPP), Q), Q);
  F Y3 = f_dot2(f_sub(Q, X3), R, f_neg(S1), PPP);
#else
  F V = f_mul(S1, PPP);
  F X3 = f_sub(f_sub(f_sub(f_sqr(R), PPP), Q), Q);
  F Y3 = f_sub(f_mul(f_sub(Q, X3), R), V);
#endif
  p.x = X3;
  p.y = Y3;
  p.zz = f_mul(f_mul(p.zz, q.zz), PP);
  p.zzz = f_mul(f_mul(p.zzz, q.zzz), PPP);
}

// Jacobian doubling for a = 0 (dbl-2009-l: 2M + 5S; the reference's G1Jac.DoubleAssign, g1.go:396-424, is the
// same point in the same coordinate system).  Valid for infinity (Z = 0 stays 0).  Used by the serial Horner
// of k_finalize, where it replaces the 6M + 3S extended-Jacobian doubling.
template <class F>
GMSM_HD Jac<F> jac_double(const Jac<F>& p) {
  F A = f_sqr(p.x);
  F B = f_sqr(p.y);
  F C = f_sqr(B);
  F D = f_sub(f_sub(f_sqr(f_add(p.x, B)), A), C);
  D = f_dbl(D);
  F E = f_add(f_dbl(A), A);
  F Fq = f_sqr(E);
  Jac<F> r;
  r.z = f_dbl(f_mul(p.y, p.z));
  r.x = f_sub(Fq, f_dbl(D));
  F C8 = f_dbl(f_dbl(f_dbl(C)));
  r.y = f_sub(f_mul(E, f_sub(D, r.x)), C8);
  return r;
}

// One batch of a window-table level (k_table_level, kernels.cuh): out_i = 2^c * in_i for cnt <= TAB_M affine points,
// c Jacobian doublings each and ONE shared inversion (Montgomery's trick over the non-zero Z's).  Infinity -- on
// input, or reached by the doublings (a point of 2-power order) -- stays (0, 0).  ld(i) / st(i, a) / dbl(j) are the
// caller's load, store and doubling (the kernel passes 256-bit loads and an out-of-line doubling; the CPU formula
// check of tests/test_hostcheck.py passes plain ones).
static constexpr int TAB_M = 8;
template <class F, class Ld, class St, class Dbl>
GMSM_HD void table_level_batch(int cnt, int c, Ld ld, St st, Dbl dbl) {
  Jac<F> pts[TAB_M];
  F pref[TAB_M];
  F prod = F::one();
  for (int i = 0; i < cnt; i++) {
    const Affine<F> a = ld(i);
    Jac<F> j = a.is_inf() ? Jac<F>{F::zero(), F::zero(), F::zero()} : Jac<F>{a.x, a.y, F::one()};
    for (int l = 0; l < c; l++) j = dbl(j);
    pts[i] = j;
    pref[i] = prod;   // product of the non-zero Z's before i
    if (!j.z.is_zero()) prod = f_mul(prod, j.z);
  }
  F inv = f_inv(prod);
  for (int i = cnt - 1; i >= 0; i--) {
    Affine<F> a = Affine<F>::inf();
    if (!pts[i].z.is_zero()) {
      const F zi = f_mul(inv, pref[i]);   // 1 / Z_i
      inv = f_mul(inv, pts[i].z);
      const F z2 = f_sqr(zi);
      a.x = f_mul(pts[i].x, z2);
      a.y = f_mul(f_mul(pts[i].y, z2), zi);
    }
    st(i, a);
  }
}

// Jacobian (X, Y, Z) -> extended Jacobian (X, Y, Z^2, Z^3): same X, Y
template <class F>
GMSM_HD XYZZ<F> jac_to_xyzz(const Jac<F>& p) {
  XYZZ<F> r;
  r.x = p.x;
  r.y = p.y;
  r.zz = f_sqr(p.z);
  r.zzz = f_mul(r.zz, p.z);
  return r;
}

// unsafeFromJacExtended g1.go:726-731; infinity (ZZ = ZZZ = 0) maps to (0,0,0)
template <class F>
GMSM_HD Jac<F> xyzz_to_jac(const XYZZ<F>& p) {
  if (p.zz.is_zero()) return Jac<F>{F::zero(), F::zero(), F::zero()};
  Jac<F> j;
  j.x = f_mul(f_sqr(p.zz), p.x);
  j.y = f_mul(f_sqr(p.zzz), p.y);
  j.z = p.zzz;
  return j;
}

// FromJacobian g1.go:150-166
template <class F>
GMSM_HD Affine<F> jac_to_affine(const Jac<F>& j) {
  if (j.z.is_zero()) return Affine<F>::inf();
  F a = f_inv(j.z);
  F b = f_sqr(a);
  Affine<F> r;
  r.x = f_mul(j.x, b);
  r.y = f_mul(f_mul(j.y, b), a);
  return r;
}

}  // namespace gmsm
