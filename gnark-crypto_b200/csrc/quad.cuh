// Lane-parallel group law for the latency-bound tail of an MSM (carry join, bucket reduction, Horner over the windows).
//
// Those stages run a few thousand threads, each through a serial chain of 50 .. 250 point operations (the reference does the
// same chains on one goroutine: the running sum of multiexp_jacobian.go:44-52, the Horner of msmReduceChunk,
// multiexp.go:302-315).  A lone thread issues one dependent carry-chain instruction every ~6 cycles, so one Montgomery
// product costs ~1000 cycles of latency and a 14-product addition ~7 us -- the GPU idles while the chain ticks.  Here a QUAD
// (four adjacent lanes) executes ONE point operation: every lane holds the same operands and the same results (so every
// branch on a field value is quad-uniform by construction), and the independent products of a formula step are spread over
// the lanes -- each lane selects its operand pair, multiplies, and the products are broadcast back with shuffles:
//   full addition   (add-2008-s,  g1.go:736-788): 14 products in 4 steps
//   doubling        (dbl-2008-s-1, g1.go:795-817):  9 products in 3 steps
//   Jacobian double (g1.go:396-424 / dbl-2009-l):   7 products in 3 steps
// For the Fp2 groups the parallelism is taken one level down instead (register budget): the three base-field products of
// a Karatsuba Fp2 product (e2_bn254.go:28-38) go to three lanes, the two of a squaring to two.
// Values are the same unique reduced field elements whichever lane computes them, so results are bit-identical to curve.cuh.
#pragma once
#include "curve.cuh"

namespace gmsm {

struct Quad {
  int ql;          // lane within the quad, 0..3
  unsigned mask;   // the quad's four lanes within the warp
};

GMSM_D Quad quad_of_thread() {
#if defined(__CUDA_ARCH__)
  const unsigned lane = threadIdx.x & 31u;
#else
  const unsigned lane = threadIdx.x & 31u;
#endif
  return Quad{(int)(lane & 3u), 0xFu << (lane & ~3u)};
}

template <class P>
GMSM_D Fp<P> quad_pick(const Quad& q, const Fp<P>& a0, const Fp<P>& a1, const Fp<P>& a2, const Fp<P>& a3) {
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < P::N; i++) r.l[i] = (q.ql == 0) ? a0.l[i] : (q.ql == 1) ? a1.l[i] : (q.ql == 2) ? a2.l[i] : a3.l[i];
  return r;
}
template <class P>
GMSM_D Fp<P> quad_bcast(const Quad& q, const Fp<P>& v, int src) {
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < P::N; i++) r.l[i] = __shfl_sync(q.mask, v.l[i], src, 4);
  return r;
}

// r_k = a_k * b_k for k < 4, one product per lane
template <class P>
GMSM_D void quad_mul4(const Quad& q, const Fp<P>& a0, const Fp<P>& b0, const Fp<P>& a1, const Fp<P>& b1, const Fp<P>& a2, const Fp<P>& b2,
                      const Fp<P>& a3, const Fp<P>& b3, Fp<P>& r0, Fp<P>& r1, Fp<P>& r2, Fp<P>& r3) {
  const Fp<P> t = fp_mul(quad_pick(q, a0, a1, a2, a3), quad_pick(q, b0, b1, b2, b3));
  r0 = quad_bcast(q, t, 0);
  r1 = quad_bcast(q, t, 1);
  r2 = quad_bcast(q, t, 2);
  r3 = quad_bcast(q, t, 3);
}
template <class P>
GMSM_D void quad_mul3(const Quad& q, const Fp<P>& a0, const Fp<P>& b0, const Fp<P>& a1, const Fp<P>& b1, const Fp<P>& a2, const Fp<P>& b2,
                      Fp<P>& r0, Fp<P>& r1, Fp<P>& r2) {
  const Fp<P> t = fp_mul(quad_pick(q, a0, a1, a2, a2), quad_pick(q, b0, b1, b2, b2));
  r0 = quad_bcast(q, t, 0);
  r1 = quad_bcast(q, t, 1);
  r2 = quad_bcast(q, t, 2);
}
template <class P>
GMSM_D void quad_mul2(const Quad& q, const Fp<P>& a0, const Fp<P>& b0, const Fp<P>& a1, const Fp<P>& b1, Fp<P>& r0, Fp<P>& r1) {
  const Fp<P> t = fp_mul(quad_pick(q, a0, a1, a0, a1), quad_pick(q, b0, b1, b0, b1));
  r0 = quad_bcast(q, t, 0);
  r1 = quad_bcast(q, t, 1);
}

// ---- Fp2: field-level parallelism (same formulas as fp2.cuh's f_mul / f_sqr) ----
template <class P>
GMSM_D Fp2<P> quad_fmul(const Quad& q, const Fp2<P>& x, const Fp2<P>& y) {
  Fp<P> a, b, c;
  quad_mul3(q, fp_add(x.a0, x.a1), fp_add(y.a0, y.a1), x.a0, y.a0, x.a1, y.a1, a, b, c);
  Fp2<P> z;
  z.a1 = fp_sub(fp_sub(a, b), c);
  if (P::FP2_NONRES == -5) c = fp_mul_by5(c);
  z.a0 = fp_sub(b, c);
  return z;
}
template <class P>
GMSM_D Fp2<P> quad_fsqr(const Quad& q, const Fp2<P>& x) {
  Fp<P> a, b;
  quad_mul2(q, fp_add(x.a0, x.a1), (P::FP2_NONRES == -5) ? fp_sub(x.a0, fp_mul_by5(x.a1)) : fp_sub(x.a0, x.a1), x.a0, x.a1, a, b);
  b = fp_dbl(b);
  if (P::FP2_NONRES == -5) a = fp_add(a, fp_dbl(b));
  return Fp2<P>{a, b};
}

// ------------------------------------------------------------------------------------------
// point operations, Fp coordinates: formula-level parallelism
// ------------------------------------------------------------------------------------------
// double g1.go:795-817 (valid for infinity)
template <class P>
GMSM_D XYZZ<Fp<P>> xyzz_double_quad(const Quad& q, const XYZZ<Fp<P>>& a) {
  using F = Fp<P>;
  const F U = f_dbl(a.y);
  F V, XX;
  quad_mul2(q, U, U, a.x, a.x, V, XX);
  const F M = f_add(f_dbl(XX), XX);
  F W, S, ZZ3, MM;
  quad_mul4(q, U, V, a.x, V, V, a.zz, M, M, W, S, ZZ3, MM);
  XYZZ<F> r;
  r.x = f_sub(f_sub(MM, S), S);
  F ZZZ3, WY, T;
  quad_mul3(q, W, a.zzz, W, a.y, f_sub(S, r.x), M, ZZZ3, WY, T);
  r.y = f_sub(T, WY);
  r.zz = ZZ3;
  r.zzz = ZZZ3;
  return r;
}

// p += a   g1.go:736-788
template <class P>
GMSM_D void xyzz_add_quad(const Quad& q, XYZZ<Fp<P>>& p, const XYZZ<Fp<P>>& a) {
  using F = Fp<P>;
  if (a.zz.is_zero()) return;
  if (p.zz.is_zero()) {
    p = a;
    return;
  }
  F U2, U1, S2, S1;
  quad_mul4(q, a.x, p.zz, p.x, a.zz, a.y, p.zzz, p.y, a.zzz, U2, U1, S2, S1);
  const F Pp = f_sub(U2, U1);
  const F R = f_sub(S2, S1);
  if (Pp.is_zero()) {
    if (R.is_zero()) {
      p = xyzz_double_quad(q, a);
    } else {
      p.zz = F::zero();
      p.zzz = F::zero();
    }
    return;
  }
  F PP, RR, Z2, Z3;
  quad_mul4(q, Pp, Pp, R, R, p.zz, a.zz, p.zzz, a.zzz, PP, RR, Z2, Z3);
  F PPP, Q, ZZ3;
  quad_mul3(q, Pp, PP, U1, PP, Z2, PP, PPP, Q, ZZ3);
  const F X3 = f_sub(f_sub(f_sub(RR, PPP), Q), Q);
  F V, ZZZ3, T;
  quad_mul3(q, S1, PPP, Z3, PPP, f_sub(Q, X3), R, V, ZZZ3, T);
  p.x = X3;
  p.y = f_sub(T, V);
  p.zz = ZZ3;
  p.zzz = ZZZ3;
}

// Jacobian doubling (jac_double of curve.cuh, dbl-2009-l), valid for infinity
template <class P>
GMSM_D Jac<Fp<P>> jac_double_quad(const Quad& q, const Jac<Fp<P>>& p) {
  using F = Fp<P>;
  F A, B, YZ;
  quad_mul3(q, p.x, p.x, p.y, p.y, p.y, p.z, A, B, YZ);
  const F E = f_add(f_dbl(A), A);
  const F XB = f_add(p.x, B);
  F C, T, Fq;
  quad_mul3(q, B, B, XB, XB, E, E, C, T, Fq);
  const F D = f_dbl(f_sub(f_sub(T, A), C));
  Jac<F> r;
  r.z = f_dbl(YZ);
  r.x = f_sub(Fq, f_dbl(D));
  const F C8 = f_dbl(f_dbl(f_dbl(C)));
  r.y = f_sub(fp_mul(E, f_sub(D, r.x)), C8);   // a single product: every lane computes it, no exchange
  return r;
}
template <class P>
GMSM_D XYZZ<Fp<P>> jac_to_xyzz_quad(const Quad&, const Jac<Fp<P>>& p) {
  return jac_to_xyzz(p);   // two dependent products: nothing to spread
}
template <class P>
GMSM_D Jac<Fp<P>> xyzz_to_jac_quad(const Quad& q, const XYZZ<Fp<P>>& p) {
  using F = Fp<P>;
  if (p.zz.is_zero()) return Jac<F>{F::zero(), F::zero(), F::zero()};
  F z2, z3;
  quad_mul2(q, p.zz, p.zz, p.zzz, p.zzz, z2, z3);
  Jac<F> j;
  quad_mul2(q, z2, p.x, z3, p.y, j.x, j.y);
  j.z = p.zzz;
  return j;
}

// ------------------------------------------------------------------------------------------
// point operations, Fp2 coordinates: the formulas of curve.cuh with lane-parallel Fp2 products
// ------------------------------------------------------------------------------------------
template <class P>
GMSM_D XYZZ<Fp2<P>> xyzz_double_quad(const Quad& q, const XYZZ<Fp2<P>>& a) {
  using F = Fp2<P>;
  const F U = f_dbl(a.y);
  const F V = quad_fsqr(q, U);
  const F W = quad_fmul(q, U, V);
  const F S = quad_fmul(q, a.x, V);
  const F XX = quad_fsqr(q, a.x);
  const F M = f_add(f_dbl(XX), XX);
  const F U2 = quad_fmul(q, W, a.y);
  XYZZ<F> r;
  r.x = f_sub(f_sub(quad_fsqr(q, M), S), S);
  r.y = f_sub(quad_fmul(q, f_sub(S, r.x), M), U2);
  r.zz = quad_fmul(q, V, a.zz);
  r.zzz = quad_fmul(q, W, a.zzz);
  return r;
}
template <class P>
GMSM_D void xyzz_add_quad(const Quad& q, XYZZ<Fp2<P>>& p, const XYZZ<Fp2<P>>& a) {
  using F = Fp2<P>;
  if (a.zz.is_zero()) return;
  if (p.zz.is_zero()) {
    p = a;
    return;
  }
  const F U2 = quad_fmul(q, a.x, p.zz);
  const F U1 = quad_fmul(q, p.x, a.zz);
  const F S2 = quad_fmul(q, a.y, p.zzz);
  const F S1 = quad_fmul(q, p.y, a.zzz);
  const F Pp = f_sub(U2, U1);
  const F R = f_sub(S2, S1);
  if (Pp.is_zero()) {
    if (R.is_zero()) {
      p = xyzz_double_quad(q, a);
    } else {
      p.zz = F::zero();
      p.zzz = F::zero();
    }
    return;
  }
  const F PP = quad_fsqr(q, Pp);
  const F PPP = quad_fmul(q, Pp, PP);
  const F Q = quad_fmul(q, U1, PP);
  const F V = quad_fmul(q, S1, PPP);
  const F X3 = f_sub(f_sub(f_sub(quad_fsqr(q, R), PPP), Q), Q);
  p.x = X3;
  p.y = f_sub(quad_fmul(q, f_sub(Q, X3), R), V);
  p.zz = quad_fmul(q, quad_fmul(q, p.zz, a.zz), PP);
  p.zzz = quad_fmul(q, quad_fmul(q, p.zzz, a.zzz), PPP);
}
template <class P>
GMSM_D Jac<Fp2<P>> jac_double_quad(const Quad& q, const Jac<Fp2<P>>& p) {
  using F = Fp2<P>;
  const F A = quad_fsqr(q, p.x);
  const F B = quad_fsqr(q, p.y);
  const F C = quad_fsqr(q, B);
  F D = f_sub(f_sub(quad_fsqr(q, f_add(p.x, B)), A), C);
  D = f_dbl(D);
  const F E = f_add(f_dbl(A), A);
  const F Fq = quad_fsqr(q, E);
  Jac<F> r;
  r.z = f_dbl(quad_fmul(q, p.y, p.z));
  r.x = f_sub(Fq, f_dbl(D));
  const F C8 = f_dbl(f_dbl(f_dbl(C)));
  r.y = f_sub(quad_fmul(q, E, f_sub(D, r.x)), C8);
  return r;
}
template <class P>
GMSM_D XYZZ<Fp2<P>> jac_to_xyzz_quad(const Quad& q, const Jac<Fp2<P>>& p) {
  XYZZ<Fp2<P>> r;
  r.x = p.x;
  r.y = p.y;
  r.zz = quad_fsqr(q, p.z);
  r.zzz = quad_fmul(q, r.zz, p.z);
  return r;
}
template <class P>
GMSM_D Jac<Fp2<P>> xyzz_to_jac_quad(const Quad& q, const XYZZ<Fp2<P>>& p) {
  using F = Fp2<P>;
  if (p.zz.is_zero()) return Jac<F>{F::zero(), F::zero(), F::zero()};
  Jac<F> j;
  j.x = quad_fmul(q, quad_fsqr(q, p.zz), p.x);
  j.y = quad_fmul(q, quad_fsqr(q, p.zzz), p.y);
  j.z = p.zzz;
  return j;
}

}  // namespace gmsm
