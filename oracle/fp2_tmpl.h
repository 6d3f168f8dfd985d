/* fp2_tmpl.h -- E2 = A0 + A1 u, u^2 = -1 (oracle; TEST INFRASTRUCTURE).
 *   #define FP(name) base field prefix   #define F2(name) extension prefix   #define F2_NONRES5 0|1 (u^2 = -1 | -5)
 * Restates ecc/bn254/internal/fptower/e2.go:104-126, e2_bn254.go:28-73
 * (same in ecc/bls12-381/internal/fptower/e2_bls381.go:16-74).
 */
typedef struct { FP(t) a0, a1; } F2(t);
static inline void F2(mul5)(FP(t)* c) { FP(t) t; FP(dbl)(&t, c); FP(dbl)(&t, &t); FP(add)(c, &t, c); } /* fp.MulBy5 */
static inline int F2(is_zero)(const F2(t)* a) { return FP(is_zero)(&a->a0) && FP(is_zero)(&a->a1); }
static inline int F2(eq)(const F2(t)* a, const F2(t)* b) { return FP(eq)(&a->a0, &b->a0) && FP(eq)(&a->a1, &b->a1); }
static inline void F2(set_zero)(F2(t)* a) { FP(set_zero)(&a->a0); FP(set_zero)(&a->a1); }
static inline void F2(set_one)(F2(t)* a) { FP(set_one)(&a->a0); FP(set_zero)(&a->a1); }
static inline void F2(add)(F2(t)* z, const F2(t)* x, const F2(t)* y) { FP(add)(&z->a0, &x->a0, &y->a0); FP(add)(&z->a1, &x->a1, &y->a1); }
static inline void F2(sub)(F2(t)* z, const F2(t)* x, const F2(t)* y) { FP(sub)(&z->a0, &x->a0, &y->a0); FP(sub)(&z->a1, &x->a1, &y->a1); }
static inline void F2(dbl)(F2(t)* z, const F2(t)* x) { FP(dbl)(&z->a0, &x->a0); FP(dbl)(&z->a1, &x->a1); }
static inline void F2(neg)(F2(t)* z, const F2(t)* x) { FP(neg)(&z->a0, &x->a0); FP(neg)(&z->a1, &x->a1); }
static inline void F2(mul)(F2(t)* z, const F2(t)* x, const F2(t)* y) {
  FP(t) a, b, c;
  FP(add)(&a, &x->a0, &x->a1);
  FP(add)(&b, &y->a0, &y->a1);
  FP(mul)(&a, &a, &b);
  FP(mul)(&b, &x->a0, &y->a0);
  FP(mul)(&c, &x->a1, &y->a1);
  FP(sub)(&z->a1, &a, &b);
  FP(sub)(&z->a1, &z->a1, &c);
  if (F2_NONRES5) F2(mul5)(&c);   /* e2_bls377.go:12-23 */
  FP(sub)(&z->a0, &b, &c);
}
static inline void F2(sqr)(F2(t)* z, const F2(t)* x) {
  FP(t) a, b;
  FP(add)(&a, &x->a0, &x->a1);
  if (F2_NONRES5) { FP(t) t = x->a1; F2(mul5)(&t); FP(sub)(&b, &x->a0, &t); } else FP(sub)(&b, &x->a0, &x->a1);
  FP(mul)(&a, &a, &b);
  FP(mul)(&b, &x->a0, &x->a1);
  FP(dbl)(&b, &b);
  if (F2_NONRES5) { FP(t) t; FP(dbl)(&t, &b); FP(add)(&a, &a, &t); }   /* e2_bls377.go:26-38 */
  z->a0 = a;
  z->a1 = b;
}
static void F2(inv)(F2(t)* z, const F2(t)* x) {
  FP(t) t0, t1;
  FP(sqr)(&t0, &x->a0);
  FP(sqr)(&t1, &x->a1);
  if (F2_NONRES5) F2(mul5)(&t1);
  FP(add)(&t0, &t0, &t1);
  FP(inv)(&t1, &t0);
  FP(mul)(&z->a0, &x->a0, &t1);
  FP(mul)(&z->a1, &x->a1, &t1);
  FP(neg)(&z->a1, &z->a1);
}
static inline void F2(set_u64)(F2(t)* z, uint64_t v) { FP(set_u64)(&z->a0, v); FP(set_zero)(&z->a1); }
