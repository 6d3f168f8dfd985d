/* fp_tmpl.h -- prime-field template (TEST INFRASTRUCTURE: part of the CPU oracle, never linked into
 * the product).  Instantiate with:
 *   #define NL <64-bit limbs>   #define FP(name) <prefix>_##name
 *   static const uint64_t FP(Q)[NL], FP(QINV) defined beforehand.
 * Restates ecc/bn254/fp/element.go (reference tree): Add :386-401, Double :403-418, Sub :420-438,
 * Neg :440-454, _mulGeneric (CIOS) :470-591, fromMont :593-642, toMont :782-784, Inverse(0)=0
 * :1170-1172 (any correct inverse gives the same unique reduced value; Fermat used here).
 */
typedef struct { uint64_t l[NL]; } FP(t);

static inline int FP(is_zero)(const FP(t)* a) {
  uint64_t o = 0;
  for (int i = 0; i < NL; i++) o |= a->l[i];
  return o == 0;
}
static inline int FP(eq)(const FP(t)* a, const FP(t)* b) {
  uint64_t o = 0;
  for (int i = 0; i < NL; i++) o |= a->l[i] ^ b->l[i];
  return o == 0;
}
static inline void FP(set_zero)(FP(t)* a) { for (int i = 0; i < NL; i++) a->l[i] = 0; }
/* a >= q ? */
static inline int FP(geq_q)(const uint64_t* a) {
  for (int i = NL - 1; i >= 0; i--) {
    if (a[i] > FP(Q)[i]) return 1;
    if (a[i] < FP(Q)[i]) return 0;
  }
  return 1;
}
static inline void FP(sub_q)(uint64_t* a) {
  unsigned __int128 br = 0;
  for (int i = 0; i < NL; i++) {
    unsigned __int128 d = (unsigned __int128)a[i] - FP(Q)[i] - br;
    a[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
}
static inline void FP(add)(FP(t)* z, const FP(t)* x, const FP(t)* y) {
  unsigned __int128 c = 0;
  uint64_t t[NL];
  for (int i = 0; i < NL; i++) {
    c += (unsigned __int128)x->l[i] + y->l[i];
    t[i] = (uint64_t)c;
    c >>= 64;
  }
  if (c || FP(geq_q)(t)) FP(sub_q)(t); /* c is always 0 when q has a spare top bit; full-width moduli (secp256k1) carry out */
  for (int i = 0; i < NL; i++) z->l[i] = t[i];
}
static inline void FP(dbl)(FP(t)* z, const FP(t)* x) { FP(add)(z, x, x); }
static inline void FP(sub)(FP(t)* z, const FP(t)* x, const FP(t)* y) {
  unsigned __int128 br = 0;
  uint64_t t[NL];
  for (int i = 0; i < NL; i++) {
    unsigned __int128 d = (unsigned __int128)x->l[i] - y->l[i] - br;
    t[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
  if (br) {
    unsigned __int128 c = 0;
    for (int i = 0; i < NL; i++) {
      c += (unsigned __int128)t[i] + FP(Q)[i];
      t[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  for (int i = 0; i < NL; i++) z->l[i] = t[i];
}
static inline void FP(neg)(FP(t)* z, const FP(t)* x) {
  if (FP(is_zero)(x)) { FP(set_zero)(z); return; }
  FP(t) q;
  for (int i = 0; i < NL; i++) q.l[i] = FP(Q)[i];
  FP(sub)(z, &q, x);
}
/* Montgomery multiplication.  "No-carry" CIOS with the two carry chains of a row interleaved, the form
 * of the reference's portable path (ecc/bn254/fp/element_purego.go:46-213; the spare top bit of q makes
 * the final word addition carry-free, field/generator/config/field_config.go:203-206); same result as the
 * textbook CIOS _mulGeneric fp/element.go:470-591. */
#ifdef FP_FULL
/* full-width modulus (no spare top bit: secp256k1 fp and fr): the reference generates the textbook CIOS with two extra words
 * for these (_mulGeneric, ecc/secp256k1/fp/element.go; noCarry is false in field/generator/config/field_config.go:203-206) and
 * the final subtraction looks at the carry word */
static inline void FP(mul)(FP(t)* z, const FP(t)* x, const FP(t)* y) {
  uint64_t t[NL + 2];
  for (int j = 0; j < NL + 2; j++) t[j] = 0;
  for (int i = 0; i < NL; i++) {
    unsigned __int128 A;
    uint64_t c = 0;
    for (int j = 0; j < NL; j++) {
      A = (unsigned __int128)x->l[j] * y->l[i] + t[j] + c;
      t[j] = (uint64_t)A;
      c = (uint64_t)(A >> 64);
    }
    A = (unsigned __int128)t[NL] + c;
    t[NL] = (uint64_t)A;
    t[NL + 1] = (uint64_t)(A >> 64);
    const uint64_t m = t[0] * FP(QINV);
    A = (unsigned __int128)m * FP(Q)[0] + t[0];
    c = (uint64_t)(A >> 64);
    for (int j = 1; j < NL; j++) {
      A = (unsigned __int128)m * FP(Q)[j] + t[j] + c;
      t[j - 1] = (uint64_t)A;
      c = (uint64_t)(A >> 64);
    }
    A = (unsigned __int128)t[NL] + c;
    t[NL - 1] = (uint64_t)A;
    t[NL] = t[NL + 1] + (uint64_t)(A >> 64);
  }
  if (t[NL] || FP(geq_q)(t)) FP(sub_q)(t);
  for (int i = 0; i < NL; i++) z->l[i] = t[i];
}
#else
static inline void FP(mul)(FP(t)* z, const FP(t)* x, const FP(t)* y) {
  uint64_t t[NL];
  for (int j = 0; j < NL; j++) t[j] = 0;
#pragma GCC unroll 8
  for (int i = 0; i < NL; i++) {
    const uint64_t yi = y->l[i];
    unsigned __int128 A = (unsigned __int128)x->l[0] * yi + t[0];
    const uint64_t m = (uint64_t)A * FP(QINV);
    unsigned __int128 B = (unsigned __int128)m * FP(Q)[0] + (uint64_t)A;
    uint64_t ca = (uint64_t)(A >> 64), cb = (uint64_t)(B >> 64);
#pragma GCC unroll 8
    for (int j = 1; j < NL; j++) {
      A = (unsigned __int128)x->l[j] * yi + t[j] + ca;
      ca = (uint64_t)(A >> 64);
      B = (unsigned __int128)m * FP(Q)[j] + (uint64_t)A + cb;
      cb = (uint64_t)(B >> 64);
      t[j - 1] = (uint64_t)B;
    }
    t[NL - 1] = ca + cb;
  }
  if (FP(geq_q)(t)) FP(sub_q)(t);
  for (int i = 0; i < NL; i++) z->l[i] = t[i];
}
#endif
static inline void FP(sqr)(FP(t)* z, const FP(t)* x) { FP(mul)(z, x, x); }
static inline void FP(from_mont)(FP(t)* z, const FP(t)* x) {
  FP(t) one;
  FP(set_zero)(&one);
  one.l[0] = 1;
  FP(mul)(z, x, &one);
}
static inline void FP(to_mont)(FP(t)* z, const FP(t)* x) {
  FP(t) r2;
  for (int i = 0; i < NL; i++) r2.l[i] = FP(R2)[i];
  FP(mul)(z, x, &r2);
}
static inline void FP(set_one)(FP(t)* z) {
  for (int i = 0; i < NL; i++) z->l[i] = FP(ONE)[i];
}
static inline void FP(set_u64)(FP(t)* z, uint64_t v) {
  FP(t) c;
  FP(set_zero)(&c);
  c.l[0] = v;
  FP(to_mont)(z, &c);
}
/* x^(q-2) */
static void FP(inv)(FP(t)* z, const FP(t)* x) {
  if (FP(is_zero)(x)) { FP(set_zero)(z); return; }
  uint64_t e[NL];
  unsigned __int128 br = 2;
  for (int i = 0; i < NL; i++) {
    unsigned __int128 d = (unsigned __int128)FP(Q)[i] - br;
    e[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
  FP(t) acc, base = *x;
  FP(set_one)(&acc);
  int started = 0;
  for (int i = 64 * NL - 1; i >= 0; i--) {
    if (started) FP(sqr)(&acc, &acc);
    if ((e[i >> 6] >> (i & 63)) & 1) {
      if (started) FP(mul)(&acc, &acc, &base); else { acc = base; started = 1; }
    }
  }
  *z = acc;
}
