// CPU build of the arithmetic headers (portable C++ path of field.cuh) for formula tests without a GPU.
// NOT part of the product library: libgmsm.so never links this file and has no CPU path.  Built by
// tests/test_hostcheck.py with g++ into build/libgmsm_hostcheck.so.
#include <cstddef>
#include <cstdint>

#include "testops.cuh"

using namespace gmsm;

template <class G>
static int run(int op, const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) {
  int wa, wb, wo;
  test_op_sizes<G>(op, &wa, &wb, &wo);
  if (wo == 0) return 1;
  for (size_t i = 0; i < n; i++) test_op_one<G>(op, a + i * wa, b + i * wb, o + i * wo);
  return 0;
}

// -DHC_GROUP=k compiles one group per translation unit (hostcheck_op_k) so the test can build them in parallel;
// without it this file is the dispatcher + the window plan.
#if defined(HC_GROUP)
#define HC_CAT2(a, b) a##b
#define HC_CAT(a, b) HC_CAT2(a, b)
#if HC_GROUP == 0
using HcG = bn254_g1;
#elif HC_GROUP == 1
using HcG = bn254_g2;
#elif HC_GROUP == 2
using HcG = bls12381_g1;
#elif HC_GROUP == 3
using HcG = bls12381_g2;
#elif HC_GROUP == 4
using HcG = bls12377_g1;
#elif HC_GROUP == 5
using HcG = bls12377_g2;
#elif HC_GROUP == 6
using HcG = secp256k1_g1;
#elif HC_GROUP == 7
using HcG = bw6761_g1;
#elif HC_GROUP == 8
using HcG = bw6761_g2;
#elif HC_GROUP == 9
using HcG = bls24315_g1;
#elif HC_GROUP == 10
using HcG = bls24317_g1;
#elif HC_GROUP == 11
using HcG = bw6633_g1;
#elif HC_GROUP == 12
using HcG = bw6633_g2;
#else
// field-only pseudo group: the secp256k1 SCALAR field as coordinate field, so that the stress vectors of the test reach the
// second full-width modulus through the same multiplier (the engine itself only needs fromMont of it)
using HcG = GroupT<13, secp256k1_fr, secp256k1_fr, false>;
#endif
extern "C" int HC_CAT(hostcheck_op_, HC_GROUP)(int op, const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) {
  return run<HcG>(op, a, b, o, n);
}
// out = (x*y + u*v) in the coordinate field (f_dot2: the fused two-product routine when built with -DGMSM_DOT2=1)
extern "C" int HC_CAT(hostcheck_dot2_, HC_GROUP)(const uint32_t* x, const uint32_t* y, const uint32_t* u, const uint32_t* v, uint32_t* o, size_t n) {
  using F = typename HcG::F;
  constexpr int W = sizeof(F) / 4;
  for (size_t i = 0; i < n; i++) wr(o + i * W, f_dot2(rd<F>(x + i * W), rd<F>(y + i * W), rd<F>(u + i * W), rd<F>(v + i * W)));
  return 0;
}
// one window-table level on the CPU, batched exactly like k_table_level: out[i] = 2^c * in[i] (affine, u32 words)
extern "C" int HC_CAT(hostcheck_table_level_, HC_GROUP)(int c, const uint32_t* in, size_t n, uint32_t* out) {
  using F = typename HcG::F;
  constexpr int AW = 2 * F::N;
  for (size_t first = 0; first < n; first += TAB_M) {
    const int cnt = (n - first < (size_t)TAB_M) ? (int)(n - first) : TAB_M;
    table_level_batch<F>(
        cnt, c, [&](int i) { return rd<Affine<F>>(in + (first + i) * AW); },
        [&](int i, const Affine<F>& a) { wr(out + (first + i) * AW, a); }, [](const Jac<F>& j) { return jac_double(j); });
  }
  return 0;
}
#else
#define HC_FOR_EACH(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13)
extern "C" {
#define X(k) int hostcheck_op_##k(int, const uint32_t*, const uint32_t*, uint32_t*, size_t); \
  int hostcheck_table_level_##k(int, const uint32_t*, size_t, uint32_t*); \
  int hostcheck_dot2_##k(const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t*, size_t);
HC_FOR_EACH(X)
#undef X
}
extern "C" int hostcheck_dot2(int curve, const uint32_t* x, const uint32_t* y, const uint32_t* u, const uint32_t* v, uint32_t* o, size_t n) {
  switch (curve) {   // meaningful for the groups whose coordinate field is Fp (for Fp2 it is the Fp2 sum of products)
#define X(k) case k: return hostcheck_dot2_##k(x, y, u, v, o, n);
    HC_FOR_EACH(X)
#undef X
  }
  return 1;
}
extern "C" int hostcheck_table_level(int curve, int c, const uint32_t* in, size_t n, uint32_t* out) {
  switch (curve) {
#define X(k) case k: return hostcheck_table_level_##k(c, in, n, out);
    HC_FOR_EACH(X)
#undef X
  }
  return 1;
}
extern "C" int hostcheck_op(int curve, int op, const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) {
  switch (curve) {
#define X(k) case k: return hostcheck_op_##k(op, a, b, o, n);
    HC_FOR_EACH(X)
#undef X
  }
  return 1;
}

extern "C" void hostcheck_plan(int fr_bits, int c, int* out6) {
  WindowPlan p = make_plan(fr_bits, c);
  out6[0] = p.c; out6[1] = p.nwin; out6[2] = p.last_c; out6[3] = (int)p.nb; out6[4] = (int)p.nb_last; out6[5] = (int)p.nb_total;
}
#endif
