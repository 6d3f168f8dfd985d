// Microbenchmarks behind DESIGN.md's multiplier discussion ("settle FP64 with a measured
// microbenchmark").  Standalone binary, no product code path:
//   part A  issue rate of the instructions a Montgomery product is made of, per SM and clock:
//           IMAD.WIDE (64-bit accumulate), IMAD (lo), IMAD.HI, DFMA, DADD, IADD3, and mixes of them in one warp
//   part B  modmul/s of the product routines: field.cuh's CIOS multiplier / dedicated squaring / fused two-product
//           (32-bit limbs, IMAD.WIDE) against a 52-bit-limb Montgomery product on the FP64 pipe (DFMA hi/lo split),
//           alone and with the two kinds of warps co-resident on every SM sub-partition
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I gnark-crypto_b200/csrc -o tools/ubench/ubench tools/ubench/ubench.cu
// Run  :  tools/ubench/ubench            (prints one line per measurement + a self-check of the FP64 product)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define GMSM_SQR_DEDICATED 1
#define GMSM_DOT2 1
#include "field.cuh"

using namespace gmsm;

#define CKU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// ------------------------------------------------------------------------------------------------
// part A: instruction issue rates
// ------------------------------------------------------------------------------------------------
enum { OP_IMADWIDE = 0, OP_IMADLO, OP_IMADHI, OP_DFMA, OP_DADD, OP_IADD3, OP_MADC_PAIR, OP_MIX_W_D, OP_MIX_W_2D, OP_MIX_W_D_A, OP_MIX_W_A, OP_LOP3, OP_COUNT };
static const char* OP_NAME[OP_COUNT] = {"IMAD.WIDE.U32 (mad.wide.u32, 64-bit acc)", "IMAD (mad.lo.u32)", "IMAD.HI (mad.hi.u32)", "DFMA (fma.rz.f64)",
                                        "DADD (add.f64)", "IADD3 (3-input add)", "mad.lo.cc+madc.hi.cc pair (counted as 1)", "mix: 1 IMAD.WIDE + 1 DFMA",
                                        "mix: 1 IMAD.WIDE + 2 DFMA", "mix: 1 IMAD.WIDE + 2 DFMA + 1 IADD3", "mix: 1 IMAD.WIDE + 2 IADD3", "LOP3 (xor-and)"};
static const int OP_PER_ITER[OP_COUNT] = {1, 1, 1, 1, 1, 1, 1, 2, 3, 4, 3, 1};

template <int OP>
__global__ void __launch_bounds__(512) k_rate(uint64_t* out, int iters, long long* cyc) {
  constexpr int C = 8;   // independent chains per thread
  uint64_t w[C];
  uint32_t a[C], b[C], s[C];
  double d[C], e[C];
  const uint32_t seed = threadIdx.x * 2654435761u + blockIdx.x;
#pragma unroll
  for (int i = 0; i < C; i++) {
    w[i] = seed + i; a[i] = seed * (i + 3) | 1u; b[i] = (seed >> 3) + i * 77u; s[i] = seed ^ (i * 0x9e3779b9u);
    d[i] = 1.0 + 1e-9 * (double)(seed % 1000 + i); e[i] = 1.0 + 1e-12 * (double)(i + 1);
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < C; i++) {
      if (OP == OP_IMADWIDE || OP == OP_MIX_W_D || OP == OP_MIX_W_2D || OP == OP_MIX_W_D_A || OP == OP_MIX_W_A)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));
      if (OP == OP_IMADLO) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s[i]) : "r"(a[i]), "r"(b[i]));
      if (OP == OP_IMADHI) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(s[i]) : "r"(a[i]), "r"(b[i]));
      if (OP == OP_DFMA || OP == OP_MIX_W_D || OP == OP_MIX_W_2D || OP == OP_MIX_W_D_A)
        asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e[i]), "d"(e[(i + 1) % C]));
      if (OP == OP_MIX_W_2D || OP == OP_MIX_W_D_A)
        asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(e[i]) : "d"(d[(i + 3) % C]), "d"(d[(i + 5) % C]));
      if (OP == OP_DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(e[i]));
      if (OP == OP_IADD3 || OP == OP_MIX_W_D_A || OP == OP_MIX_W_A)
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(s[i]) : "r"(a[i]), "r"(b[i]));
      if (OP == OP_MIX_W_A)
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(s[(i + 1) % C]), "r"(b[i]));
      if (OP == OP_LOP3) asm volatile("{ .reg .u32 t; xor.b32 t, %0, %1; and.b32 %0, t, %2; }" : "+r"(s[i]) : "r"(a[i]), "r"(b[i]));
      if (OP == OP_MADC_PAIR) {
        uint32_t lo = (uint32_t)w[i], hi = (uint32_t)(w[i] >> 32);
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a[i]), "r"(b[i]));
        w[i] = ((uint64_t)hi << 32) | lo;
      }
    }
  }
  const long long t1 = clock64();
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < C; i++) acc += w[i] + s[i] + a[i] + (uint64_t)__double_as_longlong(d[i]) + (uint64_t)__double_as_longlong(e[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run_rate(int nsm, int threads, uint64_t* d_out, long long* d_cyc) {
  const int iters = 4000;
  const int bps = threads / 512 > 0 ? threads / 512 : 1, tpb = threads / bps;   // blocks of <= 512 threads (128 registers each)
  k_rate<OP><<<nsm * bps, tpb>>>(d_out, 100, d_cyc);
  k_rate<OP><<<nsm * bps, tpb>>>(d_out, iters, d_cyc);
  CKU(cudaDeviceSynchronize());
  std::vector<long long> c(nsm * bps);
  CKU(cudaMemcpy(c.data(), d_cyc, nsm * bps * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (long long v : c) avg += (double)v;
  avg /= nsm * bps;
  const double ops = (double)threads * iters * 8 * OP_PER_ITER[OP];
  printf("RATE %-46s threads/SM=%4d  %.2f thread-instr/clk/SM  (%.3f warp-instr/clk/SMSP)\n", OP_NAME[OP], threads, ops / avg, ops / avg / 128.0);
}

// ------------------------------------------------------------------------------------------------
// part B: Montgomery products
// ------------------------------------------------------------------------------------------------
// 52-bit-limb Montgomery product on the FP64 pipe: z = x*y*2^(-52*NL) mod q (+q), limbs kept as integers < 2^52.
// One 52x52 -> 104-bit limb product = 2 DFMA + 1 DADD:  hi = fma_rz(a, b, 2^104) = 2^104 + floor(ab / 2^52) * 2^52 (truncation makes the
// mantissa field of hi exactly the high half), lo = fma_rz(a, b, (2^104 + 2^52) - hi) = 2^52 + (ab mod 2^52), exact.  The raw bit
// patterns of hi and lo are exponent | integer, so the column sums are plain 64-bit integer additions and the exponent
// fields are removed as compile-time constants on the high words.
template <int NL> struct Q52;
template <> struct Q52<5> {   // bn254 fp
  static constexpr uint64_t INV = 0x20782e4866389ull;
  __host__ __device__ static constexpr uint64_t q(int i) {
    constexpr uint64_t t[5] = {0x8c16d87cfd47ull, 0x916871ca8d3c2ull, 0x181585d97816aull, 0xa029b85045b68ull, 0x30644e72e131ull};
    return t[i];
  }
};
template <> struct Q52<8> {   // bls12-381 fp
  static constexpr uint64_t INV = 0x3fffcfffcfffdull;
  __host__ __device__ static constexpr uint64_t q(int i) {
    constexpr uint64_t t[8] = {0xeffffffffaaabull, 0xfeb153ffffb9full, 0x6b0f6241eabffull, 0x12bf6730d2a0full, 0x764774b84f385ull, 0x1ba7b6434bacdull, 0x1ea397fe69a4bull, 0x1a011ull};
    return t[i];
  }
};
static constexpr uint64_t MASK52 = (1ull << 52) - 1;
static constexpr uint64_t LO_BIAS = 0x433ull << 52;   // bits of 2^52
static constexpr uint64_t HI_BIAS = 0x467ull << 52;   // bits of 2^104

__device__ __forceinline__ double u52_to_double(uint64_t x) {   // x < 2^52
  return __longlong_as_double((long long)(x | LO_BIAS)) - 4503599627370496.0;
}

template <int NL>
__device__ __forceinline__ void mont52(const uint64_t* x, const uint64_t* y, uint64_t* z) {
  const double C1 = 20282409603651670423947251286016.0;            // 2^104
  const double C12 = 20282409603651674927546878656512.0;           // 2^104 + 2^52
  double a[NL], b[NL], qd[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) { a[i] = u52_to_double(x[i]); b[i] = u52_to_double(y[i]); qd[i] = (double)Q52<NL>::q(i); }
  uint64_t c[NL + 1];
#pragma unroll
  for (int i = 0; i <= NL; i++) c[i] = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
#pragma unroll
    for (int j = 0; j < NL; j++) {
      const double hi = __fma_rz(a[j], b[i], C1);
      const double lo = __fma_rz(a[j], b[i], C12 - hi);
      c[j] += (uint64_t)__double_as_longlong(lo);
      c[j + 1] += (uint64_t)__double_as_longlong(hi);
    }
    // the exponent fields are multiples of 2^52: c[0] mod 2^52 is already right
    const uint64_t m = ((c[0] & MASK52) * Q52<NL>::INV) & MASK52;
    const double md = u52_to_double(m);
#pragma unroll
    for (int j = 0; j < NL; j++) {
      const double hi = __fma_rz(md, qd[j], C1);
      const double lo = __fma_rz(md, qd[j], C12 - hi);
      c[j] += (uint64_t)__double_as_longlong(lo);
      c[j + 1] += (uint64_t)__double_as_longlong(hi);
    }
    // remove the exponent fields of this row (2 lo + 2 hi terms per inner column), then shift one limb
    c[0] -= 2 * LO_BIAS;
#pragma unroll
    for (int j = 1; j < NL; j++) c[j] -= 2 * LO_BIAS + 2 * HI_BIAS;
    c[NL] -= 2 * HI_BIAS;
    const uint64_t carry = c[0] >> 52;
    c[0] = c[1] + carry;
#pragma unroll
    for (int j = 1; j < NL; j++) c[j] = c[j + 1];
    c[NL] = 0;
  }
  uint64_t carry = 0;
#pragma unroll
  for (int j = 0; j < NL - 1; j++) {
    const uint64_t v = c[j] + carry;
    z[j] = v & MASK52;
    carry = v >> 52;
  }
  z[NL - 1] = c[NL - 1] + carry;
}

enum { K_MUL32 = 0, K_SQR32, K_DOT2_32, K_MUL52, K_MIXWARP, K_MUL32_12, K_MUL52_8, K_MIXWARP_381 };

template <class P>
__device__ __forceinline__ void chain32(int kind, uint32_t seed, int iters, uint32_t* out) {
  Fp<P> x, y;
#pragma unroll
  for (int i = 0; i < P::N; i++) { x.l[i] = seed * (i + 1) + 0x1234567u * i; y.l[i] = (seed ^ 0x5a5a5a5au) + i; }
  x.l[P::N - 1] &= 0x0fffffffu; y.l[P::N - 1] &= 0x0fffffffu;
  if (kind == K_SQR32) {
    for (int it = 0; it < iters; it++) { x = fp_sqr(x); y = fp_sqr(y); }
  } else if (kind == K_DOT2_32) {
    for (int it = 0; it < iters; it++) { Fp<P> t = fp_dot2(x, y, y, x); y = x; x = t; }   // counted as 2 products
  } else {
    for (int it = 0; it < iters; it++) { x = fp_mul(x, y); y = fp_mul(y, x); }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < P::N; i++) acc ^= x.l[i] + y.l[i];
  *out = acc;
}
template <int NL>
__device__ __forceinline__ void chain52(uint32_t seed, int iters, uint32_t* out) {
  uint64_t x[NL], y[NL], t[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) { x[i] = ((uint64_t)seed * 0x9e3779b97f4a7c15ull + i) & MASK52; y[i] = ((uint64_t)(seed + 7) * 0xc2b2ae3d27d4eb4full + i) & MASK52; }
  x[NL - 1] &= (1ull << 40) - 1; y[NL - 1] &= (1ull << 40) - 1;
  for (int it = 0; it < iters; it++) {
    mont52<NL>(x, y, t);
#pragma unroll
    for (int i = 0; i < NL; i++) x[i] = t[i];
    mont52<NL>(y, x, t);
#pragma unroll
    for (int i = 0; i < NL; i++) y[i] = t[i];
  }
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) acc ^= x[i] + y[i];
  *out = (uint32_t)acc ^ (uint32_t)(acc >> 32);
}

template <int KIND>
__global__ void __launch_bounds__(256) k_modmul(uint32_t* out, int iters) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t seed = gid * 747796405u + 2891336453u;
  if (KIND == K_MUL32 || KIND == K_SQR32 || KIND == K_DOT2_32) chain32<bn254_fp>(KIND, seed, iters, out + gid);
  if (KIND == K_MUL32_12) chain32<bls12381_fp>(K_MUL32, seed, iters, out + gid);
  if (KIND == K_MUL52) chain52<5>(seed, iters, out + gid);
  if (KIND == K_MUL52_8) chain52<8>(seed, iters, out + gid);
  if (KIND == K_MIXWARP) {   // alternate warps: both kinds resident on every SM sub-partition (warp id % 4 = SMSP; pairs 0..3 / 4..7)
    if (((threadIdx.x >> 5) >> 2) & 1) chain52<5>(seed, iters, out + gid); else chain32<bn254_fp>(K_MUL32, seed, iters, out + gid);
  }
  if (KIND == K_MIXWARP_381) {
    if (((threadIdx.x >> 5) >> 2) & 1) chain52<8>(seed, iters, out + gid); else chain32<bls12381_fp>(K_MUL32, seed, iters, out + gid);
  }
}

template <int KIND>
static double run_modmul(const char* name, int blocks_per_sm, int nsm, uint32_t* d_out, double per_iter) {
  const int iters = 400;
  const int blocks = nsm * blocks_per_sm;
  cudaEvent_t e0, e1;
  CKU(cudaEventCreate(&e0)); CKU(cudaEventCreate(&e1));
  k_modmul<KIND><<<blocks, 256>>>(d_out, 10);
  CKU(cudaDeviceSynchronize());
  CKU(cudaEventRecord(e0));
  k_modmul<KIND><<<blocks, 256>>>(d_out, iters);
  CKU(cudaEventRecord(e1));
  CKU(cudaEventSynchronize(e1));
  float ms = 0;
  CKU(cudaEventElapsedTime(&ms, e0, e1));
  const double n = (double)blocks * 256 * iters * per_iter;
  const double rate = n / (ms * 1e-3);
  printf("MODMUL %-58s warps/SM=%2d  %.3e modmul/s  (%.2f ms)\n", name, blocks_per_sm * 8, rate, ms);
  return rate;
}

// self-check of mont52<5>: z * 2^260 == x * y (mod q), verified on the host with 128-bit arithmetic on 52-bit limbs
template <int NL>
__global__ void k_check52(const uint64_t* x, const uint64_t* y, uint64_t* z, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t a[NL], b[NL], r[NL];
  for (int k = 0; k < NL; k++) { a[k] = x[i * NL + k]; b[k] = y[i * NL + k]; }
  mont52<NL>(a, b, r);
  for (int k = 0; k < NL; k++) z[i * NL + k] = r[k];
}
template <int NL>
static void host_mont52(const uint64_t* x, const uint64_t* y, uint64_t* z) {
  unsigned __int128 c[NL + 2] = {};
  for (int i = 0; i < NL; i++) {
    for (int j = 0; j < NL; j++) c[j] += (unsigned __int128)x[j] * y[i];
    for (int j = 0; j < NL; j++) { c[j + 1] += c[j] >> 52; c[j] &= MASK52; }
    const uint64_t m = ((uint64_t)c[0] * Q52<NL>::INV) & MASK52;
    for (int j = 0; j < NL; j++) c[j] += (unsigned __int128)m * Q52<NL>::q(j);
    for (int j = 0; j < NL; j++) { c[j + 1] += c[j] >> 52; c[j] &= MASK52; }
    for (int j = 0; j < NL; j++) c[j] = c[j + 1];
    c[NL] = 0;
  }
  for (int j = 0; j < NL; j++) z[j] = (uint64_t)c[j];
}
template <int NL>
static bool check52() {
  const int n = 4096;
  std::vector<uint64_t> x(n * NL), y(n * NL), z(n * NL);
  uint64_t s = 0x1234567887654321ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (int i = 0; i < n; i++)
    for (int k = 0; k < NL; k++) {
      x[i * NL + k] = rnd() & MASK52; y[i * NL + k] = rnd() & MASK52;
      if (k == NL - 1) { x[i * NL + k] = (i & 1) ? Q52<NL>::q(k) : (x[i * NL + k] % (Q52<NL>::q(k) * 7)); y[i * NL + k] %= (Q52<NL>::q(k) * 7); }   // values up to ~8q
      if (i == 0) { x[k] = MASK52 * (k != NL - 1) + (k == NL - 1) * Q52<NL>::q(k); }
    }
  uint64_t *dx, *dy, *dz;
  CKU(cudaMalloc(&dx, n * NL * 8)); CKU(cudaMalloc(&dy, n * NL * 8)); CKU(cudaMalloc(&dz, n * NL * 8));
  CKU(cudaMemcpy(dx, x.data(), n * NL * 8, cudaMemcpyHostToDevice));
  CKU(cudaMemcpy(dy, y.data(), n * NL * 8, cudaMemcpyHostToDevice));
  k_check52<NL><<<(n + 127) / 128, 128>>>(dx, dy, dz, n);
  CKU(cudaMemcpy(z.data(), dz, n * NL * 8, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < n; i++) {
    uint64_t w[NL];
    host_mont52<NL>(&x[i * NL], &y[i * NL], w);
    for (int k = 0; k < NL; k++) if (w[k] != z[i * NL + k]) { bad++; break; }
  }
  printf("CHECK mont52<%d>: %d / %d products differ from the exact 128-bit host evaluation of the same recurrence\n", NL, bad, n);
  if (NL == 5) {   // a few samples for an independent big-integer check (tools/ubench/check_samples.py)
    for (int i = 0; i < 4; i++) {
      printf("SAMPLE52");
      for (int k = 0; k < NL; k++) printf(" %llx", (unsigned long long)x[i * NL + k]);
      printf(" |");
      for (int k = 0; k < NL; k++) printf(" %llx", (unsigned long long)y[i * NL + k]);
      printf(" |");
      for (int k = 0; k < NL; k++) printf(" %llx", (unsigned long long)z[i * NL + k]);
      printf("\n");
    }
  }
  cudaFree(dx); cudaFree(dy); cudaFree(dz);
  return bad == 0;
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  CKU(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, max SM clock %.0f MHz\n", prop.name, nsm, clk_khz / 1e3);
  uint64_t* d_out; long long* d_cyc;
  CKU(cudaMalloc(&d_out, (size_t)nsm * 8 * 1024 * 8));
  CKU(cudaMalloc(&d_cyc, nsm * 8 * 8));
  for (int threads : {256, 512, 1024}) {
    run_rate<OP_IMADWIDE>(nsm, threads, d_out, d_cyc);
    run_rate<OP_MADC_PAIR>(nsm, threads, d_out, d_cyc);
    run_rate<OP_IMADLO>(nsm, threads, d_out, d_cyc);
    run_rate<OP_IMADHI>(nsm, threads, d_out, d_cyc);
    run_rate<OP_DFMA>(nsm, threads, d_out, d_cyc);
    run_rate<OP_DADD>(nsm, threads, d_out, d_cyc);
    run_rate<OP_IADD3>(nsm, threads, d_out, d_cyc);
    run_rate<OP_LOP3>(nsm, threads, d_out, d_cyc);
    run_rate<OP_MIX_W_D>(nsm, threads, d_out, d_cyc);
    run_rate<OP_MIX_W_2D>(nsm, threads, d_out, d_cyc);
    run_rate<OP_MIX_W_D_A>(nsm, threads, d_out, d_cyc);
    run_rate<OP_MIX_W_A>(nsm, threads, d_out, d_cyc);
  }
  bool ok = check52<5>();
  ok = check52<8>() && ok;
  uint32_t* d_o32 = reinterpret_cast<uint32_t*>(d_out);
  for (int bps : {1, 2, 4}) {
    run_modmul<K_MUL32>("bn254 fp  CIOS 32-bit limbs (IMAD.WIDE): mul", bps, nsm, d_o32, 2);
    run_modmul<K_SQR32>("bn254 fp  dedicated squaring (IMAD.WIDE)", bps, nsm, d_o32, 2);
    run_modmul<K_DOT2_32>("bn254 fp  fused x*y+u*v, one reduction (as 2 products)", bps, nsm, d_o32, 2);
    run_modmul<K_MUL52>("bn254 fp  52-bit limbs on the FP64 pipe (DFMA)", bps, nsm, d_o32, 2);
    run_modmul<K_MIXWARP>("bn254 fp  half the warps IMAD.WIDE, half DFMA", bps, nsm, d_o32, 2);
    run_modmul<K_MUL32_12>("bls12-381 fp CIOS 32-bit limbs (IMAD.WIDE): mul", bps, nsm, d_o32, 2);
    run_modmul<K_MUL52_8>("bls12-381 fp 52-bit limbs on the FP64 pipe (DFMA)", bps, nsm, d_o32, 2);
    run_modmul<K_MIXWARP_381>("bls12-381 fp half the warps IMAD.WIDE, half DFMA", bps, nsm, d_o32, 2);
  }
  printf(ok ? "UBENCH OK\n" : "UBENCH CHECK FAILED\n");
  return ok ? 0 : 1;
}
