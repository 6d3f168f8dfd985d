// TEST INFRASTRUCTURE ONLY -- the Fr FFT kernels (gnark-crypto_b200/csrc/fft_kernels.cuh, the header fft.cu includes) on
// the CPU, launched in the order of fft.cu's domain_build / run_fft: twiddle table by k_fft_powers, coset scaling,
// strided DIF / DIT stages, the shared-memory tile kernel (barriers: cooperative launcher), final scaling, bit reversal.
// The domain constants (Generator, GeneratorInv, CardinalityInv, coset shift and its inverse; Montgomery limbs) are
// passed in by the test, which takes them from the oracle's restatement of fft.NewDomain.
#include <algorithm>
#include <cstring>
#include <vector>

#include "fft_kernels.cuh"

namespace {
// the dynamic shared memory of k_fft_tile (`extern __shared__ smem_raw[]`): 1024 elements of 32 bytes
thread_local __attribute__((aligned(16))) unsigned char smem_raw[TILE * 32];

template <class P>
int emu_fft(uint32_t* a_words, uint64_t n, int logn, int inverse, int decimation, int coset, const uint32_t* consts5, int bit_reverse_only) {
  using F = Fp<P>;
  F* a = reinterpret_cast<F*>(a_words);
  auto grid = [](uint64_t work) { return (unsigned)std::min<uint64_t>((work + 255) / 256, 8u); };
  if (bit_reverse_only) {
    emu_launch(k_fft_bit_reverse<P>, dim3(grid(n)), 256u, a, n, logn);
    return 0;
  }
  F gen, gen_inv, card_inv, shift, shift_inv;
  std::memcpy(gen.l, consts5, 32); std::memcpy(gen_inv.l, consts5 + 8, 32); std::memcpy(card_inv.l, consts5 + 16, 32);
  std::memcpy(shift.l, consts5 + 24, 32); std::memcpy(shift_inv.l, consts5 + 32, 32);
  // domain_build: pw[0..63] = shift^(2^k), pw[64..127] = shift^-(2^k), pw[128..191] = w^(2^k); tw[j] = w^j, j < n/2
  std::vector<F> pw(192);
  {
    F x = shift, y = shift_inv, w = inverse ? gen_inv : gen;
    for (int k = 0; k < 64; k++) { pw[k] = x; pw[64 + k] = y; pw[128 + k] = w; x = fp_sqr(x); y = fp_sqr(y); w = fp_sqr(w); }
  }
  const uint64_t half = n >> 1;
  std::vector<F> tw(std::max<uint64_t>(half, 1));
  if (half) emu_launch(k_fft_powers<P>, dim3(grid(half)), 256u, tw.data(), half, (const F*)(pw.data() + 128), logn > 0 ? logn - 1 : 0);
  // run_fft
  const F one = F::one();
  if (!inverse && coset) emu_launch(k_fft_scale<P>, dim3(grid(n)), 256u, a, n, logn, (const F*)pw.data(), 1, (int)(decimation == 0), one, 0);
  if (n > 1) {
    const uint32_t tile = (uint32_t)std::min<uint64_t>(n, TILE);
    if (decimation == 1) {
      for (uint64_t h = half; h >= tile; h >>= 1) emu_launch(k_fft_dif_stage<P>, dim3(grid(half)), 256u, a, (const F*)tw.data(), half, h, half / h);
      emu_launch_coop(k_fft_tile<P, true>, dim3((unsigned)(n / tile)), tile / 2, a, (const F*)tw.data(), n, tile);
    } else {
      emu_launch_coop(k_fft_tile<P, false>, dim3((unsigned)(n / tile)), tile / 2, a, (const F*)tw.data(), n, tile);
      for (uint64_t h = tile; h <= half; h <<= 1) emu_launch(k_fft_dit_stage<P>, dim3(grid(half)), 256u, a, (const F*)tw.data(), half, h, half / h);
    }
  }
  if (inverse) emu_launch(k_fft_scale<P>, dim3(grid(n)), 256u, a, n, logn, (const F*)(pw.data() + 64), coset ? 1 : 0, (int)(decimation == 1), card_inv, 1);
  return 0;
}
}  // namespace

// field: 0 bn254 fr, 1 bls12-381 fr, 2 bls12-377 fr (GMSM_FR_*); a: n x 8 u32 Montgomery limbs, transformed in place
extern "C" int emu_fft_run(int field, uint32_t* a, uint64_t n, int logn, int inverse, int decimation, int coset, const uint32_t* consts5,
                           int bit_reverse_only) {
  if (n == 0 || (n & (n - 1)) || (1ull << logn) != n) return 1;
  switch (field) {
    case 0: return emu_fft<bn254_fr>(a, n, logn, inverse, decimation, coset, consts5, bit_reverse_only);
    case 1: return emu_fft<bls12381_fr>(a, n, logn, inverse, decimation, coset, consts5, bit_reverse_only);
    case 2: return emu_fft<bls12377_fr>(a, n, logn, inverse, decimation, coset, consts5, bit_reverse_only);
  }
  return 1;
}
