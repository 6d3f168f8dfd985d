// Next-row N3: radix-2 FFT over Fr on the GPU, behind gnark-crypto's fft.Domain interface.
//
// Replaces (reference tree; ecc/bls12-381/fr/fft is the same generated code with its own constants):
//   fft.NewDomain / Domain{Cardinality, CardinalityInv, Generator, GeneratorInv, FrMultiplicativeGen(Inv)}
//                                                             ecc/bn254/fr/fft/domain.go:24-110
//   fr.Generator(m) (2-adic root of unity, maxOrderRoot)      ecc/bn254/fr/generator.go:18-36
//   (*Domain).FFT / FFTInverse (DIF: natural in, bit-reversed out; DIT: bit-reversed in, natural out; coset
//   option; FFTInverse scales by CardinalityInv)              ecc/bn254/fr/fft/fft.go:31-190, difFFT :195+, ditFFT :262+
//   BitReverse                                                 ecc/bn254/fr/fft/bitreverse.go:17-42
//
// Data is the reference's []fr.Element image (4 x u64 Montgomery limbs).  Kernels: one launch per butterfly
// stage for the strided stages, one shared-memory kernel for the last (DIF) / first (DIT) TILE_LOG stages,
// twiddles w^j (j < n/2) precomputed per domain like the reference's Domain.twiddles; coset powers are
// computed on the fly from u^(2^k).  HBM-bound streaming work: 32 B per element per pass.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "engine.h"
#include "field.cuh"

using namespace gmsm;

#include "fft_kernels.cuh"

namespace {

// ---- host-side field helpers (portable path of field.cuh) ----
template <class P>
Fp<P> host_from_u64(uint64_t v) {
  Fp<P> c = Fp<P>::zero();
  c.l[0] = (uint32_t)v;
  c.l[1] = (uint32_t)(v >> 32);
  return fp_to_mont(c);
}
template <class P>
Fp<P> host_from_decimal(const char* dec) {  // canonical integer < q given in decimal -> Montgomery
  Fp<P> acc = Fp<P>::zero(), ten = host_from_u64<P>(10);
  for (const char* p = dec; *p; p++) acc = fp_add(fp_mul(acc, ten), host_from_u64<P>((uint64_t)(*p - '0')));
  return acc;
}
template <class P>
Fp<P> host_pow2k(Fp<P> x, int k) {  // x^(2^k)
  for (int i = 0; i < k; i++) x = fp_sqr(x);
  return x;
}

struct FrConsts {
  const char* root;   // fr.Generator's rootOfUnity (decimal), generator.go:23
  int max_order;      // maxOrderRoot, generator.go:24
  uint64_t mult_gen;  // GeneratorFullMultiplicativeGroup, fft/domain.go:55-63
};
const FrConsts FR_BN254 = {"19103219067921713944291392827692070036145651957329286315305642004821462161904", 28, 5};
const FrConsts FR_BLS12381 = {"10238227357739495823651030575849232062558860180284477541189508159991286009131", 32, 7};
// ecc/bls12-377/fr/generator.go:23-24, fr/fft/domain.go:59
const FrConsts FR_BLS12377 = {"8065159656716812877374967518403273466521432693661810619979959746626482506078", 47, 22};

}  // namespace

struct gmsm_fft_domain {
  int field = 0, device = 0, logn = 0;
  uint64_t n = 0;
  uint64_t consts[5][4] = {};   // Generator, GeneratorInv, CardinalityInv, FrMultiplicativeGen, FrMultiplicativeGenInv
  void *d_tw = nullptr, *d_tw_inv = nullptr;       // w^j, w^-j for j < n/2
  void *d_pw = nullptr;                            // [0..63]: u^(2^k); [64..127]: u^-(2^k); [128..191]: scratch for twiddle builds
  void* d_buf = nullptr;                           // staging for the host entry points
  std::mutex mu;
};

template <class P>
static int domain_build(gmsm_fft_domain* d, const FrConsts& fc, const uint64_t* shift_mont) {
  using F = Fp<P>;
  F gen = host_pow2k(host_from_decimal<P>(fc.root), fc.max_order - d->logn);
  F gen_inv = fp_inv(gen);
  F card_inv = fp_inv(host_from_u64<P>(d->n));
  F shift;
  if (shift_mont) memcpy(shift.l, shift_mont, 32); else shift = host_from_u64<P>(fc.mult_gen);
  F shift_inv = fp_inv(shift);
  memcpy(d->consts[0], gen.l, 32); memcpy(d->consts[1], gen_inv.l, 32); memcpy(d->consts[2], card_inv.l, 32);
  memcpy(d->consts[3], shift.l, 32); memcpy(d->consts[4], shift_inv.l, 32);
  F pw[192];
  F a = shift, b = shift_inv, g = gen, gi = gen_inv;
  for (int k = 0; k < 64; k++) { pw[k] = a; pw[64 + k] = b; a = fp_sqr(a); b = fp_sqr(b); }
  CK(cudaMalloc(&d->d_pw, sizeof(pw)));
  const uint64_t half = d->n >> 1;
  CK(cudaMalloc(&d->d_tw, (half ? half : 1) * sizeof(F)));
  CK(cudaMalloc(&d->d_tw_inv, (half ? half : 1) * sizeof(F)));
  for (int pass = 0; pass < 2; pass++) {
    F w = pass ? gi : g;
    for (int k = 0; k < 64; k++) { pw[128 + k] = w; w = fp_sqr(w); }
    CK(cudaMemcpy(d->d_pw, pw, sizeof(pw), cudaMemcpyHostToDevice));
    if (half) {
      unsigned blocks = (unsigned)std::min<uint64_t>((half + 255) / 256, 148u * 16u);
      k_fft_powers<P><<<blocks, 256>>>(reinterpret_cast<F*>(pass ? d->d_tw_inv : d->d_tw), half,
                                       reinterpret_cast<const F*>(d->d_pw) + 128, d->logn > 0 ? d->logn - 1 : 0);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
    }
  }
  return GMSM_OK;
}

template <class P>
static int run_fft(gmsm_fft_domain* d, void* d_a, int inverse, int decimation, int coset, cudaStream_t st) {
  using F = Fp<P>;
  F* a = reinterpret_cast<F*>(d_a);
  const uint64_t n = d->n, half = n >> 1;
  const F* pw = reinterpret_cast<const F*>(d->d_pw);
  const F* tw = reinterpret_cast<const F*>(inverse ? d->d_tw_inv : d->d_tw);
  auto grid = [](uint64_t work) { return (unsigned)std::min<uint64_t>((work + 255) / 256, 148u * 32u); };
  F one = F::one();
  if (!inverse && coset) {
    // FFT: a[i] *= u^i (DIF, natural input) or u^bitrev(i) (DIT, bit-reversed input)   fft.go:44-86
    k_fft_scale<P><<<grid(n), 256, 0, st>>>(a, n, d->logn, pw, 1, decimation == 0 /*DIT*/, one, 0);
  }
  if (n > 1) {
    const uint32_t tile = (uint32_t)std::min<uint64_t>(n, TILE);
    const size_t smem = (size_t)tile * sizeof(F);
    if (decimation == 1) {  // DIF: large strides first, then the tile kernel
      for (uint64_t h = half; h >= tile; h >>= 1) k_fft_dif_stage<P><<<grid(half), 256, 0, st>>>(a, tw, half, h, half / h);
      k_fft_tile<P, true><<<(unsigned)(n / tile), tile / 2, smem, st>>>(a, tw, n, tile);
    } else {                // DIT: the tile kernel first, then growing strides
      k_fft_tile<P, false><<<(unsigned)(n / tile), tile / 2, smem, st>>>(a, tw, n, tile);
      for (uint64_t h = tile; h <= half; h <<= 1) k_fft_dit_stage<P><<<grid(half), 256, 0, st>>>(a, tw, half, h, half / h);
    }
  }
  if (inverse) {
    // FFTInverse: scale by CardinalityInv, and on a coset by u^-i (DIT, natural output) or u^-bitrev(i) (DIF)
    F ci;
    memcpy(ci.l, d->consts[2], 32);
    k_fft_scale<P><<<grid(n), 256, 0, st>>>(a, n, d->logn, pw + 64, coset ? 1 : 0, decimation == 1 /*DIF*/, ci, 1);
  }
  CK(cudaGetLastError());
  return GMSM_OK;
}

extern "C" gmsm_fft_domain_t* gmsm_fft_domain_create(int fr_field, uint64_t m, const uint64_t* shift, int device) {
  if (fr_field < 0 || fr_field > 2) { set_err(GMSM_EINVAL, "unknown scalar field %d", fr_field); return nullptr; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { set_err(GMSM_ENODEV, "no CUDA device (%s); this engine has no CPU fallback", cudaGetErrorString(e)); return nullptr; }
  if (device < 0 || device >= ndev) { set_err(GMSM_EINVAL, "device %d out of range", device); return nullptr; }
  cudaSetDevice(device);
  const FrConsts& fc = fr_field == 0 ? FR_BN254 : (fr_field == 1 ? FR_BLS12381 : FR_BLS12377);
  uint64_t x = 1;
  int logn = 0;
  while (x < m) { x <<= 1; logn++; }   // ecc.NextPowerOfTwo(m)
  if (logn > fc.max_order) {
    set_err(GMSM_EINVAL, "m (%llu) is too big: the required root of unity does not exist", (unsigned long long)m);  // generator.go:29
    return nullptr;
  }
  gmsm_fft_domain* d = new gmsm_fft_domain();
  d->field = fr_field; d->device = device; d->n = x; d->logn = logn;
  int rc = fr_field == 0 ? domain_build<bn254_fr>(d, fc, shift)
                         : (fr_field == 1 ? domain_build<bls12381_fr>(d, fc, shift) : domain_build<bls12377_fr>(d, fc, shift));
  if (rc == GMSM_OK && cudaMalloc(&d->d_buf, x * 32) != cudaSuccess) rc = set_err(GMSM_ENOMEM, "cudaMalloc(%llu) failed", (unsigned long long)(x * 32));
  if (rc != GMSM_OK) { cudaFree(d->d_tw); cudaFree(d->d_tw_inv); cudaFree(d->d_pw); cudaFree(d->d_buf); delete d; return nullptr; }
  return d;
}

extern "C" void gmsm_fft_domain_free(gmsm_fft_domain_t* d) {
  if (!d) return;
  cudaSetDevice(d->device);
  cudaFree(d->d_tw); cudaFree(d->d_tw_inv); cudaFree(d->d_pw); cudaFree(d->d_buf);
  delete d;
}

extern "C" uint64_t gmsm_fft_domain_cardinality(const gmsm_fft_domain_t* d) { return d ? d->n : 0; }

extern "C" int gmsm_fft_domain_constants(const gmsm_fft_domain_t* d, uint64_t out[20]) {
  if (!d) return set_err(GMSM_EINVAL, "null domain");
  memcpy(out, d->consts, sizeof(d->consts));
  return GMSM_OK;
}

// unlocked dispatcher: callers hold d->mu
static int fft_dispatch(gmsm_fft_domain_t* d, void* d_a, int inverse, int decimation, int coset, cudaStream_t st) {
  CK(cudaSetDevice(d->device));
  if (d->field == 0) return run_fft<bn254_fr>(d, d_a, inverse, decimation, coset, st);
  if (d->field == 1) return run_fft<bls12381_fr>(d, d_a, inverse, decimation, coset, st);
  return run_fft<bls12377_fr>(d, d_a, inverse, decimation, coset, st);
}

extern "C" int gmsm_fft_device(gmsm_fft_domain_t* d, void* d_a, size_t n, int inverse, int decimation, int coset, void* stream) {
  if (!d) return set_err(GMSM_EINVAL, "null domain");
  if (n != d->n) return set_err(GMSM_EINVAL, "len(a) = %zu must equal the domain cardinality %llu", n, (unsigned long long)d->n);
  if (decimation != 0 && decimation != 1) return set_err(GMSM_EINVAL, "not implemented");  // fft.go:108
  std::lock_guard<std::mutex> lk(d->mu);
  return fft_dispatch(d, d_a, inverse, decimation, coset, (cudaStream_t)stream);
}

// host vector in, host vector out: the domain's staging buffer d_buf is shared by all callers, so the domain mutex is held
// across the whole H2D -> transform -> D2H sequence (two concurrent calls used to interleave on the buffer)
static int fft_host(gmsm_fft_domain_t* d, uint64_t* a, size_t n, int inverse, int decimation, int coset) {
  if (!d) return set_err(GMSM_EINVAL, "null domain");
  if (n != d->n) return set_err(GMSM_EINVAL, "len(a) = %zu must equal the domain cardinality %llu", n, (unsigned long long)d->n);
  if (decimation != 0 && decimation != 1) return set_err(GMSM_EINVAL, "not implemented");  // fft.go:108
  std::lock_guard<std::mutex> lk(d->mu);
  CK(cudaSetDevice(d->device));
  CK(cudaMemcpy(d->d_buf, a, n * 32, cudaMemcpyHostToDevice));
  if (int rc = fft_dispatch(d, d->d_buf, inverse, decimation, coset, nullptr)) return rc;
  CK(cudaMemcpy(a, d->d_buf, n * 32, cudaMemcpyDeviceToHost));
  return GMSM_OK;
}
extern "C" int gmsm_fft(gmsm_fft_domain_t* d, uint64_t* a, size_t n, int decimation, int coset) { return fft_host(d, a, n, 0, decimation, coset); }
extern "C" int gmsm_fft_inverse(gmsm_fft_domain_t* d, uint64_t* a, size_t n, int decimation, int coset) { return fft_host(d, a, n, 1, decimation, coset); }

extern "C" int gmsm_fft_bit_reverse_device(gmsm_fft_domain_t* d, void* d_a, size_t n, void* stream) {
  if (!d) return set_err(GMSM_EINVAL, "null domain");
  if (n != d->n) return set_err(GMSM_EINVAL, "len(a) must be the domain cardinality");
  std::lock_guard<std::mutex> lk(d->mu);
  CK(cudaSetDevice(d->device));
  unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, 148u * 32u);
  if (d->field == 0) k_fft_bit_reverse<bn254_fr><<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<Fp<bn254_fr>*>(d_a), n, d->logn);
  else if (d->field == 1) k_fft_bit_reverse<bls12381_fr><<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<Fp<bls12381_fr>*>(d_a), n, d->logn);
  else k_fft_bit_reverse<bls12377_fr><<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<Fp<bls12377_fr>*>(d_a), n, d->logn);
  CK(cudaGetLastError());
  return GMSM_OK;
}
