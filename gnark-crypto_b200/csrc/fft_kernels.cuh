// Device side of the Fr FFT (next-row N3): the kernels of fft.cu, in a header of their own so that the CPU kernel
// emulation of tests/emu/ can compile and run them too (tests/test_emu_kernels.py); fft.cu includes this file verbatim.
// See fft.cu for the reference citations (ecc/bn254/fr/fft/fft.go:31-190, 195+, 262+, bitreverse.go:17-42).
#pragma once
#include <cuda_runtime.h>

#include "field.cuh"

using namespace gmsm;

namespace {

constexpr int TILE_LOG = 10;            // stages fused in shared memory: 2^10 elements x 32 B = 32 KB per block
constexpr int TILE = 1 << TILE_LOG;

template <class T>
__device__ __forceinline__ T ldv(const T* p) {
  T r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) {
    uint4 v = s[i];
    w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
  }
  return r;
}
template <class T>
__device__ __forceinline__ void stv(T* p, const T& r) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&r);
  uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

// tw[t] = w^t for t < count, from pw[k] = w^(2^k)
template <class P>
__global__ void k_fft_powers(Fp<P>* __restrict__ tw, uint64_t count, const Fp<P>* __restrict__ pw, int nbits) {
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (uint64_t)gridDim.x * blockDim.x) {
    Fp<P> acc = Fp<P>::one();
    for (int k = 0; k < nbits; k++)
      if ((t >> k) & 1ull) acc = fp_mul(acc, ldv(pw + k));
    stv(tw + t, acc);
  }
}

// one DIF stage with half-size h >= TILE: (x, y) -> (x + y, (x - y) * w^(j * stride))
template <class P>
__global__ void k_fft_dif_stage(Fp<P>* __restrict__ a, const Fp<P>* __restrict__ tw, uint64_t half_n, uint64_t h, uint64_t stride) {
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < half_n; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t j = t & (h - 1), blk = t / h;
    const uint64_t i0 = blk * 2 * h + j, i1 = i0 + h;
    Fp<P> x = ldv(a + i0), y = ldv(a + i1);
    stv(a + i0, fp_add(x, y));
    Fp<P> d = fp_sub(x, y);
    stv(a + i1, j ? fp_mul(d, ldv(tw + j * stride)) : d);
  }
}
// one DIT stage with half-size h >= TILE: (x, y) -> (x + y w, x - y w)
template <class P>
__global__ void k_fft_dit_stage(Fp<P>* __restrict__ a, const Fp<P>* __restrict__ tw, uint64_t half_n, uint64_t h, uint64_t stride) {
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < half_n; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t j = t & (h - 1), blk = t / h;
    const uint64_t i0 = blk * 2 * h + j, i1 = i0 + h;
    Fp<P> x = ldv(a + i0), y = ldv(a + i1);
    if (j) y = fp_mul(y, ldv(tw + j * stride));
    stv(a + i0, fp_add(x, y));
    stv(a + i1, fp_sub(x, y));
  }
}

// the stages with half-size < tile (tile = min(n, TILE)) on one tile per block, in shared memory.
// DIF: the LAST log2(tile) stages; DIT: the FIRST log2(tile) stages.  blockDim.x = tile / 2.
template <class P, bool IS_DIF>
__global__ void k_fft_tile(Fp<P>* __restrict__ a, const Fp<P>* __restrict__ tw, uint64_t n, uint32_t tile) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Fp<P>* s = reinterpret_cast<Fp<P>*>(smem_raw);
  const uint64_t base = (uint64_t)blockIdx.x * tile;
  const uint32_t tid = threadIdx.x, half = tile >> 1;
  stv(s + tid, ldv(a + base + tid));
  stv(s + tid + half, ldv(a + base + tid + half));
  __syncthreads();
  if (IS_DIF) {
    for (uint32_t h = half; h >= 1; h >>= 1) {
      const uint32_t j = tid & (h - 1), blk = tid / h;
      const uint32_t i0 = blk * 2 * h + j, i1 = i0 + h;
      Fp<P> x = ldv(s + i0), y = ldv(s + i1);
      Fp<P> d = fp_sub(x, y);
      if (j) d = fp_mul(d, ldv(tw + (uint64_t)j * ((n >> 1) / h)));
      stv(s + i0, fp_add(x, y));
      stv(s + i1, d);
      __syncthreads();
    }
  } else {
    for (uint32_t h = 1; h <= half; h <<= 1) {
      const uint32_t j = tid & (h - 1), blk = tid / h;
      const uint32_t i0 = blk * 2 * h + j, i1 = i0 + h;
      Fp<P> x = ldv(s + i0), y = ldv(s + i1);
      if (j) y = fp_mul(y, ldv(tw + (uint64_t)j * ((n >> 1) / h)));
      stv(s + i0, fp_add(x, y));
      stv(s + i1, fp_sub(x, y));
      __syncthreads();
    }
  }
  stv(a + base + tid, ldv(s + tid));
  stv(a + base + tid + half, ldv(s + tid + half));
}

// a[i] *= scalar * u^(e(i)), e(i) = i or bitrev(i); pw[k] = u^(2^k) (nbits entries); use_shift = 0: scalar only
template <class P>
__global__ void k_fft_scale(Fp<P>* __restrict__ a, uint64_t n, int logn, const Fp<P>* __restrict__ pw, int use_shift, int bitrev,
                            Fp<P> scalar, int use_scalar) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Fp<P> v = ldv(a + i);
    if (use_scalar) v = fp_mul(v, scalar);
    if (use_shift) {
      const uint64_t e = bitrev ? (logn ? (__brevll(i) >> (64 - logn)) : 0ull) : i;
      for (int k = 0; k < logn; k++)
        if ((e >> k) & 1ull) v = fp_mul(v, ldv(pw + k));
    }
    stv(a + i, v);
  }
}

template <class P>
__global__ void k_fft_bit_reverse(Fp<P>* __restrict__ a, uint64_t n, int logn) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = logn ? (__brevll(i) >> (64 - logn)) : 0;
    if (r > i) {
      Fp<P> x = ldv(a + i), y = ldv(a + r);
      stv(a + i, y);
      stv(a + r, x);
    }
  }
}

}  // namespace
