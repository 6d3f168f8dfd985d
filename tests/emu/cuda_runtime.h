// TEST INFRASTRUCTURE ONLY -- a stand-in for <cuda_runtime.h> that lets g++ compile gnark-crypto_b200/csrc/kernels.cuh
// and run its kernels one thread at a time on the CPU (tests/emu/emu_engine.cpp, tests/test_emu_kernels.py).  It is
// found instead of the real header because the emulation build puts tests/emu first on the include path; nothing
// in the product (libgmsm.so) ever sees it, and the product has no CPU path.
//
// What is emulated: the execution-space qualifiers (dropped), threadIdx / blockIdx / blockDim / gridDim (set by the
// launcher), uint4 / dim3, atomicAdd / atomicSub (sequential, so plain read-modify-write), __ldg, __clz,
// __funnelshift_r, __syncthreads (a no-op: kernels that NEED a barrier or warp shuffles -- the three scan kernels --
// are not run; the emulated engine scans on the host instead; k_finalize only needs lane 0 to run last).
#pragma once
#include <cassert>
#include <cstddef>
#include <cstdint>

#define __global__
#define __device__
#define __host__
#define __noinline__ __attribute__((noinline))
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static thread_local

struct uint4 {
  uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline uint32_t atomicSub(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o - v; return o; }
static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) {
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> (shift & 31u));
}
static inline void __syncthreads() {}
static inline uint32_t __shfl_up_sync(uint32_t, uint32_t, int) {
  assert(!"warp shuffles are not emulated: the scan kernels are replaced by a host scan");
  return 0;
}

// emulated launch: every thread of every block in turn; lanes in DESCENDING order inside a block, so that the
// "lane 0 continues after the barrier" pattern of k_finalize sees the other lanes' work.  A GPU runs blocks in no
// particular order (which decides, e.g., the order of the entries inside a bucket after the atomic scatter):
// emu_block_order = 0 ascending, 1 descending, >= 2 a pseudo-random permutation seeded by it.
inline unsigned emu_block_order = 0;
template <class K, class... A>
static inline void emu_launch(K kernel, dim3 grid, unsigned block, A... args) {
  gridDim = grid;
  blockDim = dim3(block);
  const uint64_t nb = (uint64_t)grid.x * grid.y;
  // permutation i -> (a*i + b) mod nb with gcd(a, nb) = 1
  uint64_t a = 1, b = 0;
  if (emu_block_order == 1) { a = nb - 1 ? nb - 1 : 1; b = nb - 1; }
  if (emu_block_order >= 2 && nb > 2) {
    a = (0x9E3779B97F4A7C15ull * emu_block_order) % nb;
    auto gcd = [](uint64_t x, uint64_t y) { while (y) { uint64_t t = x % y; x = y; y = t; } return x; };
    while (a == 0 || gcd(a, nb) != 1) a = (a + 1) % nb;
    b = (0xD1B54A32D192ED03ull * emu_block_order) % nb;
  }
  for (uint64_t i = 0; i < nb; i++) {
    const uint64_t k = (a * i + b) % nb;
    blockIdx = dim3((unsigned)(k % grid.x), (unsigned)(k / grid.x), 0);
    for (unsigned t = block; t-- > 0;) {
      threadIdx = dim3(t, 0, 0);
      kernel(args...);
    }
  }
}
