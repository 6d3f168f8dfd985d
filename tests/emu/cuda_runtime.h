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
#include <ucontext.h>

#include <algorithm>
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <map>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __noinline__ __attribute__((noinline))
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ thread_local   /* block scope: implies static; `extern __shared__ x[]` works too */

struct uint4 {
  uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct uint2 {
  uint32_t x, y;
};
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
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
static inline unsigned long long __brevll(unsigned long long v) {
  unsigned long long r = 0;
  for (int i = 0; i < 64; i++) r |= ((v >> i) & 1ull) << (63 - i);
  return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) {
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> (shift & 31u));
}
static inline uint32_t atomicOr(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
static inline uint32_t atomicMax(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = std::max(o, v); return o; }
using std::max;
using std::min;
#define __align__(n) __attribute__((aligned(n)))

// ---- cooperative blocks: one fiber (ucontext) per CUDA thread, so that __syncthreads and warp shuffles work -------
// emu_launch_coop runs ONE block at a time; a fiber runs until it blocks in a barrier / shuffle or returns; a barrier
// opens when every live thread of the block (of the warp) has arrived -- exited threads count as arrived, as on the
// hardware.  `__shared__` variables (static thread_local here) are shared by the fibers of the running block.
namespace emu {
struct Block {
  unsigned nthreads = 0, live = 0;
  std::vector<ucontext_t> ctx;
  std::vector<char*> stacks;
  std::vector<char> done;
  ucontext_t sched;
  unsigned cur = 0;
  unsigned bar_arrived = 0, bar_gen = 0;
  unsigned warp_live[32] = {}, warp_arrived[32] = {}, warp_gen[32] = {};
  uint32_t slot[1024] = {};
  std::map<uint64_t, std::pair<unsigned, unsigned>> grp;   // (warp, mask) -> (arrived, generation): barriers of lane groups
  std::function<void()> body;
};
inline Block* g_block = nullptr;   // non-null while a cooperative launch is running
inline void yield() {
  Block& B = *g_block;
  swapcontext(&B.ctx[B.cur], &B.sched);
}
inline void block_barrier() {
  Block& B = *g_block;
  const unsigned gen = B.bar_gen;
  if (++B.bar_arrived >= B.live) { B.bar_arrived = 0; B.bar_gen++; return; }
  while (B.bar_gen == gen) yield();
}
inline void warp_barrier() {
  Block& B = *g_block;
  const unsigned w = B.cur >> 5, gen = B.warp_gen[w];
  if (++B.warp_arrived[w] >= B.warp_live[w]) { B.warp_arrived[w] = 0; B.warp_gen[w]++; return; }
  while (B.warp_gen[w] == gen) yield();
}
inline void thread_exit() {   // an exiting thread may complete a barrier the others are waiting in
  Block& B = *g_block;
  const unsigned w = B.cur >> 5;
  B.done[B.cur] = 1;
  B.live--;
  B.warp_live[w]--;
  if (B.live && B.bar_arrived >= B.live) { B.bar_arrived = 0; B.bar_gen++; }
  if (B.warp_live[w] && B.warp_arrived[w] >= B.warp_live[w]) { B.warp_arrived[w] = 0; B.warp_gen[w]++; }
}
inline void trampoline() {
  g_block->body();
  thread_exit();
  yield();   // never resumed
}
inline uint32_t shfl(uint32_t v, unsigned src_lane, bool valid) {
  Block& B = *g_block;
  const unsigned tid = B.cur, base = tid & ~31u;
  B.slot[tid] = v;
  warp_barrier();
  const uint32_t r = (valid && base + src_lane < B.nthreads) ? B.slot[base + src_lane] : v;
  warp_barrier();
  return r;
}
// shuffle among the lanes named in `mask` only (the quads of quad.cuh): a barrier over the live lanes of the mask, so that
// groups of one warp may run different code paths, as independent thread scheduling allows on the hardware
inline void group_barrier(unsigned w, unsigned /*key*/, uint32_t mask) {
  Block& B = *g_block;
  const unsigned base = w << 5;
  unsigned members = 0;
  for (unsigned l = 0; l < 32; l++)
    if (((mask >> l) & 1u) && base + l < B.nthreads && !B.done[base + l]) members++;
  const uint64_t id = ((uint64_t)w << 32) | mask;     // distinct masks (a quad, a set of peers, the whole warp) never share a counter
  const unsigned gen = B.grp[id].second;
  if (++B.grp[id].first >= members) { B.grp[id].first = 0; B.grp[id].second++; return; }
  while (B.grp[id].second == gen) yield();
}
inline uint32_t shfl_masked(uint32_t mask, uint32_t v, unsigned src_lane) {
  Block& B = *g_block;
  const unsigned tid = B.cur, base = tid & ~31u, w = tid >> 5, key = (unsigned)__builtin_ctz(mask);
  assert(((mask >> (tid & 31u)) & 1u) && "calling lane must be named in the mask");
  B.slot[tid] = v;
  group_barrier(w, key, mask);
  const uint32_t r = (base + src_lane < B.nthreads) ? B.slot[base + src_lane] : v;
  group_barrier(w, key, mask);
  return r;
}
}  // namespace emu

static inline void __syncthreads() {
  if (emu::g_block) emu::block_barrier();   // sequential launches (emu_launch): no-op, see there
}
static inline uint32_t __shfl_up_sync(uint32_t, uint32_t v, int delta) {
  assert(emu::g_block && "warp shuffles need emu_launch_coop");
  const unsigned lane = emu::g_block->cur & 31u;
  return emu::shfl(v, lane - (unsigned)delta, lane >= (unsigned)delta);
}
static inline uint32_t __shfl_xor_sync(uint32_t, uint32_t v, int mask) {
  assert(emu::g_block && "warp shuffles need emu_launch_coop");
  const unsigned lane = emu::g_block->cur & 31u;
  return emu::shfl(v, lane ^ (unsigned)mask, true);
}

static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
// lanes of `mask` (all live lanes named must call) whose key equals the caller's
static inline unsigned __match_any_sync(uint32_t mask, uint32_t key) {
  assert(emu::g_block && "__match_any_sync needs emu_launch_coop");
  emu::Block& B = *emu::g_block;
  const unsigned tid = B.cur, base = tid & ~31u, w = tid >> 5, gk = (unsigned)__builtin_ctz(mask);
  B.slot[tid] = key;
  emu::group_barrier(w, gk, mask);
  unsigned peers = 0;
  for (unsigned l = 0; l < 32; l++)
    if (((mask >> l) & 1u) && base + l < B.nthreads && !B.done[base + l] && B.slot[base + l] == key) peers |= 1u << l;
  emu::group_barrier(w, gk, mask);
  return peers;
}
// full-warp vote (every live lane of the warp must call)
static inline unsigned __ballot_sync(uint32_t mask, int pred) {
  assert(emu::g_block && "__ballot_sync needs emu_launch_coop");
  emu::Block& B = *emu::g_block;
  const unsigned tid = B.cur, base = tid & ~31u, w = tid >> 5;
  B.slot[tid] = pred ? 1u : 0u;
  emu::group_barrier(w, 0, mask);
  unsigned r = 0;
  for (unsigned l = 0; l < 32; l++)
    if (((mask >> l) & 1u) && base + l < B.nthreads && !B.done[base + l] && B.slot[base + l]) r |= 1u << l;
  emu::group_barrier(w, 0, mask);
  return r;
}
static inline void __syncwarp(uint32_t mask = 0xffffffffu) {
  assert(emu::g_block && "__syncwarp needs emu_launch_coop");
  emu::group_barrier(emu::g_block->cur >> 5, (unsigned)__builtin_ctz(mask), mask);
}
static inline uint32_t __shfl_sync(uint32_t mask, uint32_t v, int src, int width = 32) {
  assert(emu::g_block && "warp shuffles need emu_launch_coop");
  const unsigned lane = emu::g_block->cur & 31u;
  const unsigned seg = lane & ~((unsigned)width - 1u);
  return emu::shfl_masked(mask, v, seg + ((unsigned)src & ((unsigned)width - 1u)));
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

// cooperative launch: fibers, blocks one after the other (in emu_block_order), threads of a block interleaved at barriers
template <class K, class... A>
static inline void emu_launch_coop(K kernel, dim3 grid, unsigned block, A... args) {
  gridDim = grid;
  blockDim = dim3(block);
  const uint64_t nb = (uint64_t)grid.x * grid.y;
  constexpr size_t STACK = 512 * 1024;
  emu::Block B;
  B.nthreads = block;
  B.ctx.resize(block);
  B.stacks.resize(block, nullptr);
  B.done.resize(block);
  for (unsigned t = 0; t < block; t++) B.stacks[t] = (char*)std::malloc(STACK);
  B.body = [&]() { kernel(args...); };
  for (uint64_t i = 0; i < nb; i++) {
    const uint64_t k = (emu_block_order == 1) ? nb - 1 - i : i;
    blockIdx = dim3((unsigned)(k % grid.x), (unsigned)(k / grid.x), 0);
    B.live = block;
    B.bar_arrived = 0;
    B.grp.clear();
    for (unsigned w = 0; w < 32; w++) { B.warp_arrived[w] = 0; B.warp_live[w] = (block > w * 32) ? std::min(32u, block - w * 32) : 0; }
    for (unsigned t = 0; t < block; t++) {
      B.done[t] = 0;
      getcontext(&B.ctx[t]);
      B.ctx[t].uc_stack.ss_sp = B.stacks[t];
      B.ctx[t].uc_stack.ss_size = STACK;
      B.ctx[t].uc_link = nullptr;
      makecontext(&B.ctx[t], (void (*)())emu::trampoline, 0);
    }
    emu::g_block = &B;
    while (B.live) {
      for (unsigned t = 0; t < block; t++) {
        if (B.done[t]) continue;
        B.cur = t;
        threadIdx = dim3(t, 0, 0);
        swapcontext(&B.sched, &B.ctx[t]);
      }
    }
    emu::g_block = nullptr;
  }
  for (unsigned t = 0; t < block; t++) std::free(B.stacks[t]);
}
