// Batch-affine bucket accumulation (the GPU restatement of the reference's processChunkG1BatchAffine,
// ecc/bn254/multiexp_affine.go:24-231 + batchAddG1Affine g1.go:1122-1182; G2: multiexp_affine.go:351+,
// g2.go:1081-1141).
//
// The reference keeps affine buckets and adds up to `batchSize` (80..640) independent (bucket, point)
// pairs with ONE shared field inversion (Montgomery's trick), queueing conflicting adds.  On the GPU the
// independent pairs come from a balanced tree over the bucket-ordered entry list:
//
//   level 0 : the M entries, grouped by bucket (run of bucket b = len_0[b] entries)
//   level l+1: every run is halved: out[k] = in[2k] + in[2k+1] (or in[2k] alone when the run is odd),
//              len_{l+1}[b] = ceil(len_l[b] / 2);  after ceil(log2(max len)) levels one affine point per
//              non-empty bucket is left -- the bucket sum.
//
// All additions of a level are independent, so a level shares a single inversion:
//   forward  : every thread walks B consecutive output slots, computes the denominators
//              (x2-x1, or 2y for a doubling), stores the running product BEFORE each one (pref[o]) and
//              its total product (totals[t]);
//   scan     : hierarchical product scans over the thread totals give each thread the inverse of its own
//              total:  inv(tot_t) = inv(prod all) * prod_{u<t} tot_u * prod_{u>t} tot_u   (one Fermat
//              inversion per level for the whole GPU);
//   backward : every thread walks its slots in reverse: inv_i = invrun * pref[i]; invrun *= den_i;
//              lambda = num * inv_i; x3 = lambda^2 - x1 - x2; y3 = lambda (x1 - x3) - y1.
// Cost per addition: 5 M + 1 S (+ ~0.2 amortised) instead of 8 M + 2 S for the extended-Jacobian mixed
// add -- the same saving the reference gets on the CPU.
//
// Exceptional pairs never enter the shared product (their denominator is treated as 1):
//   no partner / P = inf / Q = inf -> copy;  P = -Q (or equal x, unrelated y) -> inf;
//   P = Q -> doubling with denominator 2y, numerator 3x^2 (y = 0 cannot occur on these curves; if it
//   does, off-curve input, the result is defined as inf).
#pragma once
#include "kernels.cuh"

namespace gmsm {

enum : int { AFF_COPY_P = 0, AFF_COPY_Q = 1, AFF_INF = 2, AFF_ADD = 3, AFF_DBL = 4 };

template <class F>
GMSM_D int aff_classify(const Affine<F>& P, const Affine<F>& Q, bool has_q, F& den) {
  if (!has_q) return AFF_COPY_P;
  if (P.is_inf()) return AFF_COPY_Q;
  if (Q.is_inf()) return AFF_COPY_P;
  if (P.x == Q.x) {
    if (P.y == Q.y && !P.y.is_zero()) {
      den = f_dbl(P.y);
      return AFF_DBL;
    }
    return AFF_INF;
  }
  den = f_sub(Q.x, P.x);
  return AFF_ADD;
}

// input element i of a level: level 0 reads the entry list and gathers the base point (negated for
// negative digits: subMixed, g1.go:878-930); higher levels read the previous level's output buffer
template <class G, bool L0>
GMSM_D Affine<typename G::F> aff_get(const Affine<typename G::F>* __restrict__ bases, const uint32_t* __restrict__ entries,
                                     const Affine<typename G::F>* __restrict__ src, uint32_t i) {
  using F = typename G::F;
  if (L0) {
    uint32_t e = __ldg(entries + i);
    Affine<F> a = load_vec_ro(bases + (e >> 1));
    if (e & 1u) a.y = f_neg(a.y);
    return a;
  } else {
    return load_vec(src + i);
  }
}

// counts[b] = ceil(len_0[b] / 2^level) for b < nb; counts[nb] = 0 (so the exclusive scan has nb+1 entries)
static __global__ void k_aff_level_counts(const uint32_t* __restrict__ offsets0, uint32_t nb, int level,
                                          uint32_t* __restrict__ counts) {
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b <= nb; b += gridDim.x * blockDim.x) {
    uint32_t v = 0;
    if (b < nb) {
      uint32_t len = offsets0[b + 1] - offsets0[b];
      v = (uint32_t)(((uint64_t)len + ((1ull << level) - 1ull)) >> level);
    }
    counts[b] = v;
  }
}

static __global__ void k_aff_max_len(const uint32_t* __restrict__ offsets0, uint32_t nb, uint32_t* __restrict__ out_max) {
  uint32_t m = 0;
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x)
    m = max(m, offsets0[b + 1] - offsets0[b]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out_max, m);
}

// ---- slot geometry --------------------------------------------------------------------------------
// Output slots of a level are dealt to threads WARP-INTERLEAVED: warp w owns slots
// [w*32*B, (w+1)*32*B) and lane L takes slot w*32*B + i*32 + L at iteration i.  In every iteration the
// 32 lanes touch 32 consecutive slots, so the pref / output stores, the entry reads and (levels >= 1)
// the input reads are fully coalesced; only the level-0 base-point gathers are random (inherently).
// Each lane still runs its own Montgomery batch over its B slots.
struct AffCursor {   // a lane's position in the bucket structure
  uint32_t b, bstart_n, bend_n, bstart_l, blen_l;
};

GMSM_D void aff_cursor_set(AffCursor& c, uint32_t b, const uint32_t* __restrict__ off_l, const uint32_t* __restrict__ off_n) {
  c.b = b;
  c.bstart_n = off_n[b];
  c.bend_n = off_n[b + 1];
  c.bstart_l = off_l[b];
  c.blen_l = off_l[b + 1] - c.bstart_l;
}
// move to the bucket containing output slot o (o may be before or after the current bucket)
GMSM_D void aff_cursor_seek(AffCursor& c, uint32_t o, const uint32_t* __restrict__ off_l, const uint32_t* __restrict__ off_n,
                            uint32_t nb) {
  if (o >= c.bstart_n && o < c.bend_n) return;
  uint32_t b = c.b;
  bool found = false;
  if (o >= c.bend_n) {
#pragma unroll 1
    for (int s = 0; s < 4; s++) {  // short forward walk, then binary search
      b++;
      if (b >= nb) break;
      if (off_n[b + 1] > o) { found = (off_n[b] <= o); break; }
    }
  }
  if (!found) b = upper_bound_u32(off_n, nb + 1, o) - 1u;
  aff_cursor_set(c, b, off_l, off_n);
}

template <class G, bool L0>
struct AffPair {
  Affine<typename G::F> P, Q;
  bool has_q;
};

template <class G, bool L0>
GMSM_D void aff_load_pair(AffPair<G, L0>& pr, const AffCursor& c, uint32_t o, const Affine<typename G::F>* __restrict__ bases,
                          const uint32_t* __restrict__ entries, const Affine<typename G::F>* __restrict__ src) {
  const uint32_t k = o - c.bstart_n;
  const uint32_t i0 = c.bstart_l + 2u * k;
  pr.has_q = (2u * k + 1u < c.blen_l);
  pr.P = aff_get<G, L0>(bases, entries, src, i0);
  pr.Q = pr.has_q ? aff_get<G, L0>(bases, entries, src, i0 + 1u) : pr.P;
}

// ---- forward pass -------------------------------------------------------------------------------
template <class G, bool L0>
__global__ void __launch_bounds__(128)
k_aff_forward(const Affine<typename G::F>* __restrict__ bases, const uint32_t* __restrict__ entries,
              const Affine<typename G::F>* __restrict__ src, const uint32_t* __restrict__ off_l,
              const uint32_t* __restrict__ off_n, uint32_t nb, uint32_t B, uint32_t T,
              typename G::F* __restrict__ pref, typename G::F* __restrict__ totals) {
  using F = typename G::F;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint32_t m_out = off_n[nb];
  const uint32_t lane = t & 31u;
  const uint64_t first64 = (uint64_t)(t >> 5) * 32u * B + lane;
  F run = F::one();
  if (first64 < m_out) {
    uint32_t o = (uint32_t)first64;
    AffCursor cur;
    aff_cursor_set(cur, upper_bound_u32(off_n, nb + 1, o) - 1u, off_l, off_n);
    AffPair<G, L0> pr;
    aff_load_pair<G, L0>(pr, cur, o, bases, entries, src);
    for (uint32_t i = 0; i < B; i++) {
      // software pipeline: locate + load the next slot's pair before this slot's arithmetic
      const uint64_t on64 = (uint64_t)o + 32u;
      const bool has_next = (i + 1 < B) && (on64 < m_out);
      AffPair<G, L0> nx;
      if (has_next) {
        aff_cursor_seek(cur, (uint32_t)on64, off_l, off_n, nb);
        aff_load_pair<G, L0>(nx, cur, (uint32_t)on64, bases, entries, src);
      }
      F den;
      const int kind = aff_classify(pr.P, pr.Q, pr.has_q, den);
      if (kind >= AFF_ADD) {
        store_vec(pref + o, run);
        run = f_mul(run, den);
      }
      if (!has_next) break;
      pr = nx;
      o = (uint32_t)on64;
    }
  }
  store_vec(totals + t, run);
}

// ---- product scans over the thread totals ---------------------------------------------------------
static constexpr int PSCAN_THREADS = 256;
static constexpr int PSCAN_ITEMS = 4;
static constexpr int PSCAN_TILE = PSCAN_THREADS * PSCAN_ITEMS;

// inclusive Kogge-Stone product scan over one value per thread (blockDim.x == PSCAN_THREADS)
template <class F>
GMSM_D F block_incl_prod_scan(F v, F* smem) {
  const int tid = threadIdx.x;
  store_vec(smem + tid, v);
  __syncthreads();
  for (int d = 1; d < PSCAN_THREADS; d <<= 1) {
    F cur = load_vec(smem + tid);
    if (tid >= d) {
      F lo = load_vec(smem + tid - d);
      cur = f_mul(lo, cur);
    }
    __syncthreads();
    store_vec(smem + tid, cur);
    __syncthreads();
  }
  F r = load_vec(smem + tid);
  return r;
}

// S1: per tile of PSCAN_TILE totals: ps[t] = (product of the tile's totals before t) * (product after t);
//     BP[tile] = product of the whole tile
template <class G>
__global__ void __launch_bounds__(PSCAN_THREADS)
k_aff_scan_tiles(const typename G::F* __restrict__ totals, uint32_t T, typename G::F* __restrict__ ps,
                 typename G::F* __restrict__ BP) {
  using F = typename G::F;
  __shared__ __align__(16) unsigned char smem_raw[sizeof(F) * PSCAN_THREADS];
  F* smem = reinterpret_cast<F*>(smem_raw);
  const int tid = threadIdx.x;
  const uint32_t base = blockIdx.x * PSCAN_TILE + tid * PSCAN_ITEMS;
  F x[PSCAN_ITEMS];
#pragma unroll
  for (int j = 0; j < PSCAN_ITEMS; j++) x[j] = (base + j < T) ? load_vec(totals + base + j) : F::one();
  // thread-local exclusive prefix / suffix products
  F lp[PSCAN_ITEMS], ls[PSCAN_ITEMS];
  lp[0] = F::one();
#pragma unroll
  for (int j = 1; j < PSCAN_ITEMS; j++) lp[j] = f_mul(lp[j - 1], x[j - 1]);
  F tp = f_mul(lp[PSCAN_ITEMS - 1], x[PSCAN_ITEMS - 1]);
  ls[PSCAN_ITEMS - 1] = F::one();
#pragma unroll
  for (int j = PSCAN_ITEMS - 2; j >= 0; j--) ls[j] = f_mul(ls[j + 1], x[j + 1]);
  // block scans: prefix over tid, suffix over reversed tid
  F incl = block_incl_prod_scan(tp, smem);
  F excl_pre = (tid > 0) ? load_vec(smem + tid - 1) : F::one();
  F tile_prod = load_vec(smem + PSCAN_THREADS - 1);
  __syncthreads();
  // reversed: thread tid contributes at position PSCAN_THREADS-1-tid
  {
    store_vec(smem + (PSCAN_THREADS - 1 - tid), tp);
    __syncthreads();
    F v = load_vec(smem + tid);
    __syncthreads();
    (void)block_incl_prod_scan(v, smem);
  }
  const int rpos = PSCAN_THREADS - 1 - tid;  // my position in the reversed order
  F excl_suf = (rpos > 0) ? load_vec(smem + rpos - 1) : F::one();
  (void)incl;
#pragma unroll
  for (int j = 0; j < PSCAN_ITEMS; j++) {
    if (base + j < T) {
      F a = f_mul(excl_pre, lp[j]);
      F b = f_mul(excl_suf, ls[j]);
      store_vec(ps + base + j, f_mul(a, b));
    }
  }
  if (tid == 0) store_vec(BP + blockIdx.x, tile_prod);
}

// S2 (one block): invBP[i] = inv(BP[i]) = inv(prod all) * prod_{u<i} BP[u] * prod_{u>i} BP[u]
template <class G>
__global__ void __launch_bounds__(PSCAN_THREADS)
k_aff_scan_top(const typename G::F* __restrict__ BP, uint32_t NB, typename G::F* __restrict__ preB,
               typename G::F* __restrict__ invBP) {
  using F = typename G::F;
  __shared__ __align__(16) unsigned char smem_raw[sizeof(F) * PSCAN_THREADS];
  __shared__ __align__(16) unsigned char carry_raw[sizeof(F)];
  F* smem = reinterpret_cast<F*>(smem_raw);
  F* carry = reinterpret_cast<F*>(carry_raw);
  const int tid = threadIdx.x;
  if (tid == 0) store_vec(carry, F::one());
  __syncthreads();
  // forward: exclusive prefix products
  for (uint32_t base = 0; base < NB; base += PSCAN_THREADS) {
    const uint32_t i = base + tid;
    F v = (i < NB) ? load_vec(BP + i) : F::one();
    F c = load_vec(carry);
    (void)block_incl_prod_scan(v, smem);
    F ex = (tid > 0) ? load_vec(smem + tid - 1) : F::one();
    if (i < NB) store_vec(preB + i, f_mul(c, ex));
    F tot = load_vec(smem + PSCAN_THREADS - 1);
    __syncthreads();
    if (tid == 0) store_vec(carry, f_mul(c, tot));
    __syncthreads();
  }
  // one inversion for the whole level
  if (tid == 0) {
    F tot = load_vec(carry);
    store_vec(carry, f_inv(tot));
  }
  __syncthreads();
  const F ginv = load_vec(carry);
  __syncthreads();
  if (tid == 0) store_vec(carry, F::one());
  __syncthreads();
  // backward: exclusive suffix products, tile by tile from the end
  const uint32_t ntiles = (NB + PSCAN_THREADS - 1) / PSCAN_THREADS;
  for (uint32_t tile = ntiles; tile-- > 0;) {
    const uint32_t base = tile * PSCAN_THREADS;
    // reversed position r <-> element base + (PSCAN_THREADS-1-r)
    const uint32_t i_rev = base + (PSCAN_THREADS - 1 - tid);
    F v = (i_rev < NB) ? load_vec(BP + i_rev) : F::one();
    F c = load_vec(carry);
    (void)block_incl_prod_scan(v, smem);
    F ex = (tid > 0) ? load_vec(smem + tid - 1) : F::one();  // product of elements after i_rev within the tile
    if (i_rev < NB) {
      F suf = f_mul(c, ex);
      F pre = load_vec(preB + i_rev);
      store_vec(invBP + i_rev, f_mul(ginv, f_mul(pre, suf)));
    }
    F tot = load_vec(smem + PSCAN_THREADS - 1);
    __syncthreads();
    if (tid == 0) store_vec(carry, f_mul(c, tot));
    __syncthreads();
  }
}

// ---- backward pass ------------------------------------------------------------------------------
template <class G, bool L0>
__global__ void __launch_bounds__(128)
k_aff_backward(const Affine<typename G::F>* __restrict__ bases, const uint32_t* __restrict__ entries,
               const Affine<typename G::F>* __restrict__ src, const uint32_t* __restrict__ off_l,
               const uint32_t* __restrict__ off_n, uint32_t nb, uint32_t B, uint32_t T,
               const typename G::F* __restrict__ pref, const typename G::F* __restrict__ ps,
               const typename G::F* __restrict__ invBP, Affine<typename G::F>* __restrict__ dst) {
  using F = typename G::F;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint32_t m_out = off_n[nb];
  const uint32_t lane = t & 31u;
  const uint64_t first64 = (uint64_t)(t >> 5) * 32u * B + lane;
  if (first64 >= m_out) return;
  const uint32_t first = (uint32_t)first64;
  // number of slots of this lane: i = 0 .. cnt-1 with first + 32 i < m_out
  uint32_t cnt = (m_out - first + 31u) / 32u;
  if (cnt > B) cnt = B;
  F invrun = f_mul(load_vec(invBP + t / PSCAN_TILE), load_vec(ps + t));
  uint32_t o = first + 32u * (cnt - 1u);
  AffCursor cur;
  aff_cursor_set(cur, upper_bound_u32(off_n, nb + 1, o) - 1u, off_l, off_n);
  AffPair<G, L0> pr;
  aff_load_pair<G, L0>(pr, cur, o, bases, entries, src);
  for (uint32_t i = cnt; i-- > 0;) {
    const bool has_next = (i > 0);
    AffPair<G, L0> nx;
    if (has_next) {
      const uint32_t on = o - 32u;
      if (on < cur.bstart_n) aff_cursor_set(cur, upper_bound_u32(off_n, nb + 1, on) - 1u, off_l, off_n);
      aff_load_pair<G, L0>(nx, cur, on, bases, entries, src);
    }
    F den;
    const int kind = aff_classify(pr.P, pr.Q, pr.has_q, den);
    Affine<F> out;
    if (kind >= AFF_ADD) {
      F pf = load_vec(pref + o);
      F inv = f_mul(invrun, pf);
      invrun = f_mul(invrun, den);
      F num;
      if (kind == AFF_ADD) {
        num = f_sub(pr.Q.y, pr.P.y);
      } else {
        F xx = f_sqr(pr.P.x);
        num = f_add(f_dbl(xx), xx);
      }
      F lam = f_mul(num, inv);
      F x3 = f_sub(f_sub(f_sqr(lam), pr.P.x), pr.Q.x);
      out.x = x3;
      out.y = f_sub(f_mul(lam, f_sub(pr.P.x, x3)), pr.P.y);
    } else if (kind == AFF_COPY_P) {
      out = pr.P;
    } else if (kind == AFF_COPY_Q) {
      out = pr.Q;
    } else {
      out = Affine<F>::inf();
    }
    store_vec(dst + o, out);
    if (has_next) {
      pr = nx;
      o -= 32u;
    }
  }
}

// final: one affine point (or none) per bucket -> extended-Jacobian bucket array
template <class G, bool L0>
__global__ void k_aff_to_buckets(const Affine<typename G::F>* __restrict__ bases, const uint32_t* __restrict__ entries,
                                 const Affine<typename G::F>* __restrict__ src, const uint32_t* __restrict__ off_l,
                                 uint32_t nb, XYZZ<typename G::F>* __restrict__ buckets) {
  using F = typename G::F;
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
    XYZZ<F> r = XYZZ<F>::inf();
    if (off_l[b + 1] > off_l[b]) {
      Affine<F> a = aff_get<G, L0>(bases, entries, src, off_l[b]);
      r = xyzz_from_affine(a);
    }
    store_vec(buckets + b, r);
  }
}

}  // namespace gmsm
