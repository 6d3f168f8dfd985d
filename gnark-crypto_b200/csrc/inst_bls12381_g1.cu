// explicit instantiation of the engine for bls12381_g1
// out-of-line field multiplier: measured faster for the 12-limb and Fp2 groups (instruction-cache bound
// when inlined: profiles/r01_ncu_accumulate_*), slower for bn254 G1 -- see field.cuh
#define GMSM_MUL_NOINLINE 1
// Per-group build choices, each measured on B200 (profiles/r02_ab_multiplier_variants_call2.txt, r02_ab_noprefetch_call7.txt):
//  * dedicated squaring + fused two-product y-coordinate (field.cuh): bn254 G1 41.4 -> 40.3 ms, bls12-381 G1 102.0 -> 92.1 ms
//  * bn254 G1 only: no software prefetch of the next point (the gather latency is covered by the other warps), which frees its
//    16 registers: 128 registers, 4 blocks / SM, no spills: accumulate 40.9 -> 39.4 ms at 2^24 (for the 12-limb G1 groups the
//    prefetch stays: 90.1 vs 91.0 ms)
#ifndef GMSM_SQR_DEDICATED
#define GMSM_SQR_DEDICATED 1
#endif
#ifndef GMSM_DOT2
#define GMSM_DOT2 1
#endif
#include "engine_impl.cuh"
namespace gmsm {
GMSM_INSTANTIATE(bls12381_g1, vt_bls12381_g1)
}
