// explicit instantiation of the engine for bn254_g1
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
#ifndef GMSM_ACC_NOPREFETCH
#define GMSM_ACC_NOPREFETCH 1
#endif
#ifndef GMSM_ACC_MINBLOCKS_SMALL
#define GMSM_ACC_MINBLOCKS_SMALL 4
#endif
#include "engine_impl.cuh"
namespace gmsm {
GMSM_INSTANTIATE(bn254_g1, vt_bn254_g1)
}
