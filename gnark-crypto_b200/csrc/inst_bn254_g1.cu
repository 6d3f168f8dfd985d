// explicit instantiation of the engine for bn254_g1
// multiplier variants measured on B200 for this group (profiles/r02_ab_multiplier_variants_call2.txt): dedicated squaring + the
// fused two-product y-coordinate (field.cuh) -- bn254 G1 41.4 -> 40.3 ms at 2^24 with 3 blocks per SM (154 registers, no spills;
// at 4 blocks the fused routine spills), bls12-381 G1 102.0 -> 92.1 ms; the Fp2 groups gain nothing (their products are
// Karatsuba triples, no fused form) and keep the plain multiplier.
#ifndef GMSM_SQR_DEDICATED
#define GMSM_SQR_DEDICATED 1
#endif
#ifndef GMSM_DOT2
#define GMSM_DOT2 1
#endif
#ifndef GMSM_ACC_MINBLOCKS_SMALL
#define GMSM_ACC_MINBLOCKS_SMALL 3
#endif
#include "engine_impl.cuh"
namespace gmsm {
GMSM_INSTANTIATE(bn254_g1, vt_bn254_g1)
}
