// explicit instantiation of the engine for bn254_g2
// out-of-line field multiplier: measured faster for the 12-limb and Fp2 groups (instruction-cache bound
// when inlined: profiles/r01_ncu_accumulate_*), slower for bn254 G1 -- see field.cuh
#define GMSM_MUL_NOINLINE 1
// Per-group build choices, each measured on B200 (profiles/r02_ab_fp2_dot2_call6.txt, r02_ab_fp2_lazy_vs_dot2_call7.txt,
// r02_ab_noprefetch_call7.txt):
//  * no software prefetch of the next point (frees 32 .. 48 registers of a 255-register kernel): bn254 G2 37.7 -> 35.9 ms,
//    bls12-381 G2 21.3 -> 20.7 ms
//  * bn254 G2: the Fp2 product as two fused two-product reductions (fp2.cuh, GMSM_FP2_DOT2): 37.7 -> 34.2 ms; the lazy-reduction
//    product over the separated wide product / REDC routines (GMSM_FP2_LAZY, 336 instead of 408 IMAD.WIDE) measured 35.2 ms
//    (bls12-381 G2: 27.2 ms against 21.3) and is not used
#ifndef GMSM_SQR_DEDICATED
#define GMSM_SQR_DEDICATED 1
#endif
#ifndef GMSM_DOT2
#define GMSM_DOT2 1
#endif
#ifndef GMSM_FP2_DOT2
#define GMSM_FP2_DOT2 1
#endif
#ifndef GMSM_ACC_NOPREFETCH
#define GMSM_ACC_NOPREFETCH 1
#endif
// the y-coordinate of a point addition as two four-product fused reductions (field.cuh fp_dot4): 40.4 -> 39.3 ms at 2^22
// (profiles/r02_ab_dot4_call10.txt; the 12-limb G2 groups spill with it and stay on the two-product form)
#ifndef GMSM_DOT4
#define GMSM_DOT4 1
#endif
#include "engine_impl.cuh"
namespace gmsm {
GMSM_INSTANTIATE(bn254_g2, vt_bn254_g2)
}
