// explicit instantiation of the engine for secp256k1 G1 (next-row N4, ecc/secp256k1/multiexp.go:32)
// Both moduli of this curve fill all 256 bits (Params::FULL): the multiplier keeps the two carries per row that a spare top
// bit would make zero (field.cuh fp_mul_inline), additions hand their carry-out to the final subtraction, and the squaring /
// fused-product variants -- which need two spare bits -- fall back to the multiplier.  Otherwise the 8-limb kernel shape of
// bn254 G1: inlined multiplier, no software prefetch, 4 blocks per SM.
#ifndef GMSM_ACC_NOPREFETCH
#define GMSM_ACC_NOPREFETCH 1
#endif
#ifndef GMSM_ACC_MINBLOCKS_SMALL
#define GMSM_ACC_MINBLOCKS_SMALL 4
#endif
#include "engine_impl.cuh"
namespace gmsm {
GMSM_INSTANTIATE(secp256k1_g1, vt_secp256k1_g1)
}
