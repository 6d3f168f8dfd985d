// explicit instantiation of the engine for secp256k1 G1 (next-row N4, ecc/secp256k1/multiexp.go:32)
// Both moduli of this curve fill all 256 bits (Params::FULL): the field layer takes the carry-aware textbook CIOS
// (field.cuh fp_mul_cios) instead of the even/odd-accumulator form, and the squaring / fused-product variants, which need
// spare top bits, fall back to it.  Out of line: one multiplier body per kernel instead of ten.
#define GMSM_MUL_NOINLINE 1
#define GMSM_ACC_NOPREFETCH 1
#include "engine_impl.cuh"
namespace gmsm {
GMSM_INSTANTIATE(secp256k1_g1, vt_secp256k1_g1)
}
