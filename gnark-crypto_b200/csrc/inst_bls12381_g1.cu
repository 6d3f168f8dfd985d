// explicit instantiation of the engine for bls12381_g1
// out-of-line field multiplier: measured faster for the 12-limb and Fp2 groups (instruction-cache bound
// when inlined: profiles/r01_ncu_accumulate_*), slower for bn254 G1 -- see field.cuh
#define GMSM_MUL_NOINLINE 1
#include "engine_impl.cuh"
namespace gmsm {
GMSM_INSTANTIATE(bls12381_g1, vt_bls12381_g1)
}
