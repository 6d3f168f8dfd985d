// explicit instantiation of the engine for bn254_g2
// out-of-line field multiplier: measured faster for the 12-limb and Fp2 groups (instruction-cache bound
// when inlined: profiles/r01_ncu_accumulate_*), slower for bn254 G1 -- see field.cuh
#define GMSM_MUL_NOINLINE 1
#include "engine_impl.cuh"
namespace gmsm {
GMSM_INSTANTIATE(bn254_g2, vt_bn254_g2)
}
