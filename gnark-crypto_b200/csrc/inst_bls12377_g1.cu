// explicit instantiation of the engine for bls12377_g1 (next-row N4)
#define GMSM_MUL_NOINLINE 1
#include "engine_impl.cuh"
namespace gmsm {
GMSM_INSTANTIATE(bls12377_g1, vt_bls12377_g1)
}
