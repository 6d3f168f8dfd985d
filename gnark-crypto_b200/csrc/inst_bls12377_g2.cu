// explicit instantiation of the engine for bls12377_g2 (next-row N4; Fp2 with u^2 = -5)
#define GMSM_MUL_NOINLINE 1
#include "engine_impl.cuh"
namespace gmsm {
GMSM_INSTANTIATE(bls12377_g2, vt_bls12377_g2)
}
