// explicit instantiation of the engine for bls12381_g1
#include "engine_impl.cuh"
namespace gmsm {
GMSM_INSTANTIATE(bls12381_g1, vt_bls12381_g1)
}
