// explicit instantiation of the engine for bn254_g1
#include "engine_impl.cuh"
namespace gmsm {
GMSM_INSTANTIATE(bn254_g1, vt_bn254_g1)
}
