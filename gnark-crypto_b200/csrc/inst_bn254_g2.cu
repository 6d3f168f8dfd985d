// explicit instantiation of the engine for bn254_g2
#include "engine_impl.cuh"
namespace gmsm {
GMSM_INSTANTIATE(bn254_g2, vt_bn254_g2)
}
