"""gnark-crypto_b200: B200-native multi-scalar multiplication behind gnark-crypto's MultiExp.

Layout:
  csrc/        sm_100a CUDA kernels + the C ABI (include/gmsm.h) -> libgmsm.so (built in-tree)
  _native.py   ctypes loader of libgmsm.so (fails loudly when the library or a GPU is missing)
  multiexp.py  host-side mirror of the reference interface for this path:
               G1Affine/G1Jac/G2Affine/G2Jac .MultiExp(points, scalars, MultiExpConfig)
               (ecc/bn254/multiexp.go:20,32,345,357; ecc/ecc.go:107-110) + the device-level Engine
  dist.py      multi-GPU: one process per GPU, shard points/scalars, all-gather the per-window partials
"""
from . import _native  # noqa: F401
from .multiexp import (  # noqa: F401
    BatchScalarMultiplication,
    CURVES,
    Engine,
    G1Affine,
    G1Jac,
    G2Affine,
    G2Jac,
    MultiExpConfig,
    MultiExpError,
    curve_package,
)

__all__ = ["BatchScalarMultiplication", "CURVES", "Engine", "G1Affine", "G1Jac", "G2Affine", "G2Jac", "MultiExpConfig", "MultiExpError", "curve_package"]
