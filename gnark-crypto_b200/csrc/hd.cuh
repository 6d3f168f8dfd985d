// Host/device portability macros.  The arithmetic headers compile both with nvcc (device: inline
// PTX carry chains) and with plain g++ (portable C++ path, used by the CPU-side formula tests).
#pragma once
#if defined(__CUDACC__)
#define GMSM_HD __host__ __device__ __forceinline__
#define GMSM_D __device__ __forceinline__
#else
#define GMSM_HD inline __attribute__((always_inline))
#define GMSM_D inline __attribute__((always_inline))
#endif

// GMSM_PTX_PATH selects the carry-chain formulation of the field arithmetic (field.cuh): on the device it is inline PTX;
// with -DGMSM_EMULATE_PTX a host build runs the SAME source over emulated primitives (one carry flag per thread), so the
// exact limb/carry schedule that ptxas sees is checked on a CPU by tests/test_hostcheck.py.
#if defined(__CUDA_ARCH__) || (defined(GMSM_EMULATE_PTX) && !defined(__CUDACC__))
#define GMSM_PTX_PATH 1
#endif
