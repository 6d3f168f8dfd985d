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
