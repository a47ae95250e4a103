// TEST INFRASTRUCTURE: stands in for <cuda_runtime.h> when libeddsa_b200/csrc/kernel_common.cuh and kernels_*.cu are compiled
// for the host under tests/host_sim/ptx_emul.h / simt_emul.h (which supply the keywords, intrinsics and thread geometry); only
// what those files name is declared.  The runtime calls resolve to the simulator (cudasim.cpp).
#pragma once
#include <stddef.h>
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
typedef struct CUstream_st *cudaStream_t;
extern "C" cudaError_t cudaGetLastError(void);
extern "C" cudaError_t cudaMemsetAsync(void *devPtr, int value, size_t count, cudaStream_t stream);
// resident blocks per SM as on sm_100a for the shipped launch bounds (4 blocks of 128 threads); the comb kernel's single block
// of 512 threads is a constant of its launcher
template <typename K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *blocks, K, int, size_t) { *blocks = 4; return cudaSuccess; }
template <typename K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
