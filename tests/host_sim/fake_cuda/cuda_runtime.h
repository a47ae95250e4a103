// TEST INFRASTRUCTURE: stands in for <cuda_runtime.h> when libeddsa_b200/csrc/kernel_common.cuh is compiled for the host under
// tests/host_sim/ptx_emul.h (which supplies the keywords and intrinsics); only what the header's templates name is declared.
#pragma once
typedef int cudaError_t;
enum { cudaSuccess = 0 };
