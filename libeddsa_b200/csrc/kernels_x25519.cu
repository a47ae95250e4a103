// X25519 variable-base ladder kernel (sm_100a): one Diffie-Hellman per thread, persistent grid-stride
// blocks, all state in registers, constant time (the scalar only feeds cswap masks).
// Replaces x25519() / DH(): /root/reference/lib/x25519.c:129-150, 215-222, 236-243.
#include "kernel_common.cuh"
using namespace edg;

namespace {

__global__ void __launch_bounds__(kThreads) k_x25519(size_t n, uint8_t *out, const uint8_t *scalar, const uint8_t *point) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u32 s[8], p[8], o[8];
        load8(s, scalar, i);
        load8(p, point, i);
        x25519_op(o, s, p);
        store8(out, i, o);
    }
}

}  // namespace

extern "C" {

int edg_launch_x25519(size_t n, uint8_t *out, const uint8_t *scalar, const uint8_t *point, int sm_count, void *stream) {
    if (n == 0) return 0;
    int g = grid_for(k_x25519, n, 0, sm_count, nullptr);
    k_x25519<<<g, kThreads, 0, (cudaStream_t)stream>>>(n, out, scalar, point);
    return (int)cudaGetLastError();
}

}  // extern "C"
