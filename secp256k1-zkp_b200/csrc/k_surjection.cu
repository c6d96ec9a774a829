// k_surjection.cu -- surjection proof verification / Pedersen tally kernels (one thread per item).
#include "surjection.cuh"
#include "launch.h"

namespace b200 {

__global__ void __launch_bounds__(128, 5) k_surjection_verify(uint8_t* __restrict__ ok, const uint8_t* __restrict__ proofs, size_t stride,
        const uint32_t* __restrict__ plen, const uint8_t* __restrict__ in_tags64, const uint32_t* __restrict__ tag_off,
        const uint8_t* __restrict__ out_tags64, size_t n, ecmult_gtable gt) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tab_local tab;   // local memory, see k_verify.cu
    uint32_t t0 = tag_off[i], t1 = tag_off[i + 1];
    if (plen[i] > stride || t1 < t0) { ok[i] = 0; return; }          // device-side lengths are not trusted (the _dev call cannot pre-check them)
    ok[i] = (uint8_t)surjection_verify_item(proofs + stride * i, plen[i], in_tags64 + 64 * (size_t)t0, t1 - t0, out_tags64 + 64 * i, gt, tab);
}
__global__ void __launch_bounds__(128) k_pedersen_tally(uint8_t* __restrict__ ok, const uint8_t* __restrict__ pos33, const uint32_t* __restrict__ pos_off,
        const uint8_t* __restrict__ neg33, const uint32_t* __restrict__ neg_off, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ok[i] = (uint8_t)pedersen_tally_item(pos33 + 33 * (size_t)pos_off[i], pos_off[i + 1] - pos_off[i], neg33 + 33 * (size_t)neg_off[i], neg_off[i + 1] - neg_off[i]);
}

cudaError_t launch_surjection_verify(uint8_t* ok, const uint8_t* proofs, size_t stride, const uint32_t* plen, const uint8_t* in_tags64,
                                     const uint32_t* tag_off, const uint8_t* out_tags64, size_t n, const gtable_dev& gt, cudaStream_t s, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    const int smem = 0;
    ecmult_gtable g; g.data = gt.data; g.gw = gt.gw; g.windows = gt.windows;
    k_surjection_verify<<<(unsigned)((n + 127) / 128), 128, smem, s>>>(ok, proofs, stride, plen, in_tags64, tag_off, out_tags64, n, g);
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}
cudaError_t launch_pedersen_tally(uint8_t* ok, const uint8_t* pos33, const uint32_t* pos_off, const uint8_t* neg33, const uint32_t* neg_off, size_t n,
                                  cudaStream_t s, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    k_pedersen_tally<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(ok, pos33, pos_off, neg33, neg_off, n);
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

}  // namespace b200
