// k_gtable.cu -- builds the generator comb table on the device at engine init.
#include "gtable.cuh"
#include "launch.h"

namespace b200 {

__global__ void k_gtable_bases(uint32_t* bases, int gw, int windows) {
    if (blockIdx.x == 0 && threadIdx.x == 0) gtable_compute_bases(bases, gw, windows);
}
__global__ void __launch_bounds__(128) k_gtable_entries(uint32_t* tab, const uint32_t* bases, int gw, int windows) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t per = (size_t)1 << gw;
    if (i >= per * (size_t)windows) return;
    uint32_t d = (uint32_t)(i & (per - 1));
    int j = (int)(i >> gw);
    uint32_t* out = tab + i * 16;
    if (d == 0) {
#pragma unroll
        for (int k = 0; k < 16; k++) out[k] = 0;
        return;
    }
    uint32_t e[16];
    gtable_compute_entry(e, bases + 16 * j, d, gw);
    uint4* o4 = reinterpret_cast<uint4*>(out);
    o4[0] = make_uint4(e[0], e[1], e[2], e[3]);
    o4[1] = make_uint4(e[4], e[5], e[6], e[7]);
    o4[2] = make_uint4(e[8], e[9], e[10], e[11]);
    o4[3] = make_uint4(e[12], e[13], e[14], e[15]);
}

cudaError_t launch_gtable_build(uint32_t* d_tab, uint32_t* d_bases, int gw, int windows, cudaStream_t s, uint64_t* launches) {
    k_gtable_bases<<<1, 32, 0, s>>>(d_bases, gw, windows);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    size_t total = ((size_t)1 << gw) * (size_t)windows;
    k_gtable_entries<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(d_tab, d_bases, gw, windows);
    if (launches) __atomic_fetch_add(launches, (uint64_t)2, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

}  // namespace b200
