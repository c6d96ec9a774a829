// k_codec.cu -- batched point parsing / serialization kernel (one thread per item).
#include "codec.cuh"
#include "launch.h"

namespace b200 {

__global__ void __launch_bounds__(128) k_codec(uint8_t* __restrict__ ok, uint8_t* __restrict__ out64, const uint8_t* __restrict__ in,
                                               size_t in_stride, uint32_t in_len, int fmt, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t buf[65], o[64];
    for (uint32_t k = 0; k < in_len && k < 65; k++) buf[k] = in[in_stride * i + k];
    int r = codec_item(o, buf, in_len, fmt);
    for (int k = 0; k < 64; k++) out64[64 * i + k] = o[k];
    ok[i] = (uint8_t)r;
}

cudaError_t launch_codec(uint8_t* ok, uint8_t* out64, const uint8_t* in, size_t in_stride, uint32_t in_len, int fmt, size_t n,
                         cudaStream_t s, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    k_codec<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(ok, out64, in, in_stride, in_len, fmt, n);
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

}  // namespace b200
