// k_fe.cu -- batched F_p micro-kernel (fe_mul / fe_sqr / ...), the roofline probe of SURVEY 2.4.
#include "k_fe_impl.cuh"
namespace b200 {

// IMAD.WIDE.U32 saturation probe: 16 independent 32x32+64 multiply-accumulate chains per thread, all in
// registers.  Its MAC/s is the measured integer-pipe peak the field kernels are compared against.
__global__ void __launch_bounds__(256) k_imad_probe(uint64_t* out, int iters, uint32_t seed) {
    uint32_t a = (threadIdx.x + 1u) * 2654435761u + seed, b = (blockIdx.x + 7u) * 40503u + seed;
    uint64_t acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = ((uint64_t)a << 20) + k;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a), "r"(b));
    }
    uint64_t x = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) x ^= acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x;
}
cudaError_t launch_imad_probe(uint64_t* out, int blocks, int threads, int iters, cudaStream_t s, uint64_t* launches) {
    k_imad_probe<<<blocks, threads, 0, s>>>(out, iters, 12345u);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_fe_op(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, int iters, cudaStream_t s, uint64_t* launches) {
    return launch_fe_op_impl(op, out, a, b, n, iters, s, launches);
}
}
