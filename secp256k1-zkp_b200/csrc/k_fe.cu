// k_fe.cu -- batched F_p micro-kernel (fe_mul / fe_sqr / ...), the roofline probe of SURVEY 2.4.
#include "k_fe_impl.cuh"
namespace b200 {

// IMAD.WIDE.U32 saturation probe: 16 independent 32x32+64 multiply-accumulate chains per thread, all in
// registers.  Its MAC/s is the measured integer-pipe peak the field kernels are compared against.
__global__ void __launch_bounds__(256) k_imad_probe(uint64_t* out, int iters, uint32_t seed) {
    // 16 independent 64-bit accumulators, 16 distinct multiplicand pairs (u[k&3], v[k>>2]); the u's change every round (4
    // adds on the ALU pipe per 16 multiply-adds), so ptxas can neither hoist a product out of the loop nor share one between
    // accumulators.  (The first version of this probe multiplied two loop-invariant registers and was folded into one
    // IMAD.WIDE plus 16 adds: it over-stated the peak.)  SASS: 16 IMAD.WIDE.U32 + 4 IADD3 per round, checked with cuobjdump.
    uint32_t a = (threadIdx.x + 1u) * 2654435761u + seed, b = (blockIdx.x * 256u + threadIdx.x + 7u) * 40503u + seed;
    uint64_t acc[16];
    uint32_t u[4], v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { u[k] = a * (2u * k + 3u) + k; v[k] = b * (2u * k + 5u) + 977u * k; }
    const uint64_t* in = out + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0;   // accumulators start from memory: kept as 64-bit pairs
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = in[k] + a;
#pragma unroll 1
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
#pragma unroll
            for (int k = 0; k < 16; k++) acc[k] += (uint64_t)u[k & 3] * v[k >> 2];
#pragma unroll
            for (int k = 0; k < 4; k++) u[k] += 0x9E3779B9u + k;
        }
    }
    uint64_t x = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) x ^= acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x;
}
cudaError_t launch_imad_probe(uint64_t* out, int blocks, int threads, int iters, cudaStream_t s, uint64_t* launches) {
    k_imad_probe<<<blocks, threads, 0, s>>>(out, iters, 12345u);
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

cudaError_t launch_fe_op(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, int iters, cudaStream_t s, uint64_t* launches) {
    return launch_fe_op_impl(op, out, a, b, n, iters, s, launches);
}
}
