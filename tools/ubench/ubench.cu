// tools/ubench/ubench.cu -- integer-pipe micro-benchmarks for sm_100a (standalone; run under gpurun).
// Reports, per instruction pattern, SM cycles per warp-instruction per SM sub-partition at saturation.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t a = (threadIdx.x + 1u) * 2654435761u + seed, b = (blockIdx.x + 7u) * 40503u + seed;
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = a + i * 977u;
    uint32_t av[8];
#pragma unroll
    for (int i = 0; i < 8; i++) av[i] = a * (2u * i + 3u) + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {        // 8 independent mad.wide.u32 (64-bit accumulate, no carry flag)
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++) {
                uint64_t acc = ((uint64_t)r[2 * k2 + 1] << 32) | r[2 * k2];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
                r[2 * k2] = (uint32_t)acc; r[2 * k2 + 1] = (uint32_t)(acc >> 32);
            }
        } else if (MODE == 1) { // two carry chains of 4 IMAD.WIDE.X (the fe_mul row pattern), 8 products
            asm volatile(
                "mad.lo.cc.u32  %0, %16, %24, %0;\n\tmadc.hi.cc.u32 %1, %16, %24, %1;\n\t"
                "madc.lo.cc.u32 %2, %17, %24, %2;\n\tmadc.hi.cc.u32 %3, %17, %24, %3;\n\t"
                "madc.lo.cc.u32 %4, %18, %24, %4;\n\tmadc.hi.cc.u32 %5, %18, %24, %5;\n\t"
                "madc.lo.cc.u32 %6, %19, %24, %6;\n\tmadc.hi.u32    %7, %19, %24, %7;\n\t"
                "mad.lo.cc.u32  %8, %20, %24, %8;\n\tmadc.hi.cc.u32 %9, %20, %24, %9;\n\t"
                "madc.lo.cc.u32 %10, %21, %24, %10;\n\tmadc.hi.cc.u32 %11, %21, %24, %11;\n\t"
                "madc.lo.cc.u32 %12, %22, %24, %12;\n\tmadc.hi.cc.u32 %13, %22, %24, %13;\n\t"
                "madc.lo.cc.u32 %14, %23, %24, %14;\n\tmadc.hi.u32    %15, %23, %24, %15;\n\t"
                : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                  "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                : "r"(av[0]), "r"(av[1]), "r"(av[2]), "r"(av[3]), "r"(av[4]), "r"(av[5]), "r"(av[6]), "r"(av[7]), "r"(b));
        } else if (MODE == 2) { // 16 independent 32-bit mad.lo (IMAD)
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[k2]) : "r"(a), "r"(b));
        } else if (MODE == 3) { // 16 mad.hi
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(r[k2]) : "r"(a), "r"(b));
        } else if (MODE == 4) { // two add carry chains of 8 (IADD3.X)
            asm volatile(
                "add.cc.u32 %0, %0, %16;\n\taddc.cc.u32 %1, %1, %17;\n\taddc.cc.u32 %2, %2, %16;\n\taddc.cc.u32 %3, %3, %17;\n\t"
                "addc.cc.u32 %4, %4, %16;\n\taddc.cc.u32 %5, %5, %17;\n\taddc.cc.u32 %6, %6, %16;\n\taddc.u32 %7, %7, %17;\n\t"
                "add.cc.u32 %8, %8, %16;\n\taddc.cc.u32 %9, %9, %17;\n\taddc.cc.u32 %10, %10, %16;\n\taddc.cc.u32 %11, %11, %17;\n\t"
                "addc.cc.u32 %12, %12, %16;\n\taddc.cc.u32 %13, %13, %17;\n\taddc.cc.u32 %14, %14, %16;\n\taddc.u32 %15, %15, %17;\n\t"
                : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                  "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                : "r"(a), "r"(b));
        } else if (MODE == 5) { // 8 mad.wide + 8 independent adds (dual-pipe overlap?)
#pragma unroll
            for (int k2 = 0; k2 < 4; k2++) {
                uint64_t acc = ((uint64_t)r[2 * k2 + 1] << 32) | r[2 * k2];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
                r[2 * k2] = (uint32_t)acc; r[2 * k2 + 1] = (uint32_t)(acc >> 32);
            }
#pragma unroll
            for (int k2 = 8; k2 < 16; k2++) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[k2]) : "r"(b));
        } else if (MODE == 6) { // 16 independent 3-input adds
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[k2]) : "r"(b));
        } else if (MODE == 7) { // mul.wide (no accumulate): 8
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++) {
                uint64_t acc;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(r[2 * k2]), "r"(b));
                r[2 * k2] = (uint32_t)acc ^ (uint32_t)(acc >> 32);
            }
        }
    }
    long long t1 = clock64();
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) x ^= r[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_iter, int blocks_per_sm) {
    int nsm = 148, threads = 256;
    int blocks = nsm * blocks_per_sm;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaMalloc(&cyc, blocks * sizeof(long long));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, cyc, 1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, cyc, 2);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = (long long*)malloc(blocks * sizeof(long long));
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    // warps per SMSP = blocks_per_sm * 8 warps / 4
    double warps_per_smsp = blocks_per_sm * (threads / 32) / 4.0;
    double cyc_per_warpinst = avg / ((double)ITERS * ops_per_iter * warps_per_smsp);
    double total_ops = (double)blocks * threads * ITERS * ops_per_iter;
    printf("{\"ubench\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"tops_per_s\": %.3f, \"cycles_per_warp_inst_per_smsp\": %.3f, \"block_cycles\": %.0f}\n",
           name, blocks_per_sm, ms, total_ops / ms / 1e9, cyc_per_warpinst, avg);
    cudaFree(out); cudaFree(cyc); free(h);
}

int main() {
    for (int bps = 2; bps <= 8; bps *= 2) {
        run<0>("mad.wide.u32 x8 indep (IMAD.WIDE)", 8, bps);
        run<1>("mad.lo.cc/madc.hi.cc 2 chains x4 (IMAD.WIDE.X)", 8, bps);
        run<2>("mad.lo.u32 x16 (IMAD)", 16, bps);
        run<3>("mad.hi.u32 x16 (IMAD.HI)", 16, bps);
        run<4>("add.cc/addc 2 chains x8 (IADD3.X)", 16, bps);
        run<5>("4 mad.wide + 8 add", 12, bps);
        run<6>("add.u32 x16 (IADD3)", 16, bps);
        run<7>("mul.wide.u32 x8", 8, bps);
    }
    return 0;
}
