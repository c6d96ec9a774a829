// tools/ubench/kara_proto.cu -- prototype: one level of Karatsuba inside the 256x256-bit multiplication (3 x 16 = 48 IMAD.WIDE instead
// of 64, paid for with ~55 add-with-carry instructions on the ALU pipe), against the production schoolbook bn_mul_wide, both followed by
// the production reduction.  Standalone (run under gpurun); prints throughput of r = r*y loops and values for a host-side check.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include "../../secp256k1-zkp_b200/csrc/fe.cuh"

using namespace b200;

// out[0..7] = a[0..3] * b[0..3]   (16 IMAD.WIDE, even/odd column chains like bn_mul_wide)
__device__ __forceinline__ void mul4x4(uint32_t* out, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__       // the PTX helpers of bn.cuh only exist in the device pass
    uint32_t e[9], o[9];
#pragma unroll
    for (int i = 0; i < 9; i++) { e[i] = 0; o[i] = 0; }
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
        bn_cmad2(e + i, a[0], a[2], b[i]);
        bn_cmad2(o + i, a[1], a[3], b[i]);
        bn_cmad2(o + i, a[0], a[2], b[i + 1]);
        bn_cmad2(e + i + 2, a[1], a[3], b[i + 1]);
    }
    out[0] = e[0];
    asm("add.cc.u32  %0, %7, %14;\n\taddc.cc.u32 %1, %8, %15;\n\taddc.cc.u32 %2, %9, %16;\n\taddc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\taddc.cc.u32 %5, %12, %19;\n\taddc.u32 %6, %13, %20;\n\t"
        : "=r"(out[1]), "=r"(out[2]), "=r"(out[3]), "=r"(out[4]), "=r"(out[5]), "=r"(out[6]), "=r"(out[7])
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]),
          "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]));
#endif
}

// r[0..3] = a[0..3] + b[0..3], returns the carry
__device__ __forceinline__ uint32_t add4(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t c;
    asm("add.cc.u32 %0, %5, %9;\n\taddc.cc.u32 %1, %6, %10;\n\taddc.cc.u32 %2, %7, %11;\n\taddc.cc.u32 %3, %8, %12;\n\taddc.u32 %4, 0, 0;\n\t"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
    return c;
}

__device__ __forceinline__ void mul_wide_kara(uint32_t* t, const uint32_t* a, const uint32_t* b) {
    uint32_t z0[8], z2[8], sa[4], sb[4], zm[8];
    mul4x4(z0, a, b);
    mul4x4(z2, a + 4, b + 4);
    uint32_t ca = add4(sa, a, a + 4), cb = add4(sb, b, b + 4);
    mul4x4(zm, sa, sb);
    // m = (sa + ca 2^128)(sb + cb 2^128) = zm + (ca*sb + cb*sa) 2^128 + ca*cb 2^256        (10 words: m[0..8], m[9] <= 3 unused beyond bit 257)
    uint32_t ma = 0u - ca, mb = 0u - cb;
    uint32_t m[9];
    m[0] = zm[0]; m[1] = zm[1]; m[2] = zm[2]; m[3] = zm[3];
    asm("add.cc.u32 %0, %5, %10;\n\taddc.cc.u32 %1, %6, %11;\n\taddc.cc.u32 %2, %7, %12;\n\taddc.cc.u32 %3, %8, %13;\n\taddc.u32 %4, %9, 0;\n\t"
        : "=r"(m[4]), "=r"(m[5]), "=r"(m[6]), "=r"(m[7]), "=r"(m[8])
        : "r"(zm[4]), "r"(zm[5]), "r"(zm[6]), "r"(zm[7]), "r"(ca & cb), "r"(sb[0] & ma), "r"(sb[1] & ma), "r"(sb[2] & ma), "r"(sb[3] & ma));
    asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\taddc.cc.u32 %3, %3, %8;\n\taddc.u32 %4, %4, 0;\n\t"
        : "+r"(m[4]), "+r"(m[5]), "+r"(m[6]), "+r"(m[7]), "+r"(m[8])
        : "r"(sa[0] & mb), "r"(sa[1] & mb), "r"(sa[2] & mb), "r"(sa[3] & mb));
    // m -= z0 + z2   (the cross term a_lo*b_hi + a_hi*b_lo, < 2^257)
    asm("sub.cc.u32 %0, %0, %9;\n\tsubc.cc.u32 %1, %1, %10;\n\tsubc.cc.u32 %2, %2, %11;\n\tsubc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\tsubc.cc.u32 %5, %5, %14;\n\tsubc.cc.u32 %6, %6, %15;\n\tsubc.cc.u32 %7, %7, %16;\n\tsubc.u32 %8, %8, 0;\n\t"
        : "+r"(m[0]), "+r"(m[1]), "+r"(m[2]), "+r"(m[3]), "+r"(m[4]), "+r"(m[5]), "+r"(m[6]), "+r"(m[7]), "+r"(m[8])
        : "r"(z0[0]), "r"(z0[1]), "r"(z0[2]), "r"(z0[3]), "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]));
    asm("sub.cc.u32 %0, %0, %9;\n\tsubc.cc.u32 %1, %1, %10;\n\tsubc.cc.u32 %2, %2, %11;\n\tsubc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\tsubc.cc.u32 %5, %5, %14;\n\tsubc.cc.u32 %6, %6, %15;\n\tsubc.cc.u32 %7, %7, %16;\n\tsubc.u32 %8, %8, 0;\n\t"
        : "+r"(m[0]), "+r"(m[1]), "+r"(m[2]), "+r"(m[3]), "+r"(m[4]), "+r"(m[5]), "+r"(m[6]), "+r"(m[7]), "+r"(m[8])
        : "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]));
    // t = z0 + m 2^128 + z2 2^256
    t[0] = z0[0]; t[1] = z0[1]; t[2] = z0[2]; t[3] = z0[3];
    asm("add.cc.u32 %0, %12, %16;\n\taddc.cc.u32 %1, %13, %17;\n\taddc.cc.u32 %2, %14, %18;\n\taddc.cc.u32 %3, %15, %19;\n\t"
        "addc.cc.u32 %4, %24, %20;\n\taddc.cc.u32 %5, %25, %21;\n\taddc.cc.u32 %6, %26, %22;\n\taddc.cc.u32 %7, %27, %23;\n\t"
        "addc.cc.u32 %8, %28, %32;\n\taddc.cc.u32 %9, %29, 0;\n\taddc.cc.u32 %10, %30, 0;\n\taddc.u32 %11, %31, 0;\n\t"
        : "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
        : "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]), "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]), "r"(m[4]), "r"(m[5]), "r"(m[6]), "r"(m[7]),
          "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]), "r"(m[8]));
}

template <int MODE>
__global__ void __launch_bounds__(256) k_proto(uint32_t* out, const uint32_t* in, size_t n, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x, y, r;
#pragma unroll
    for (int k = 0; k < 8; k++) { x.v[k] = in[(2 * i) * 8 + k]; y.v[k] = in[(2 * i + 1) * 8 + k]; }
    r = x;
#pragma unroll 1
    for (int k = 0; k < iters; k++) {
        uint32_t t[16];
        if (MODE == 0) bn_mul_wide(t, r.v, y.v); else mul_wide_kara(t, r.v, y.v);
        fe_reduce512(r, t);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) out[i * 8 + k] = r.v[k];
}

int main(int argc, char** argv) {
    size_t n = 148 * 2048;
    int iters = argc > 1 ? atoi(argv[1]) : 1000;
    uint32_t* h = (uint32_t*)malloc(n * 16 * 4);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < n * 16; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint32_t)s; }
    for (int k = 0; k < 16; k++) h[k] = 0xFFFFFFFFu;                 // item 0: all-ones operands (every carry path)
    uint32_t *din, *dout;
    cudaMalloc(&din, n * 16 * 4); cudaMalloc(&dout, n * 8 * 4);
    cudaMemcpy(din, h, n * 16 * 4, cudaMemcpyHostToDevice);
    uint32_t* ho = (uint32_t*)malloc(n * 8 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; mode++) {
        for (int chk = 0; chk < 2; chk++) {
            int it = chk ? 1 : iters;
            float best = 1e30f;
            for (int rep = 0; rep < (chk ? 1 : 4); rep++) {
                cudaEventRecord(e0);
                dim3 g((unsigned)((n + 255) / 256));
                if (mode == 0) k_proto<0><<<g, 256>>>(dout, din, n, it); else k_proto<1><<<g, 256>>>(dout, din, n, it);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            if (!chk) printf("{\"proto\": \"%s\", \"ms\": %.4f, \"G_per_s\": %.2f, \"err\": \"%s\"}\n", mode ? "fe_mul_karatsuba" : "fe_mul_schoolbook", best,
                             (double)n * iters / best / 1e6, cudaGetErrorString(cudaGetLastError()));
            else {
                cudaMemcpy(ho, dout, n * 8 * 4, cudaMemcpyDeviceToHost);
                for (int i = 0; i < 4; i++) {
                    printf("CHK %d", mode);
                    for (int k = 0; k < 16; k++) printf(" %u", h[i * 16 + k]);
                    for (int k = 0; k < 8; k++) printf(" %u", ho[i * 8 + k]);
                    printf("\n");
                }
            }
        }
    }
    return 0;
}
