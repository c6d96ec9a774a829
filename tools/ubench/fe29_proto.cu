// tools/ubench/fe29_proto.cu -- prototype of a carry-free 9x29-bit field multiplication for sm_100a
// (standalone; run under gpurun).  Compares throughput of r = r*y / r = r^2 loops against the
// production 8x32 carry-chain code (102 / 144 G/s) and prints results for a host-side correctness check.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

struct fe29 { uint32_t n[9]; };
#define M29 0x1FFFFFFFu
#define R0 31264u   // 2^261 = 2^37 + 31264 (mod p) = R1*2^29 + R0
#define R1 256u

__device__ __forceinline__ void fe29_mul_inl(fe29& r, const fe29& a, const fe29& b) {
    const uint32_t* x = a.n; const uint32_t* y = b.n;
    uint64_t c, d;
    uint32_t t[9], u;
    // column 8 first: its carry joins the high half
    d = (uint64_t)x[0] * y[8];
#pragma unroll
    for (int i = 1; i <= 8; i++) d += (uint64_t)x[i] * y[8 - i];
    t[8] = (uint32_t)d & M29; d >>= 29;
    c = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
        for (int i = 0; i <= k; i++) c += (uint64_t)x[i] * y[k - i];            // column k
#pragma unroll
        for (int i = k + 1; i <= 8; i++) d += (uint64_t)x[i] * y[9 + k - i];    // column 9+k
        u = (uint32_t)d & M29; d >>= 29;
        c += (uint64_t)u * R0;
        t[k] = (uint32_t)c & M29; c >>= 29;
        c += (uint64_t)u << 8;      // u * R1
    }
    // c: carry into column 8; d: carry into column 17 = column 8 * 2^261
    uint32_t dd = (uint32_t)d;
    c += t[8];
    c += (uint64_t)dd * R0;
    t[8] = (uint32_t)c & M29; c >>= 29;
    c += (uint64_t)dd << 8;         // column 9 -> folds to columns 0,1,2
    uint32_t cl = (uint32_t)c & M29, ch = (uint32_t)(c >> 29);
    uint64_t e = (uint64_t)cl * R0 + t[0];
    r.n[0] = (uint32_t)e & M29; e >>= 29;
    e += ((uint64_t)cl << 8) + (uint64_t)ch * R0 + t[1];
    r.n[1] = (uint32_t)e & M29; e >>= 29;
    r.n[2] = t[2] + (uint32_t)e + (ch << 8);
#pragma unroll
    for (int k = 3; k < 9; k++) r.n[k] = t[k];
}

__device__ __forceinline__ void fe29_sqr_inl(fe29& r, const fe29& a) {
    const uint32_t* x = a.n;
    uint32_t x2[9];
#pragma unroll
    for (int i = 0; i < 9; i++) x2[i] = x[i] << 1;
    uint64_t c, d;
    uint32_t t[9], u;
    d = (uint64_t)x[4] * x[4];
#pragma unroll
    for (int i = 0; i < 4; i++) d += (uint64_t)x2[i] * x[8 - i];
    t[8] = (uint32_t)d & M29; d >>= 29;
    c = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
        for (int i = 0; 2 * i < k; i++) c += (uint64_t)x2[i] * x[k - i];
        if ((k & 1) == 0) c += (uint64_t)x[k / 2] * x[k / 2];
#pragma unroll
        for (int i = k + 1; 2 * i < 9 + k; i++) d += (uint64_t)x2[i] * x[9 + k - i];
        if (((9 + k) & 1) == 0) d += (uint64_t)x[(9 + k) / 2] * x[(9 + k) / 2];
        u = (uint32_t)d & M29; d >>= 29;
        c += (uint64_t)u * R0;
        t[k] = (uint32_t)c & M29; c >>= 29;
        c += (uint64_t)u << 8;
    }
    uint32_t dd = (uint32_t)d;
    c += t[8];
    c += (uint64_t)dd * R0;
    t[8] = (uint32_t)c & M29; c >>= 29;
    c += (uint64_t)dd << 8;
    uint32_t cl = (uint32_t)c & M29, ch = (uint32_t)(c >> 29);
    uint64_t e = (uint64_t)cl * R0 + t[0];
    r.n[0] = (uint32_t)e & M29; e >>= 29;
    e += ((uint64_t)cl << 8) + (uint64_t)ch * R0 + t[1];
    r.n[1] = (uint32_t)e & M29; e >>= 29;
    r.n[2] = t[2] + (uint32_t)e + (ch << 8);
#pragma unroll
    for (int k = 3; k < 9; k++) r.n[k] = t[k];
}

static __device__ __noinline__ fe29 fe29_mul_nv(fe29 a, fe29 b) { fe29 r; fe29_mul_inl(r, a, b); return r; }
static __device__ __noinline__ fe29 fe29_sqr_nv(fe29 a) { fe29 r; fe29_sqr_inl(r, a); return r; }

template <int OP, int NV>
__global__ void __launch_bounds__(256) k_proto(uint32_t* out, const uint32_t* in, size_t n, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe29 x, y, r;
#pragma unroll
    for (int k = 0; k < 9; k++) { x.n[k] = in[(2 * i) * 9 + k]; y.n[k] = in[(2 * i + 1) * 9 + k]; }
    r = x;
#pragma unroll 1
    for (int k = 0; k < iters; k++) {
        if (OP == 0) { if (NV) r = fe29_mul_nv(r, y); else fe29_mul_inl(r, r, y); }
        else { if (NV) r = fe29_sqr_nv(r); else fe29_sqr_inl(r, r); }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) out[i * 9 + k] = r.n[k];
}

int main(int argc, char** argv) {
    size_t n = 148 * 2048;
    int iters = argc > 1 ? atoi(argv[1]) : 1000;
    uint32_t* h = (uint32_t*)malloc(n * 18 * 4);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < n * 18; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint32_t)s & M29; }
    uint32_t *din, *dout;
    cudaMalloc(&din, n * 18 * 4); cudaMalloc(&dout, n * 9 * 4);
    cudaMemcpy(din, h, n * 18 * 4, cudaMemcpyHostToDevice);
    uint32_t* ho = (uint32_t*)malloc(n * 9 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 4; mode++) {
        int op = mode & 1, nv = mode >> 1;
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            dim3 g((unsigned)((n + 255) / 256));
            if (mode == 0) k_proto<0, 0><<<g, 256>>>(dout, din, n, iters);
            if (mode == 1) k_proto<1, 0><<<g, 256>>>(dout, din, n, iters);
            if (mode == 2) k_proto<0, 1><<<g, 256>>>(dout, din, n, iters);
            if (mode == 3) k_proto<1, 1><<<g, 256>>>(dout, din, n, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        cudaError_t err = cudaGetLastError();
        printf("{\"proto\": \"fe29_%s%s\", \"ms\": %.4f, \"G_per_s\": %.2f, \"err\": \"%s\"}\n", op ? "sqr" : "mul", nv ? "_nv" : "_inl", best,
               (double)n * iters / best / 1e6, cudaGetErrorString(err));
        cudaMemcpy(ho, dout, n * 9 * 4, cudaMemcpyDeviceToHost);
        for (int i = 0; i < 3; i++) {
            printf("CHK %d %d", op, iters);
            for (int k = 0; k < 9; k++) printf(" %u", h[(2 * i) * 9 + k]);
            for (int k = 0; k < 9; k++) printf(" %u", h[(2 * i + 1) * 9 + k]);
            for (int k = 0; k < 9; k++) printf(" %u", ho[i * 9 + k]);
            printf("\n");
        }
    }
    return 0;
}
