// k_msm.cu -- Pippenger multi-scalar multiplication kernels (see msm.cuh for the algorithm).
#include "msm.cuh"
#include "gtable.cuh"
#include "launch.h"

namespace b200 {

struct msm_ws {
    uint32_t* scalars;     // [N][8]   reduced scalars, LE limbs
    uint32_t* points;      // [N][16]  affine LE storage, zero = infinity
    uint32_t* hist;        // [16][32768]
    uint32_t* offsets;     // [16][32768]
    uint32_t* cursor;      // [16][32768]
    uint32_t* entries;     // [16][N]
    gej_store* buckets;    // [16][32768]
    gej_store* chunks;     // [16][1024]
    gej_store* groups;     // [16][32]
    gej_store* windows;    // [16]
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
size_t msm_workspace_bytes(size_t n_terms) {
    size_t N = n_terms + 1;   // + optional generator term
    size_t b = 0;
    b += align256(N * 32) + align256(N * 64);
    b += 3 * align256((size_t)MSM_WINDOWS * MSM_BUCKETS * 4);
    b += align256((size_t)MSM_WINDOWS * N * 4);
    b += align256((size_t)MSM_WINDOWS * MSM_BUCKETS * sizeof(gej_store));
    b += align256((size_t)MSM_WINDOWS * MSM_CHUNKS * sizeof(gej_store));
    b += align256((size_t)MSM_WINDOWS * 32 * sizeof(gej_store));
    b += align256((size_t)MSM_WINDOWS * sizeof(gej_store));
    return b;
}
static msm_ws carve_ws(uint8_t* base, size_t N) {
    msm_ws w;
    size_t o = 0;
    w.scalars = (uint32_t*)(base + o); o += align256(N * 32);
    w.points = (uint32_t*)(base + o); o += align256(N * 64);
    w.hist = (uint32_t*)(base + o); o += align256((size_t)MSM_WINDOWS * MSM_BUCKETS * 4);
    w.offsets = (uint32_t*)(base + o); o += align256((size_t)MSM_WINDOWS * MSM_BUCKETS * 4);
    w.cursor = (uint32_t*)(base + o); o += align256((size_t)MSM_WINDOWS * MSM_BUCKETS * 4);
    w.entries = (uint32_t*)(base + o); o += align256((size_t)MSM_WINDOWS * N * 4);
    w.buckets = (gej_store*)(base + o); o += align256((size_t)MSM_WINDOWS * MSM_BUCKETS * sizeof(gej_store));
    w.chunks = (gej_store*)(base + o); o += align256((size_t)MSM_WINDOWS * MSM_CHUNKS * sizeof(gej_store));
    w.groups = (gej_store*)(base + o); o += align256((size_t)MSM_WINDOWS * 32 * sizeof(gej_store));
    w.windows = (gej_store*)(base + o);
    return w;
}

// (1) parse + histogram.  Term N-1 is the optional generator term (g_scalar * G).
__global__ void __launch_bounds__(256) k_msm_prepare(msm_ws ws, const uint8_t* __restrict__ scalars32, const uint8_t* __restrict__ points64,
                                                     const uint8_t* __restrict__ g_scalar32, size_t n, size_t N) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    sc k;
    uint32_t p16[16];
    if (i < n) {
        uint8_t b[32];
        for (int t = 0; t < 32; t++) b[t] = scalars32[32 * i + t];
        sc_set_b32(k, b);
        fe x, y;
        uint8_t pb[64];
        for (int t = 0; t < 64; t++) pb[t] = points64[64 * i + t];
        bn_from_be32(x.v, pb); bn_from_be32(y.v, pb + 32);
        for (int t = 0; t < 8; t++) { p16[t] = x.v[t]; p16[8 + t] = y.v[t]; }
    } else {
        if (g_scalar32) {
            uint8_t b[32];
            for (int t = 0; t < 32; t++) b[t] = g_scalar32[t];
            sc_set_b32(k, b);
        } else sc_set_zero(k);
        ge g; ge_set_generator(g);
        for (int t = 0; t < 8; t++) { p16[t] = g.x.v[t]; p16[8 + t] = g.y.v[t]; }
    }
    for (int t = 0; t < 8; t++) ws.scalars[8 * i + t] = k.v[t];
    uint32_t z = 0;
    for (int t = 0; t < 16; t++) { ws.points[16 * i + t] = p16[t]; z |= p16[t]; }
    if (z == 0) return;                      // infinity contributes nothing
    int32_t d[MSM_WINDOWS];
    msm_digits(d, k);
#pragma unroll
    for (int w = 0; w < MSM_WINDOWS; w++) {
        if (d[w] != 0) {
            uint32_t b = (uint32_t)(d[w] < 0 ? -d[w] : d[w]) - 1u;
            atomicAdd(&ws.hist[(size_t)w * MSM_BUCKETS + b], 1u);
        }
    }
}

// (2) exclusive scan of the 32768 bucket counts of one window per block (1024 threads x 32 buckets)
__global__ void __launch_bounds__(1024) k_msm_scan(msm_ws ws) {
    __shared__ uint32_t part[1024];
    const uint32_t w = blockIdx.x, t = threadIdx.x;
    const uint32_t* h = ws.hist + (size_t)w * MSM_BUCKETS + t * 32;
    uint32_t local[32], sum = 0;
    for (int k = 0; k < 32; k++) { local[k] = sum; sum += h[k]; }
    part[t] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partial sums
    for (uint32_t off = 1; off < 1024; off <<= 1) {
        uint32_t v = (t >= off) ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t base = part[t] - sum;
    uint32_t* o = ws.offsets + (size_t)w * MSM_BUCKETS + t * 32;
    uint32_t* c = ws.cursor + (size_t)w * MSM_BUCKETS + t * 32;
    for (int k = 0; k < 32; k++) { o[k] = base + local[k]; c[k] = 0; }
}

// (3) counting-sort scatter
__global__ void __launch_bounds__(256) k_msm_scatter(msm_ws ws, size_t N) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t z = 0;
    for (int t = 0; t < 16; t++) z |= ws.points[16 * i + t];
    if (z == 0) return;
    sc k;
    for (int t = 0; t < 8; t++) k.v[t] = ws.scalars[8 * i + t];
    int32_t d[MSM_WINDOWS];
    msm_digits(d, k);
#pragma unroll
    for (int w = 0; w < MSM_WINDOWS; w++) {
        if (d[w] != 0) {
            uint32_t neg = d[w] < 0;
            uint32_t b = (uint32_t)(neg ? -d[w] : d[w]) - 1u;
            size_t wb = (size_t)w * MSM_BUCKETS + b;
            uint32_t pos = ws.offsets[wb] + atomicAdd(&ws.cursor[wb], 1u);
            ws.entries[(size_t)w * N + pos] = (uint32_t)i | (neg << 31);
        }
    }
}

// (4) one thread per bucket
__global__ void __launch_bounds__(128) k_msm_bucket_sum(msm_ws ws, size_t N) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)MSM_WINDOWS * MSM_BUCKETS) return;
    uint32_t w = (uint32_t)(t / MSM_BUCKETS);
    uint32_t begin = ws.offsets[t], end = begin + ws.hist[t];
    gej acc;
    msm_bucket_sum(acc, ws.entries + (size_t)w * N, begin, end, ws.points);
    msm_store_gej(ws.buckets[t], acc);
}

// (5) chunk reduction
__global__ void __launch_bounds__(128) k_msm_chunks(msm_ws ws) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)MSM_WINDOWS * MSM_CHUNKS) return;
    uint32_t w = (uint32_t)(t / MSM_CHUNKS), ch = (uint32_t)(t % MSM_CHUNKS);
    gej out;
    msm_chunk_reduce(out, ws.buckets + (size_t)w * MSM_BUCKETS, ch * MSM_CHUNK);
    msm_store_gej(ws.chunks[t], out);
}
// (6a) 32 chunks -> 1 group ; (6b) 32 groups -> window sum
__global__ void __launch_bounds__(64) k_msm_groups(gej_store* dst, const gej_store* src, uint32_t ndst, uint32_t per) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ndst) return;
    gej out;
    msm_sum_range(out, src + (size_t)t * per, per);
    msm_store_gej(dst[t], out);
}
// (6c) Horner over the windows, to affine, serialize x||y big-endian (zero = infinity)
__global__ void k_msm_final(uint8_t* out64, msm_ws ws) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    gej r;
    msm_horner(r, ws.windows);
    if (r.inf) { for (int k = 0; k < 64; k++) out64[k] = 0; return; }
    ge a; ge_set_gej(a, r);
    fe_get_b32(out64, a.x); fe_get_b32(out64 + 32, a.y);
}

cudaError_t launch_msm(uint8_t* out64, const uint8_t* g_scalar32, const uint8_t* scalars32, const uint8_t* points64, size_t n,
                       uint8_t* workspace, size_t workspace_bytes, cudaStream_t s, uint64_t* launches) {
    if (workspace_bytes < msm_workspace_bytes(n)) return cudaErrorInvalidValue;
    if (n + 1 >= ((size_t)1 << 31)) return cudaErrorInvalidValue;
    const size_t N = n + 1;
    msm_ws ws = carve_ws(workspace, N);
    cudaError_t err = cudaMemsetAsync(ws.hist, 0, (size_t)MSM_WINDOWS * MSM_BUCKETS * 4, s);
    if (err != cudaSuccess) return err;
    k_msm_prepare<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(ws, scalars32, points64, g_scalar32, n, N);
    k_msm_scan<<<MSM_WINDOWS, 1024, 0, s>>>(ws);
    k_msm_scatter<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(ws, N);
    k_msm_bucket_sum<<<(MSM_WINDOWS * MSM_BUCKETS + 127) / 128, 128, 0, s>>>(ws, N);
    k_msm_chunks<<<(MSM_WINDOWS * MSM_CHUNKS + 127) / 128, 128, 0, s>>>(ws);
    k_msm_groups<<<(MSM_WINDOWS * 32 + 63) / 64, 64, 0, s>>>(ws.groups, ws.chunks, MSM_WINDOWS * 32, 32);
    k_msm_groups<<<1, 64, 0, s>>>(ws.windows, ws.groups, MSM_WINDOWS, 32);
    k_msm_final<<<1, 32, 0, s>>>(out64, ws);
    if (launches) *launches += 8;
    return cudaGetLastError();
}

}  // namespace b200
