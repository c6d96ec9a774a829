// k_halfagg.cu -- half-aggregation verification kernels (halfagg.cuh): randomizer chain (one thread, serial by construction),
// per-signature terms (one thread per signature), and the final accept bit.  The MSM in between is k_msm.cu's.
#include "halfagg.cuh"
#include "launch.h"

namespace b200 {

static __device__ __forceinline__ void load_bytes_global(uint8_t* dst, const uint8_t* src, int nbytes) {
    if (((reinterpret_cast<uintptr_t>(src) | (uintptr_t)nbytes) & 15) == 0) {
        const uint4* q = reinterpret_cast<const uint4*>(src);
        uint4* d = reinterpret_cast<uint4*>(dst);
        for (int i = 0; i < nbytes / 16; i++) d[i] = __ldg(q + i);
    } else {
        for (int i = 0; i < nbytes; i++) dst[i] = __ldg(src + i);
    }
}

__global__ void k_halfagg_z(uint8_t* z32, halfagg_midstate mid, const uint8_t* aggsig, const uint8_t* pk32, const uint8_t* msgs32, size_t n) {
    if (blockIdx.x == 0 && threadIdx.x == 0) halfagg_z_chain(z32, mid, aggsig, pk32, msgs32, n);
}

// scalars: 2n x 32, points: 2n x 64 (term 2i = z_i*R_i, term 2i+1 = z_i*e_i*P_i); g_scalar32 = -s; fail != 0 <=> reject
__global__ void __launch_bounds__(128) k_halfagg_terms(uint8_t* __restrict__ scalars, uint8_t* __restrict__ points, uint8_t* __restrict__ g_scalar32,
        int* __restrict__ fail, const uint8_t* __restrict__ aggsig, const uint8_t* __restrict__ pk32, const uint8_t* __restrict__ msgs32,
        const uint8_t* __restrict__ z32, size_t n, bip340_midstate mid) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        uint8_t ns[32], sb[32];
        for (int k = 0; k < 32; k++) sb[k] = aggsig[32 * n + k];
        if (!halfagg_neg_s(ns, sb)) atomicOr(fail, 2);
        for (int k = 0; k < 32; k++) g_scalar32[k] = ns[k];
    }
    if (i >= n) return;
    __align__(16) uint8_t in[128], scb[64], ptb[128];
    load_bytes_global(in, aggsig + 32 * i, 32);
    load_bytes_global(in + 32, pk32 + 32 * i, 32);
    load_bytes_global(in + 64, msgs32 + 32 * i, 32);
    load_bytes_global(in + 96, z32 + 32 * i, 32);
    if (!halfagg_term(scb, ptb, in, in + 32, in + 64, in + 96, mid)) atomicOr(fail, 1);
    uint4* qs = reinterpret_cast<uint4*>(scalars + 64 * i);
    uint4* qp = reinterpret_cast<uint4*>(points + 128 * i);
    const uint4* s4 = reinterpret_cast<const uint4*>(scb);
    const uint4* p4 = reinterpret_cast<const uint4*>(ptb);
#pragma unroll
    for (int k = 0; k < 4; k++) qs[k] = s4[k];
#pragma unroll
    for (int k = 0; k < 8; k++) qp[k] = p4[k];
}

__global__ void k_halfagg_finish(uint8_t* ok, const uint8_t* out64, const int* fail) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint8_t x = 0;
        for (int k = 0; k < 64; k++) x |= out64[k];
        *ok = (uint8_t)(x == 0 && *fail == 0);
    }
}

size_t halfagg_workspace_bytes(size_t n) {
    // z | scalars | points | g_scalar | out64 | fail | MSM workspace for 2n terms   (every piece 256-byte aligned)
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    return al(32 * n + 32) + al(64 * n + 64) + al(128 * n + 128) + 256 + 256 + 256 + msm_workspace_bytes(2 * n);
}

// d_z32_in: optional precomputed randomizer chain (n x 32 bytes, device); nullptr = compute it on the device (one thread)
cudaError_t launch_halfagg_verify(uint8_t* d_ok, const uint8_t* d_pk32, const uint8_t* d_msgs32, size_t n, const uint8_t* d_aggsig,
                                  const uint8_t* d_z32_in, const uint32_t* hmid, const bip340_midstate_h& mid, uint8_t* ws, size_t ws_bytes,
                                  cudaStream_t s, uint64_t* launches) {
    if (ws_bytes < halfagg_workspace_bytes(n) || (reinterpret_cast<uintptr_t>(ws) & 15)) return cudaErrorInvalidValue;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    uint8_t* z = ws;                       ws += al(32 * n + 32);
    uint8_t* scalars = ws;                 ws += al(64 * n + 64);
    uint8_t* points = ws;                  ws += al(128 * n + 128);
    uint8_t* gsc = ws;                     ws += 256;
    uint8_t* out64 = ws;                   ws += 256;
    int* fail = reinterpret_cast<int*>(ws); ws += 256;
    cudaError_t err = cudaMemsetAsync(fail, 0, 4, s);
    if (err != cudaSuccess) return err;
    if (n && !d_z32_in) {
        halfagg_midstate hm;
        for (int i = 0; i < 8; i++) hm.s[i] = hmid[i];
        k_halfagg_z<<<1, 32, 0, s>>>(z, hm, d_aggsig, d_pk32, d_msgs32, n);
        if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    }
    bip340_midstate m;
    for (int i = 0; i < 8; i++) m.s[i] = mid.s[i];
    k_halfagg_terms<<<(unsigned)((n + 128) / 128), 128, 0, s>>>(scalars, points, gsc, fail, d_aggsig, d_pk32, d_msgs32, d_z32_in ? d_z32_in : z, n, m);
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    err = launch_msm(out64, 0, gsc, scalars, points, 2 * n, ws, msm_workspace_bytes(2 * n), 0, s, launches);
    if (err != cudaSuccess) return err;
    k_halfagg_finish<<<1, 32, 0, s>>>(d_ok, out64, fail);
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

}  // namespace b200
