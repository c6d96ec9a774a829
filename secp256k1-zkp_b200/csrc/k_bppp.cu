// k_bppp.cu -- BP++ norm-product argument verification kernel (one warp per proof; see bppp.cuh).
#include "bppp.cuh"
#include "launch.h"

namespace b200 {

#define BP_WARPS 4
#define BP_TAB_WORDS (32 * 8 * 16)
#define BP_WARP_STRIDE (BP_TAB_WORDS * 4)     // tables, later reused for the 32 partial sums (32 x 100 B)

__global__ void __launch_bounds__(BP_WARPS * 32, 3) k_bppp_norm_verify(uint8_t* __restrict__ ok, const uint8_t* __restrict__ proofs, size_t stride,
        const uint32_t* __restrict__ plen, const uint8_t* __restrict__ transcripts, const uint8_t* __restrict__ rho32,
        const uint8_t* __restrict__ gens64, uint32_t g_len, const uint8_t* __restrict__ c_vec32, uint32_t h_len,
        const uint8_t* __restrict__ commits64, size_t n, ecmult_gtable gt) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t item = (size_t)blockIdx.x * BP_WARPS + warp;
    if (item >= n) return;
    uint8_t* wbase = smem_raw + (size_t)warp * BP_WARP_STRIDE;
    uint32_t* tabw = reinterpret_cast<uint32_t*>(wbase);
    gej* pts = reinterpret_cast<gej*>(wbase);
    const uint8_t* proof = proofs + item * stride;

    sha256 tr;
    {
        const uint8_t* t = transcripts + 104 * item;
        for (int k = 0; k < 8; k++) tr.s[k] = (uint32_t)t[4 * k] | ((uint32_t)t[4 * k + 1] << 8) | ((uint32_t)t[4 * k + 2] << 16) | ((uint32_t)t[4 * k + 3] << 24);
        for (int k = 0; k < 64; k++) tr.buf[k] = t[32 + k];
        uint64_t b = 0;
        for (int k = 7; k >= 0; k--) b = (b << 8) | t[96 + k];
        tr.bytes = b;
    }
    sc rho;
    int good = !sc_set_b32(rho, rho32 + 32 * item);          // the reference's callers CHECK(!overflow); reject here
    bppp_ctx c;
    good = good && plen[item] <= stride;                     // device-side lengths are not trusted
    good = good && bppp_prepare(c, proof, plen[item], tr, rho, g_len, h_len);
    if (!good) { if (lane == 0) ok[item] = 0; return; }      // uniform across the warp

    gej acc; gej_set_infinity(acc);
    tab_shared tab = {tabw + lane, 32};
    const uint32_t T = bppp_num_terms(c);
    int bad = 0;
    for (uint32_t t = lane; t < T; t += 32) {
        ge pt; sc s;
        if (!bppp_term(pt, s, c, t, proof, gens64)) { bad = 1; break; }
        if (pt.inf) continue;
        gej pj, r, sum, zero_unused;
        (void)zero_unused;
        gej_set_ge(pj, pt);
        sc zero; sc_set_zero(zero);
        ecmult(r, pj, s, zero, gt, tab);
        gej_add(sum, acc, r);
        acc = sum;
    }
    if (lane == 0) {
        // + C - v*G
        sc v, zero; sc_set_zero(zero);
        bppp_v(v, c, c_vec32 + (size_t)32 * h_len * item);
        sc_negate(v, v);
        gej inf, r, sum;
        gej_set_infinity(inf);
        ecmult(r, inf, zero, v, gt, tab);
        gej_add(sum, acc, r); acc = sum;
        ge cm;
        const uint8_t* cb = commits64 + 64 * item;
        uint8_t z = 0;
        for (int k = 0; k < 64; k++) z |= cb[k];
        if (z) {
            fe_set_b32(cm.x, cb); fe_set_b32(cm.y, cb + 32); cm.inf = 0;
            gej_add_ge(sum, acc, cm); acc = sum;
        }
    }
    bad = __any_sync(0xffffffffu, bad);
    __syncwarp();                                            // all table use finished: the region becomes the sum scratch
    pts[lane] = acc;
    __syncwarp();
    for (int d = 16; d >= 1; d >>= 1) {
        if ((int)lane < d) { gej t; gej_add(t, pts[lane], pts[lane + d]); pts[lane] = t; }
        __syncwarp();
    }
    if (lane == 0) {
        ok[item] = (uint8_t)(!bad && pts[0].inf);
    }
}

cudaError_t launch_bppp_norm_verify(uint8_t* ok, const uint8_t* proofs, size_t stride, const uint32_t* plen, const uint8_t* transcripts,
                                    const uint8_t* rho32, const uint8_t* gens64, uint32_t g_len, const uint8_t* c_vec32, uint32_t h_len,
                                    const uint8_t* commits64, size_t n, const gtable_dev& gt, cudaStream_t s, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    const int smem = BP_WARPS * BP_WARP_STRIDE;
    cudaError_t err = cudaFuncSetAttribute(k_bppp_norm_verify, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    ecmult_gtable g; g.data = gt.data; g.gw = gt.gw; g.windows = gt.windows;
    k_bppp_norm_verify<<<(unsigned)((n + BP_WARPS - 1) / BP_WARPS), BP_WARPS * 32, smem, s>>>(ok, proofs, stride, plen, transcripts, rho32, gens64, g_len,
                                                                                          c_vec32, h_len, commits64, n, g);
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

}  // namespace b200
