// k_bppp.cu -- BP++ norm-product argument verification kernel (one warp per proof; see bppp.cuh).
#include <stdlib.h>
#include "bppp.cuh"
#include "gtable.cuh"
#include "ge_w10.cuh"
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

// ---- large batches: 4 proofs per warp, comb tables for the generators ------------------------------------------------------------
// In k_bppp_norm_verify every lane runs a whole a*P for its term, 24 of the 32 for the FIXED generators G_i, H_i of the batch.  From
// BPK_MIN_PROOFS proofs on, the launcher first builds an 8-bit comb table T_i[j][d] = d 2^(8j) P_i per generator on the stream (one warp per
// generator: 31 x 8 doublings in lane form; then one thread per entry), and the verification kernel packs 4 proofs into a warp: 8 lanes per
// proof, each with one of the proof's 8 variable points (a full a*P) and three generator terms (32 table additions each, no doublings).
// r2t/r2v: 8192 distinct proofs 7.9 -> 4.15 ms; the packed path costs ~2 ms whatever the batch (0.7 ms of table build, then one warp's chain).
#define BPK_MIN_PROOFS 2048        // r2v: 2048 proofs 1.99 ms packed against 2.48 ms warp-per-proof; 1024 proofs 1.99 against 1.25
#define BPK_GW 8
#define BPK_WINDOWS 32
#define BPK_GEN_WORDS ((size_t)BPK_WINDOWS * 256 * 16)
#define BPK_WARPS 4

struct bpk_base { uint32_t w[24]; };                              // Jacobian x, y, z (a base 2^(8j) P_i; never infinity)

// one warp per generator: bases[i][j] = 2^(8 j) P_i
__global__ void __launch_bounds__(128) k_bppp_gen_bases(bpk_base* __restrict__ bases, const uint8_t* __restrict__ gens64, uint32_t count) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= count) return;                                   // whole warp
    gej a;
    fe_set_b32(a.x, gens64 + 64 * warp); fe_set_b32(a.y, gens64 + 64 * warp + 32); fe_set_int(a.z, 1); a.inf = 0;
#pragma unroll 1
    for (int j = 0; j < BPK_WINDOWS; j++) {
        if (lane < 24) {
            fe z = a.z; fe_normalize(z);
            const uint32_t v = lane < 8 ? a.x.v[lane] : (lane < 16 ? a.y.v[lane - 8] : z.v[lane - 16]);
            bases[(size_t)warp * BPK_WINDOWS + j].w[lane] = v;
        }
        gej t;
        gej_double_n_w10(t, a, BPK_GW);
        a = t;
    }
}
// one thread per table entry: d * base by double-and-add, to affine storage (entry 0 unused: zeros)
__global__ void __launch_bounds__(128) k_bppp_gen_entries(uint32_t* __restrict__ tab, const bpk_base* __restrict__ bases, uint32_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)count * BPK_WINDOWS * 256) return;
    const uint32_t d = (uint32_t)(i & 255);
    uint32_t* out = tab + i * 16;
    if (d == 0) { for (int k = 0; k < 16; k++) out[k] = 0; return; }
    const bpk_base* bs = bases + (i >> 8);
    gej b;
    for (int k = 0; k < 8; k++) { b.x.v[k] = bs->w[k]; b.y.v[k] = bs->w[8 + k]; b.z.v[k] = bs->w[16 + k]; }
    b.inf = 0;
    gej acc; gej_set_infinity(acc);
#pragma unroll 1
    for (int k = BPK_GW - 1; k >= 0; k--) {
        gej t;
        if (!acc.inf) { gej_double(t, acc); acc = t; }
        if ((d >> k) & 1u) { gej_add(t, acc, b); acc = t; }
    }
    ge r; ge_set_gej(r, acc);
    uint32_t e[16];
    ge_to_storage(e, r);
    uint4* o4 = reinterpret_cast<uint4*>(out);
    o4[0] = make_uint4(e[0], e[1], e[2], e[3]); o4[1] = make_uint4(e[4], e[5], e[6], e[7]);
    o4[2] = make_uint4(e[8], e[9], e[10], e[11]); o4[3] = make_uint4(e[12], e[13], e[14], e[15]);
}

__global__ void __launch_bounds__(BPK_WARPS * 32, 3) k_bppp_norm_verify_packed(uint8_t* __restrict__ ok, const uint8_t* __restrict__ proofs, size_t stride,
        const uint32_t* __restrict__ plen, const uint8_t* __restrict__ transcripts, const uint8_t* __restrict__ rho32,
        const uint8_t* __restrict__ gens64, uint32_t g_len, const uint8_t* __restrict__ c_vec32, uint32_t h_len,
        const uint8_t* __restrict__ commits64, size_t n, ecmult_gtable gt, const uint32_t* __restrict__ gentab) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane >> 3, q = lane & 7;
    size_t item = ((size_t)blockIdx.x * BPK_WARPS + warp) * 4 + grp;
    if (((size_t)blockIdx.x * BPK_WARPS + warp) * 4 >= n) return;          // whole warp
    const bool valid = item < n;
    if (!valid) item = n - 1;                                              // a surplus group redoes the last proof and stores nothing
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
    int good = !sc_set_b32(rho, rho32 + 32 * item);
    bppp_ctx c;
    good = good && plen[item] <= stride;
    good = good && bppp_prepare(c, proof, plen[item], tr, rho, g_len, h_len);  // the 8 lanes of a proof agree; groups may diverge

    gej acc; gej_set_infinity(acc);
    tab_shared tab = {tabw + lane, 32};
    int bad = 0;
    if (good) {
        const uint32_t T = bppp_num_terms(c), nvar = 2 * c.n_rounds;
        for (uint32_t t = q; t < T; t += 8) {
            ge pt; sc s;
            if (!bppp_term(pt, s, c, t, proof, gens64)) { bad = 1; break; }
            if (t < nvar) {                                                  // a point of the proof: full a*P
                if (pt.inf) continue;
                gej pj, r, sum;
                gej_set_ge(pj, pt);
                sc zero; sc_set_zero(zero);
                ecmult(r, pj, s, zero, gt, tab);
                gej_add(sum, acc, r);
                acc = sum;
            } else {                                                         // a generator of the batch: its comb table
                ecmult_gtable g; g.data = gentab + (size_t)(t - nvar) * BPK_GEN_WORDS; g.gw = BPK_GW; g.windows = BPK_WINDOWS;
                ecmult_gen_accumulate(acc, s, g);
            }
        }
        if (q == 0) {                                                        // + C - v*G
            sc v;
            bppp_v(v, c, c_vec32 + (size_t)32 * h_len * item);
            sc_negate(v, v);
            ecmult_gen_accumulate(acc, v, gt);
            ge cm;
            const uint8_t* cb = commits64 + 64 * item;
            uint8_t z = 0;
            for (int k = 0; k < 64; k++) z |= cb[k];
            if (z) {
                gej sum;
                fe_set_b32(cm.x, cb); fe_set_b32(cm.y, cb + 32); cm.inf = 0;
                gej_add_ge(sum, acc, cm); acc = sum;
            }
        }
    }
    const uint32_t badmask = __ballot_sync(0xffffffffu, bad || !good);
    __syncwarp();                                                            // all table use finished: the region becomes the sum scratch
    pts[lane] = acc;
    __syncwarp();
    for (int d = 4; d >= 1; d >>= 1) {
        if ((int)q < d) { gej t; gej_add(t, pts[lane], pts[lane + d]); pts[lane] = t; }
        __syncwarp();
    }
    if (q == 0 && valid) ok[item] = (uint8_t)(((badmask >> (8 * grp)) & 0xFFu) == 0 && pts[lane].inf);
}

cudaError_t launch_bppp_norm_verify(uint8_t* ok, const uint8_t* proofs, size_t stride, const uint32_t* plen, const uint8_t* transcripts,
                                    const uint8_t* rho32, const uint8_t* gens64, uint32_t g_len, const uint8_t* c_vec32, uint32_t h_len,
                                    const uint8_t* commits64, size_t n, const gtable_dev& gt, cudaStream_t s, uint64_t* launches, cudaMemPool_t pool) {
    if (n == 0) return cudaSuccess;
    static const size_t pack_min = [] {                                     // development knob: where the packed kernel takes over
        const char* v = getenv("SECP256K1_B200_BPPP_PACK_MIN");
        return v && *v ? (size_t)strtoull(v, nullptr, 10) : (size_t)BPK_MIN_PROOFS;
    }();
    if (n >= pack_min && g_len + h_len <= 128 && g_len && h_len) {
        const uint32_t count = g_len + h_len;
        const size_t tab_bytes = (size_t)count * BPK_GEN_WORDS * 4, base_bytes = (size_t)count * BPK_WINDOWS * sizeof(bpk_base);
        uint8_t* mem = nullptr;
        cudaError_t err = pool ? cudaMallocFromPoolAsync((void**)&mem, tab_bytes + base_bytes, pool, s) : cudaMallocAsync((void**)&mem, tab_bytes + base_bytes, s);
        if (err != cudaSuccess) return err;
        uint32_t* tabp = reinterpret_cast<uint32_t*>(mem);
        bpk_base* bases = reinterpret_cast<bpk_base*>(mem + tab_bytes);
        k_bppp_gen_bases<<<(count * 32 + 127) / 128, 128, 0, s>>>(bases, gens64, count);
        const size_t entries = (size_t)count * BPK_WINDOWS * 256;
        k_bppp_gen_entries<<<(unsigned)((entries + 127) / 128), 128, 0, s>>>(tabp, bases, count);
        const int smem_p = BPK_WARPS * BP_WARP_STRIDE;
        err = cudaFuncSetAttribute(k_bppp_norm_verify_packed, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p);
        if (err == cudaSuccess) {
            ecmult_gtable g; g.data = gt.data; g.gw = gt.gw; g.windows = gt.windows;
            const size_t warps = (n + 3) / 4;
            k_bppp_norm_verify_packed<<<(unsigned)((warps + BPK_WARPS - 1) / BPK_WARPS), BPK_WARPS * 32, smem_p, s>>>(ok, proofs, stride, plen, transcripts, rho32,
                                                                                                                gens64, g_len, c_vec32, h_len, commits64, n, g, tabp);
            err = cudaGetLastError();
        }
        cudaError_t ferr = cudaFreeAsync(mem, s);
        if (launches) __atomic_fetch_add(launches, (uint64_t)3, __ATOMIC_RELAXED);
        return err != cudaSuccess ? err : ferr;
    }
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
