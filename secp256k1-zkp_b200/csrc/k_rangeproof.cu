// k_rangeproof.cu -- rangeproof_verify (one warp per proof, one lane per ring) and pedersen_commit kernels.
#include "rewind.cuh"
#include "launch.h"

namespace b200 {

#ifndef RP_WARPS
#define RP_WARPS 4
#endif
// per warp in shared memory: the point-sum scratch (32 Jacobian points), m, R33[32].  The lanes' odd-multiples tables are in local
// memory (see k_verify.cu: more resident warps than the 3 blocks/SM that 16 KB of table per warp allowed).
#ifndef RP_MINBLOCKS
#define RP_MINBLOCKS 4
#endif
#define RP_TAB_WORDS ((32 * (int)sizeof(gej) + 3) / 4)
#define RP_WARP_BYTES (RP_TAB_WORDS * 4 + 32 + 32 * 33 + 32)
#define RP_WARP_STRIDE ((RP_WARP_BYTES + 127) & ~127)

// The two hashes over the whole proof -- the message hash m (commitment, generator, header, digit commitments, extra data: ~1.1 KB for a
// 64-bit proof) and the closing e0 check over the 32 ring results -- are serial chains of 17-18 compressions each.  Inside the warp-per-proof
// kernel they ran on lane 0 with 31 lanes masked (r2m: 6.5 % of the kernel's issue slots); here one THREAD per proof runs them, before and
// after the ring walks.  RP_SCRATCH_BYTES per proof of scratch carry m and the ring results between the three launches.
#define RP_SCRATCH_BYTES (32 + 33 * RP_MAX_RINGS)
__global__ void __launch_bounds__(128) k_rangeproof_prehash(uint8_t* __restrict__ scratch, const uint8_t* __restrict__ commit33, const uint8_t* __restrict__ proofs,
        size_t stride, const uint32_t* __restrict__ plen, const uint8_t* __restrict__ gen64, size_t gen_stride, const uint8_t* __restrict__ extra,
        size_t extra_stride, const uint32_t* __restrict__ extra_len, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* proof = proofs + i * stride;
    const size_t pl = plen[i];
    const uint8_t* ex = extra ? extra + extra_stride * i : nullptr;
    const size_t exl = (extra && extra_len) ? extra_len[i] : 0;
    if (pl > stride || (ex && exl > extra_stride)) return;          // k_rangeproof_verify rejects these before it looks at m
    rp_header h;
    uint64_t mn, mx;
    if (!rp_parse_header(h, mn, mx, proof, pl)) return;
    const uint8_t* g64 = gen64 + gen_stride * i;
    ge genp; rp_generator_load(genp, g64);
    uint8_t m32[32];
    rp_message_hash(m32, h, proof, commit33 + 33 * i, g64, fe_is_square(genp.y), ex, exl);
    uint8_t* out = scratch + RP_SCRATCH_BYTES * i;
    for (int k = 0; k < 32; k++) out[k] = m32[k];
}
// P5 for the proofs whose ring walks succeeded
__global__ void __launch_bounds__(128) k_rangeproof_final(uint8_t* __restrict__ ok, const uint8_t* __restrict__ scratch, const uint8_t* __restrict__ proofs,
        size_t stride, const uint32_t* __restrict__ plen, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !ok[i]) return;
    rp_header h;
    uint64_t mn, mx;
    const uint8_t* proof = proofs + i * stride;
    int good = rp_parse_header(h, mn, mx, proof, plen[i]);           // accepted by k_rangeproof_verify: cannot fail
    const uint8_t* scr = scratch + RP_SCRATCH_BYTES * i;
    ok[i] = (uint8_t)(good && rp_final_check(h, proof, scr + 32, scr));
}

__global__ void __launch_bounds__(RP_WARPS * 32, RP_MINBLOCKS) k_rangeproof_verify(uint8_t* __restrict__ ok, uint64_t* __restrict__ minv, uint64_t* __restrict__ maxv,
        const uint8_t* __restrict__ commit33, const uint8_t* __restrict__ proofs, size_t stride, const uint32_t* __restrict__ plen,
        const uint8_t* __restrict__ gen64, size_t gen_stride, const uint8_t* __restrict__ extra, size_t extra_stride,
        const uint32_t* __restrict__ extra_len, size_t n, ecmult_gtable gt, uint8_t* __restrict__ ev_out /* optional: n x 4096 bytes */,
        uint8_t* __restrict__ scratch /* n x RP_SCRATCH_BYTES: m from k_rangeproof_prehash in, ring results out */) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t item = (size_t)blockIdx.x * RP_WARPS + warp;
    if (item >= n) return;                                  // whole warp leaves together
    uint8_t* wbase = smem_raw + (size_t)warp * RP_WARP_STRIDE;
    gej* pts = reinterpret_cast<gej*>(wbase);
    uint8_t* m32 = wbase + RP_TAB_WORDS * 4;
    uint8_t* r33 = m32 + 32;
    const uint8_t* proof = proofs + item * stride;
    const size_t pl = plen[item];
    const uint8_t* c33 = commit33 + 33 * item;
    const uint8_t* g64 = gen64 + gen_stride * item;
    const uint8_t* ex = extra ? extra + extra_stride * item : nullptr;
    const size_t exl = (extra && extra_len) ? extra_len[item] : 0;
    // the _dev entry points take lengths from device memory: a length beyond its stride would read the neighbour's record
    if (pl > stride || (ex && exl > extra_stride)) { if (lane == 0) { ok[item] = 0; minv[item] = 0; maxv[item] = 0; } return; }

    rp_header h;
    uint64_t mn = 0, mx = 0;
    int hdr_ok = rp_parse_header(h, mn, mx, proof, pl);     // identical in all lanes
    if (lane == 0) { minv[item] = mn; maxv[item] = mx; }
    if (!hdr_ok) { if (lane == 0) ok[item] = 0; return; }

    // ---- P1: decompress digit commitments (lanes < rings-1) and the value commitment (lane rings-1)
    ge genp; rp_generator_load(genp, g64);
    ge mine; mine.inf = 1;
    int good = 1;
    if (lane + 1 < h.rings) good = rp_digit_commitment_load(mine, h, proof, lane);
    else if (lane + 1 == h.rings) good = rp_commitment_load(mine, c33);
    {
        gej t;
        if (lane + 1 < h.rings && good) gej_set_ge(t, mine); else gej_set_infinity(t);
        pts[lane] = t;
    }
    if (lane < 8) reinterpret_cast<uint32_t*>(m32)[lane] = reinterpret_cast<const uint32_t*>(scratch + RP_SCRATCH_BYTES * item)[lane];   // m (k_rangeproof_prehash)
    // a commitment that does not parse is rejected without looking at the proof: the reference caller never reaches
    // secp256k1_rangeproof_verify (secp256k1_pedersen_commitment_parse fails first), so min/max stay untouched (0 here)
    int commit_ok = __shfl_sync(0xffffffffu, good, (int)h.rings - 1);
    if (!commit_ok && lane == 0) { minv[item] = 0; maxv[item] = 0; }
    good = __all_sync(0xffffffffu, good);
    __syncwarp();
    // ---- P2: accj = min_value*H + sum C_i
    for (int d = 16; d >= 1; d >>= 1) {
        if ((int)lane < d) { gej t; gej_add(t, pts[lane], pts[lane + d]); pts[lane] = t; }
        __syncwarp();
    }
    gej pub; gej_set_infinity(pub);
    if (lane + 1 == h.rings && good) {
        gej acc = pts[0];
        if (h.min_value) { gej mh, t; rp_small_mult(mh, h.min_value, genp); gej_add(t, acc, mh); acc = t; }
        gej nacc; gej_neg(nacc, acc);
        gej_add_ge(pub, nacc, mine);                         // commit - accj (rangeproof_impl.h:627-631)
        if (pub.inf) good = 0;
    } else if (lane + 1 < h.rings) {
        gej_set_ge(pub, mine);
    }
    good = __all_sync(0xffffffffu, good);
    __syncwarp();
    if (!good) { if (lane == 0) ok[item] = 0; return; }
    // ---- P3/P4: ring walk
    int ring_ok = 1;
    if (lane < h.rings) {
        gej base;
        rp_ring_base(base, genp, h.exp, lane);
        tab_local tab;
        ring_ok = rp_ring_walk(r33 + 33 * lane, h, proof, m32, lane, pub, base, gt, tab, ev_out ? ev_out + 4096 * item : nullptr);
    }
    ring_ok = __all_sync(0xffffffffu, ring_ok);
    __syncwarp();
    // ---- the ring results go to k_rangeproof_final (P5)
    if (ring_ok) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(scratch + RP_SCRATCH_BYTES * item + 32);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(r33);
        for (uint32_t k = lane; k < (33 * h.rings + 3) / 4; k += 32) dst[k] = src[k];
    }
    if (lane == 0) ok[item] = (uint8_t)ring_ok;
}

// Second pass of a rewind: one thread per proof that k_rangeproof_verify accepted (ok[i] != 0 on entry).  Replays the prover's
// RFC 6979 stream from the nonce, recovers value / blinding factor / message and re-checks the commitment.
__global__ void __launch_bounds__(64) k_rangeproof_rewind(uint8_t* __restrict__ ok, uint8_t* __restrict__ blind32, uint64_t* __restrict__ value,
        uint8_t* __restrict__ msg, size_t msg_stride, uint32_t* __restrict__ msg_len, const uint8_t* __restrict__ nonce32,
        const uint8_t* __restrict__ commit33, const uint8_t* __restrict__ proofs, size_t stride, const uint32_t* __restrict__ plen,
        const uint8_t* __restrict__ gen64, size_t gen_stride, const uint8_t* __restrict__ ev, uint8_t* __restrict__ scratch /* n x 5120 */, size_t n,
        ecmult_gtable gt) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int good = ok[i];
    uint8_t bl[32];
    uint64_t v = 0;
    size_t ml = 0;
    for (int k = 0; k < 32; k++) bl[k] = 0;
    if (good) {
        rp_header h;
        uint64_t mn, mx;
        const uint8_t* proof = proofs + i * stride;
        good = rp_parse_header(h, mn, mx, proof, plen[i]);          // accepted before: cannot fail
        ge genp; rp_generator_load(genp, gen64 + gen_stride * i);
        if (good) good = rp_rewind_item(bl, v, msg ? msg + msg_stride * i : nullptr, msg ? msg_stride : 0, ml, h, proof, commit33 + 33 * i,
                                        gen64 + gen_stride * i, fe_is_square(genp.y), nonce32 + 32 * i, ev + 4096 * i, gt,
                                        reinterpret_cast<sc*>(scratch + 5120 * i + 4096), scratch + 5120 * i);
    }
    if (!good) { v = 0; ml = 0; for (int k = 0; k < 32; k++) bl[k] = 0; }
    ok[i] = (uint8_t)good;
    for (int k = 0; k < 32; k++) blind32[32 * i + k] = bl[k];
    value[i] = v;
    if (msg_len) msg_len[i] = (uint32_t)ml;
}

__global__ void __launch_bounds__(128) k_pedersen_commit(uint8_t* __restrict__ ok, uint8_t* __restrict__ out33, const uint8_t* __restrict__ blind32,
        const uint64_t* __restrict__ value, const uint8_t* __restrict__ gen64, size_t gen_stride, size_t n, ecmult_gtable gt) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t b[32], g[64], o[33];
    for (int k = 0; k < 32; k++) b[k] = blind32[32 * i + k];
    for (int k = 0; k < 64; k++) g[k] = gen64[gen_stride * i + k];
    int r = pedersen_commit_item(o, b, value[i], g, gt);
    for (int k = 0; k < 33; k++) out33[33 * i + k] = o[k];
    ok[i] = (uint8_t)r;
}

static inline ecmult_gtable to_gt(const gtable_dev& g) { ecmult_gtable t; t.data = g.data; t.gw = g.gw; t.windows = g.windows; return t; }

cudaError_t launch_rangeproof_verify(uint8_t* ok, uint64_t* minv, uint64_t* maxv, const uint8_t* commit33, const uint8_t* proofs, size_t stride,
                                     const uint32_t* plen, const uint8_t* gen64, size_t gen_stride, const uint8_t* extra, size_t extra_stride,
                                     const uint32_t* extra_len, size_t n, const gtable_dev& gt, cudaStream_t s, uint64_t* launches, uint8_t* ev_out, cudaMemPool_t pool) {
    if (n == 0) return cudaSuccess;
#ifdef RP_SMEM_MIN                                                 // development: caps the resident blocks per SM (A/B of wave shapes)
    const int smem = RP_WARPS * RP_WARP_STRIDE > RP_SMEM_MIN ? RP_WARPS * RP_WARP_STRIDE : RP_SMEM_MIN;
#else
    const int smem = RP_WARPS * RP_WARP_STRIDE;
#endif
    cudaError_t err = cudaFuncSetAttribute(k_rangeproof_verify, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    uint8_t* scratch = nullptr;                                   // stream-ordered: no synchronisation; the engine's pool keeps the block cached call after call
    err = pool ? cudaMallocFromPoolAsync((void**)&scratch, RP_SCRATCH_BYTES * n, pool, s) : cudaMallocAsync((void**)&scratch, RP_SCRATCH_BYTES * n, s);
    if (err != cudaSuccess) return err;
    const unsigned tb = (unsigned)((n + 127) / 128);
    k_rangeproof_prehash<<<tb, 128, 0, s>>>(scratch, commit33, proofs, stride, plen, gen64, gen_stride, extra, extra_stride, extra_len, n);
    k_rangeproof_verify<<<(unsigned)((n + RP_WARPS - 1) / RP_WARPS), RP_WARPS * 32, smem, s>>>(ok, minv, maxv, commit33, proofs, stride, plen,
                                                                                            gen64, gen_stride, extra, extra_stride, extra_len, n, to_gt(gt), ev_out, scratch);
    k_rangeproof_final<<<tb, 128, 0, s>>>(ok, scratch, proofs, stride, plen, n);
    err = cudaGetLastError();
    cudaError_t ferr = cudaFreeAsync(scratch, s);
    if (launches) __atomic_fetch_add(launches, (uint64_t)3, __ATOMIC_RELAXED);
    return err != cudaSuccess ? err : ferr;
}
cudaError_t launch_rangeproof_rewind(uint8_t* ok, uint8_t* blind32, uint64_t* value, uint8_t* msg, size_t msg_stride, uint32_t* msg_len,
                                     const uint8_t* nonce32, const uint8_t* commit33, const uint8_t* proofs, size_t stride, const uint32_t* plen,
                                     const uint8_t* gen64, size_t gen_stride, const uint8_t* ev, uint8_t* scratch, size_t n, const gtable_dev& gt,
                                     cudaStream_t s, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    k_rangeproof_rewind<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(ok, blind32, value, msg, msg_stride, msg_len, nonce32, commit33, proofs, stride, plen,
                                                                 gen64, gen_stride, ev, scratch, n, to_gt(gt));
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}
cudaError_t launch_pedersen_commit(uint8_t* ok, uint8_t* out33, const uint8_t* blind32, const uint64_t* value, const uint8_t* gen64, size_t gen_stride,
                                   size_t n, const gtable_dev& gt, cudaStream_t s, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    k_pedersen_commit<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(ok, out33, blind32, value, gen64, gen_stride, n, to_gt(gt));
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

}  // namespace b200
