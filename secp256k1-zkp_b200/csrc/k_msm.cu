// k_msm.cu -- Pippenger multi-scalar multiplication kernels (see msm.cuh for the algorithm and what it replaces).
//
// Launch sequence of one MSM (all on the caller's stream, no host synchronisation, every size decided on the device):
//   memset            hist / counters / buckets
//   k_msm_prepare     thread per term: parse, GLV split, signed digits -> (window,bucket) histogram
//   k_msm_scan_a/_b   two-level exclusive scan of the histogram: entry offsets, piece offsets, part offsets, list of multi-piece buckets
//   k_msm_scatter     thread per term: counting-sort scatter of (term index | sign) by (window,bucket)
//   k_msm_units       thread per piece: piece -> (first entry, count, destination)
//   k_msm_bucket_sum  persistent warps take batches of pieces from a global queue; one lane per piece, lanes pull the next piece when theirs
//                     is done (mixed additions)
//   k_msm_combine     block per multi-piece (hot) bucket: strided partial sums + warp-shuffle tree
//   k_msm_r1          running sums over chunks of 8 buckets (thread per chunk, or 4 lanes per chunk when chunks are few)
//   k_msm_r2          block per (window, sum): plain sum of the chunk L's, bit sums of the chunk A's
//   k_msm_tail        one block: per-window bit-Horner (a warp per window), two window-Horner chains (a warp per GLV half), join, to affine --
//                     all in the lane form of ge_w10.cuh (one 26-bit limb of every coordinate per lane)
#ifndef MSM_G4_OUTLINED
#define B200_G4_INLINE 1          // the 4-lane group operations inlined into the (few) latency-bound loops of this file
#endif
#include "msm.cuh"
#include "ge_g4.cuh"
#include "launch.h"

namespace b200 {

struct msm_ws {
    uint32_t* k12;       // [N][8]     |k1|, |k2| (4 + 4 LE limbs)
    uint8_t* flags;      // [N]
    uint8_t* gpoint;     // [64]       generator, x||y big-endian (term N-1)
    uint32_t* hist;      // [TB]       counts; the scatter counts them back down to zero
    uint32_t* off;       // [TB+1]     first sorted entry of a bucket
    uint32_t* pbase;     // [TB+1]     first piece (unit) of a bucket
    uint32_t* mbase;     // [TB+1]     first slot in `parts` of a multi-piece bucket
    uint32_t* blocktot;  // [nblk][4]
    uint32_t* counters;  // [4 + 2*nw*MSM_MAX_SUMS]  [0] = number of large multi-piece buckets (> MSM_SMALL_MULTI pieces), [1] = small ones,
                         //            [4 + sum] = blocks of that plain / bit sum that have finished (k_msm_r2)
    uint32_t* multi;     // [cap_multi]  large multi-piece buckets
    uint32_t* multi_s;   // [cap_multi]  small multi-piece buckets
    uint32_t* entries;   // [2*nw*N]
    msm_unit* units;     // [cap_units]
    gej_store* buckets;  // [TB]
    gej_store* parts;    // [cap_parts]
    gej_store* Lq;       // [TB/8]
    gej_store* Aq;       // [TB/8]
    gej_store* T;        // [2*nw][MSM_MAX_SUMS][MSM_R2_PARTS]
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
struct msm_caps { size_t emax, cap_multi, cap_parts, cap_units, nblk; };
static msm_caps caps_of(const msm_plan& p) {
    msm_caps c;
    c.emax = (size_t)msm_windows_owned(p) * p.N;              // one entry per term and summed window at most (all 2 nw windows unless sharded)
    if (c.emax == 0) c.emax = 1;
    c.cap_multi = c.emax / p.Tp + 1;
    c.cap_parts = 2 * (c.emax / p.Tp) + 2;
    c.cap_units = (c.emax < p.TB ? c.emax : p.TB) + c.emax / p.Tp + 1;
    c.nblk = (p.TB + MSM_SCAN_ELEMS - 1) / MSM_SCAN_ELEMS;
    return c;
}
static size_t carve(msm_ws* w, uint8_t* base, const msm_plan& p) {
    const msm_caps c = caps_of(p);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += align256(bytes); return base ? base + at : (uint8_t*)nullptr; };
    uint8_t* q;
    q = take((size_t)p.N * 32); if (w) w->k12 = (uint32_t*)q;
    q = take(p.N); if (w) w->flags = q;
    q = take(64); if (w) w->gpoint = q;
    q = take((size_t)p.TB * 4); if (w) w->hist = (uint32_t*)q;           // hist and counters are adjacent: one memset
    q = take(16 + (size_t)2 * p.nw * MSM_MAX_SUMS * 4); if (w) w->counters = (uint32_t*)q;
    q = take((size_t)(p.TB + 1) * 4); if (w) w->off = (uint32_t*)q;
    q = take((size_t)(p.TB + 1) * 4); if (w) w->pbase = (uint32_t*)q;
    q = take((size_t)(p.TB + 1) * 4); if (w) w->mbase = (uint32_t*)q;
    q = take(c.nblk * 16); if (w) w->blocktot = (uint32_t*)q;
    q = take(c.cap_multi * 4); if (w) w->multi = (uint32_t*)q;
    q = take(c.cap_multi * 4); if (w) w->multi_s = (uint32_t*)q;
    q = take(c.emax * 4); if (w) w->entries = (uint32_t*)q;
    q = take(c.cap_units * sizeof(msm_unit)); if (w) w->units = (msm_unit*)q;
    q = take((size_t)p.TB * sizeof(gej_store)); if (w) w->buckets = (gej_store*)q;
    q = take(c.cap_parts * sizeof(gej_store)); if (w) w->parts = (gej_store*)q;
    q = take((size_t)(p.TB / MSM_CHUNK) * sizeof(gej_store)); if (w) w->Lq = (gej_store*)q;
    q = take((size_t)(p.TB / MSM_CHUNK) * sizeof(gej_store)); if (w) w->Aq = (gej_store*)q;
    q = take((size_t)2 * p.nw * MSM_MAX_SUMS * MSM_R2_PARTS * sizeof(gej_store)); if (w) w->T = (gej_store*)q;
    return o;
}
// window widths a caller may force (secp256k1_b200_ecmult_multi_set_window); the workspace is sized for the worst of them
static bool c_allowed(size_t N, int c) { return c >= (N <= 4096 ? MSM_C_MIN : 8) && c <= MSM_C_MAX; }
size_t msm_workspace_bytes(size_t n_terms) {
    size_t best = 0;
    for (int c = MSM_C_MIN; c <= MSM_C_MAX; c++) {
        if (!c_allowed(n_terms + 1, c)) continue;
        msm_plan p;
        msm_make_plan_c(p, n_terms + 1, (uint32_t)c);
        size_t b = carve(nullptr, nullptr, p);
        if (b > best) best = b;
    }
    return best;
}

// ---- (1) parse + GLV split + histogram ---------------------------------------------------------------------------------------
struct emit_hist {
    uint32_t* hist;
    HD void operator()(uint32_t bucket, uint32_t) {
#if defined(__CUDA_ARCH__)
        atomicAdd(&hist[bucket], 1u);
#else
        hist[bucket]++;
#endif
    }
};
__global__ void __launch_bounds__(256) k_msm_prepare(msm_ws ws, msm_plan p, const uint8_t* __restrict__ scalars32, const uint8_t* __restrict__ points64,
                                                     const uint8_t* __restrict__ g_scalar32, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    sc k;
    int inf;
    __align__(16) uint8_t b[64];
    if (i < n) {
        const uint8_t* s = scalars32 + 32 * i;
        if ((reinterpret_cast<uintptr_t>(s) & 15) == 0) { const uint4* q = reinterpret_cast<const uint4*>(s); uint4* d = reinterpret_cast<uint4*>(b); d[0] = __ldg(q); d[1] = __ldg(q + 1); }
        else for (int t = 0; t < 32; t++) b[t] = __ldg(s + t);
        sc_set_b32(k, b);
        const uint8_t* pt = points64 + 64 * i;
        uint32_t z = 0;
        if ((reinterpret_cast<uintptr_t>(pt) & 15) == 0) {
            const uint4* q = reinterpret_cast<const uint4*>(pt);
#pragma unroll
            for (int t = 0; t < 4; t++) { uint4 v = __ldg(q + t); z |= v.x | v.y | v.z | v.w; }
        } else for (int t = 0; t < 64; t++) z |= __ldg(pt + t);
        inf = (z == 0);
    } else {
        if (g_scalar32) { for (int t = 0; t < 32; t++) b[t] = g_scalar32[t]; sc_set_b32(k, b); } else sc_set_zero(k);
        inf = 0;
        ge g; ge_set_generator(g);
        fe_get_b32(b, g.x); fe_get_b32(b + 32, g.y);
        for (int t = 0; t < 64; t++) ws.gpoint[t] = b[t];
    }
    uint32_t k12[8], flags;
    msm_prepare_term(k12, flags, k, inf);
    uint4* o = reinterpret_cast<uint4*>(ws.k12 + 8 * i);
    o[0] = make_uint4(k12[0], k12[1], k12[2], k12[3]);
    o[1] = make_uint4(k12[4], k12[5], k12[6], k12[7]);
    ws.flags[i] = (uint8_t)flags;
    if (flags & 4u) return;
    emit_hist em = {ws.hist};
    msm_emit_digits(p, k12, flags & 1u, 0, em);
    msm_emit_digits(p, k12 + 4, (flags >> 1) & 1u, 1, em);
}

// ---- (2) scan: offsets of entries / pieces / parts ---------------------------------------------------------------------------
struct scan3 { uint32_t a, b, c; };
__device__ __forceinline__ scan3 scan3_add(scan3 x, scan3 y) { scan3 r = {x.a + y.a, x.b + y.b, x.c + y.c}; return r; }
// exclusive scan of one scan3 per thread across a 1024-thread block; returns the block total through `total`
__device__ scan3 block_exclusive_scan3(scan3 v, scan3& total) {
    __shared__ scan3 warp_tot[32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    scan3 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        scan3 o;
        o.a = __shfl_up_sync(0xffffffffu, inc.a, d); o.b = __shfl_up_sync(0xffffffffu, inc.b, d); o.c = __shfl_up_sync(0xffffffffu, inc.c, d);
        if ((int)lane >= d) inc = scan3_add(inc, o);
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        scan3 w = {0, 0, 0};
        if (lane < nw) w = warp_tot[lane];
        scan3 winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            scan3 o;
            o.a = __shfl_up_sync(0xffffffffu, winc.a, d); o.b = __shfl_up_sync(0xffffffffu, winc.b, d); o.c = __shfl_up_sync(0xffffffffu, winc.c, d);
            if ((int)lane >= d) winc = scan3_add(winc, o);
        }
        warp_tot[lane] = winc;                      // inclusive over warps
    }
    __syncthreads();
    total = warp_tot[nw - 1];
    scan3 base = {0, 0, 0};
    if (wid > 0) base = warp_tot[wid - 1];
    scan3 ex = {base.a + inc.a - v.a, base.b + inc.b - v.b, base.c + inc.c - v.c};
    __syncthreads();
    return ex;
}
__device__ __forceinline__ uint32_t msm_pieces(uint32_t cnt, uint32_t Tp) { return (cnt + Tp - 1) / Tp; }

__global__ void __launch_bounds__(1024) k_msm_scan_a(msm_ws ws, msm_plan p) {
    const uint32_t first = blockIdx.x * MSM_SCAN_ELEMS + threadIdx.x * 4;
    uint32_t cnt[4];
    scan3 mine = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        cnt[k] = (first + k < p.TB) ? ws.hist[first + k] : 0;
        uint32_t pc = msm_pieces(cnt[k], p.Tp);
        mine.a += cnt[k]; mine.b += pc; mine.c += (pc > 1 ? pc : 0);
    }
    scan3 total;
    scan3 ex = block_exclusive_scan3(mine, total);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (first + k < p.TB) { ws.off[first + k] = ex.a; ws.pbase[first + k] = ex.b; ws.mbase[first + k] = ex.c; }
        uint32_t pc = msm_pieces(cnt[k], p.Tp);
        ex.a += cnt[k]; ex.b += pc; ex.c += (pc > 1 ? pc : 0);
    }
    if (threadIdx.x == 0) { ws.blocktot[4 * blockIdx.x] = total.a; ws.blocktot[4 * blockIdx.x + 1] = total.b; ws.blocktot[4 * blockIdx.x + 2] = total.c; }
}
__global__ void __launch_bounds__(1024) k_msm_scan_b(msm_ws ws, msm_plan p) {
    // base of this block = sum of the totals of the blocks before it (<= 288 of them: every thread strides, then a block reduction)
    scan3 part = {0, 0, 0};
    for (uint32_t b = threadIdx.x; b < blockIdx.x; b += blockDim.x) { part.a += ws.blocktot[4 * b]; part.b += ws.blocktot[4 * b + 1]; part.c += ws.blocktot[4 * b + 2]; }
    scan3 base;
    block_exclusive_scan3(part, base);
    const uint32_t first = blockIdx.x * MSM_SCAN_ELEMS + threadIdx.x * 4;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t i = first + k;
        if (i >= p.TB) break;
        ws.off[i] += base.a; ws.pbase[i] += base.b; ws.mbase[i] += base.c;
        const uint32_t pc = msm_pieces(ws.hist[i], p.Tp);
        if (pc > MSM_SMALL_MULTI) ws.multi[atomicAdd(&ws.counters[0], 1u)] = i;
        else if (pc > 1) ws.multi_s[atomicAdd(&ws.counters[1], 1u)] = i;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        ws.off[p.TB] = base.a + ws.blocktot[4 * blockIdx.x];
        ws.pbase[p.TB] = base.b + ws.blocktot[4 * blockIdx.x + 1];
        ws.mbase[p.TB] = base.c + ws.blocktot[4 * blockIdx.x + 2];
    }
}

// ---- (3) counting-sort scatter -----------------------------------------------------------------------------------------------
struct emit_scatter {
    const uint32_t* off; uint32_t* hist; uint32_t* entries; uint32_t idx;
    HD void operator()(uint32_t bucket, uint32_t sgn) {
#if defined(__CUDA_ARCH__)
        uint32_t pos = off[bucket] + atomicSub(&hist[bucket], 1u) - 1u;
#else
        uint32_t pos = off[bucket] + (hist[bucket]--) - 1u;
#endif
        entries[pos] = idx | (sgn << 31);
    }
};
__global__ void __launch_bounds__(256) k_msm_scatter(msm_ws ws, msm_plan p) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    uint32_t flags = ws.flags[i];
    if (flags & 4u) return;
    const uint4* q = reinterpret_cast<const uint4*>(ws.k12 + 8 * i);
    uint4 a = q[0], b = q[1];
    uint32_t k1[4] = {a.x, a.y, a.z, a.w}, k2[4] = {b.x, b.y, b.z, b.w};
    emit_scatter em = {ws.off, ws.hist, ws.entries, (uint32_t)i};
    msm_emit_digits(p, k1, flags & 1u, 0, em);
    msm_emit_digits(p, k2, (flags >> 1) & 1u, 1, em);
}

// ---- (3b) pieces -> unit descriptors -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_msm_units(msm_ws ws, msm_plan p) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t P = ws.pbase[p.TB];
    if (u >= P) return;
    // largest bucket b with pbase[b] <= u  (empty buckets share their pbase with the next one; the last of a run is the non-empty one)
    uint32_t lo = 0, hi = p.TB;                    // invariant: pbase[lo] <= u < pbase[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (ws.pbase[mid] <= u) lo = mid; else hi = mid;
    }
    const uint32_t b = lo, j = u - ws.pbase[b];
    const uint32_t beg = ws.off[b], end = ws.off[b + 1], npieces = ws.pbase[b + 1] - ws.pbase[b];
    msm_unit d;
    d.begin = beg + j * p.Tp;
    d.count = (end - d.begin < p.Tp) ? end - d.begin : p.Tp;
    d.dst = (npieces == 1) ? b : (MSM_DST_PART | (ws.mbase[b] + j));
    d.first = ws.entries[d.begin];
    ws.units[u] = d;
}

// ---- (4) bucket sums: one lane per piece, lanes of a warp pull pieces dynamically --------------------------------------------
__device__ __forceinline__ void store_gej_vec(gej_store* dst, const gej& a) {
    gej_store s;
    msm_store_gej(s, a);
    uint4* o = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int k = 0; k < 6; k++) o[k] = make_uint4(s.w[4 * k], s.w[4 * k + 1], s.w[4 * k + 2], s.w[4 * k + 3]);
}
__device__ __forceinline__ void load_gej_vec(gej& a, const gej_store* src) {
    gej_store s;
    const uint4* q = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int k = 0; k < 6; k++) { uint4 v = q[k]; s.w[4 * k] = v.x; s.w[4 * k + 1] = v.y; s.w[4 * k + 2] = v.z; s.w[4 * k + 3] = v.w; }
    msm_load_gej(a, s);
}
__device__ __forceinline__ void load_gej_cg(gej& a, const gej_store* src) {      // L2 load: data written by another block of this kernel
    gej_store s;
    const uint4* q = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int k = 0; k < 6; k++) { uint4 v = __ldcg(q + k); s.w[4 * k] = v.x; s.w[4 * k + 1] = v.y; s.w[4 * k + 2] = v.z; s.w[4 * k + 3] = v.w; }
    msm_load_gej(a, s);
}
#ifndef MSM_SUM_MINBLOCKS
#define MSM_SUM_MINBLOCKS 4        // measured (profiles/r2b): inlined addition at 4 blocks/SM 3.97 ms per 2^20-term MSM, 5 blocks 4.04, outlined/5 4.22
#endif
#ifndef MSM_SUM_OUTLINED_ADD
#define MSM_SUM_INLINE_ADD 1
#endif
#ifndef MSM_SUM_STATIC_RANGES
#define MSM_SUM_QUEUE 1            // r2r: 2^17 terms 1.089 -> 1.065 ms, 2^20 terms 3.439 -> 3.422 ms against one contiguous range per warp
#endif
#ifndef MSM_SUM_BATCH
#define MSM_SUM_BATCH 32           // at most this many pieces per take (64: no gain at 2^20, 128: slower everywhere)
#endif
// (r2o: an accumulator in XYZZ coordinates -- madd-2008-s, 8M + 2S against 8M + 3S -- was measured and rejected: the fourth coordinate
// pushes the loop over 128 registers, 3.95 ms against 3.45 ms for the 2^20-term MSM.)
__global__ void __launch_bounds__(128, MSM_SUM_MINBLOCKS) k_msm_bucket_sum(msm_ws ws, msm_plan p, const uint8_t* __restrict__ points64, size_t n) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t P = ws.pbase[p.TB];
#ifdef MSM_SUM_QUEUE
    // the pieces are a global queue (counters[2]) from which a warp takes MSM_SUM_BATCH at a time: warps whose pieces are long (the top window of
    // a half: several times fuller buckets) simply take fewer batches, and all warps run dry together
    (void)wid;
    uint32_t batch = P / (4 * nwarps);                            // small enough that every warp gets several batches
    batch = batch < 1 ? 1 : (batch > MSM_SUM_BATCH ? MSM_SUM_BATCH : batch);
    uint32_t next = 0, uend = 0;
    bool drained = false;
#else
    const uint32_t R = (P + nwarps - 1) / nwarps;                 // pieces per warp (contiguous range)
    uint64_t first = (uint64_t)wid * R;
    if (first >= P) return;                                       // whole warp
    uint32_t next = (uint32_t)first;
    const uint32_t uend = (first + R < P) ? (uint32_t)(first + R) : P;
#endif
    bool have = false;
    uint32_t e = 0, eend = 0, dst = 0;
    gej acc;
    gej_set_infinity(acc);
    ge nxt;                                                       // the point of entry e, loaded one iteration ahead of its addition
    uint32_t nxt_neg = 0;                                         // its sign, applied when it is consumed (nothing waits on the load before that)
    nxt.inf = 1;
    while (true) {
#ifdef MSM_SUM_QUEUE
        uint32_t mask = __ballot_sync(0xffffffffu, !have);
        while (mask) {                                                // warp-uniform: serve the lanes without a piece, refilling the batch when it runs dry
            if (next >= uend) {
                if (drained) break;
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&ws.counters[2], batch);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= P) { drained = true; break; }
                next = base;
                uend = (P - base < batch) ? P : base + batch;
            }
            const uint32_t avail = uend - next, rank = __popc(mask & ((1u << lane) - 1u));
            if (!have && rank < avail) {
                const uint4 d = *reinterpret_cast<const uint4*>(ws.units + next + rank);
                e = d.x; eend = d.x + d.y; dst = d.z;
                const uint32_t v = d.w;                               // = ws.entries[e], copied into the descriptor by k_msm_units
                // (the first addition of a piece is "infinity + P", a copy: r2x tried making the first point the accumulator right here instead --
                // the load it then has to wait for costs more than the idle lane-iteration it saves, 3.52 ms against 3.42 for 2^20 terms)
                gej_set_infinity(acc);
                msm_load_point(nxt, points64, n, ws.gpoint, v & 0x7FFFFFFFu);
                nxt_neg = v >> 31;
                have = true;
            }
            const uint32_t need = __popc(mask);
            next += need < avail ? need : avail;
            mask = __ballot_sync(0xffffffffu, !have);
        }
#else
        const uint32_t mask = __ballot_sync(0xffffffffu, !have);
        if (!have) {
            uint32_t my = next + __popc(mask & ((1u << lane) - 1u));
            if (my < uend) {
                const uint4 d = *reinterpret_cast<const uint4*>(ws.units + my);
                e = d.x; eend = d.x + d.y; dst = d.z;
                gej_set_infinity(acc);
                const uint32_t v = d.w;                               // = ws.entries[e], copied into the descriptor by k_msm_units
                msm_load_point(nxt, points64, n, ws.gpoint, v & 0x7FFFFFFFu);
                nxt_neg = v >> 31;
                have = true;
            }
        }
        next += __popc(mask);
        if (next > uend) next = uend;
#endif
        if (!__any_sync(0xffffffffu, have)) break;
        if (have) {
            ge pt = nxt;
            if (nxt_neg) fe_neg(pt.y, pt.y);
            if (e + 1 < eend) {                                   // issue the next gather now; the addition below hides its latency
                const uint32_t v = ws.entries[e + 1];
                msm_load_point(nxt, points64, n, ws.gpoint, v & 0x7FFFFFFFu);
                nxt_neg = v >> 31;
            }
            gej t;
#ifdef MSM_SUM_INLINE_ADD
            gej_add_ge_t<true>(t, acc, pt, nullptr);      // field multiplications inlined: this loop is the whole kernel, the body fits the I-cache
#else
            gej_add_ge(t, acc, pt);
#endif
            acc = t;
            if (++e == eend) {
                store_gej_vec((dst & MSM_DST_PART) ? ws.parts + (dst & ~MSM_DST_PART) : ws.buckets + dst, acc);
                have = false;
            }
        }
    }
}

// ---- (4b) multi-piece buckets: block per bucket, strided sums, warp-shuffle tree ----------------------------------------------
__device__ __forceinline__ void gej_shfl_down(gej& o, const gej& a, int delta) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
        o.x.v[k] = __shfl_down_sync(0xffffffffu, a.x.v[k], delta);
        o.y.v[k] = __shfl_down_sync(0xffffffffu, a.y.v[k], delta);
        o.z.v[k] = __shfl_down_sync(0xffffffffu, a.z.v[k], delta);
    }
    o.inf = __shfl_down_sync(0xffffffffu, a.inf, delta);
}
// sum of `acc` over the block's threads, result valid in thread 0.  scratch: one gej_store per warp.
__device__ void block_sum_gej(gej& acc, gej_store* scratch) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        gej o, t;
        gej_shfl_down(o, acc, d);
        if ((int)lane < d) { gej_add(t, acc, o); acc = t; }
    }
    __syncthreads();
    if (lane == 0) msm_store_gej(scratch[wid], acc);
    __syncthreads();
    if (wid == 0) {                                                     // the <= 8 warp sums: a second shuffle tree (3 additions deep, not 7 in a row)
        if (lane < nw) msm_load_gej(acc, scratch[lane]); else gej_set_infinity(acc);
#pragma unroll 1
        for (int d = 4; d >= 1; d >>= 1) {
            gej o, t;
            gej_shfl_down(o, acc, d);
            if ((int)lane < d) { gej_add(t, acc, o); acc = t; }
        }
    }
}
__global__ void __launch_bounds__(256) k_msm_combine(msm_ws ws, msm_plan p) {
    __shared__ gej_store scratch[8];
    const uint32_t nm = ws.counters[0];
    for (uint32_t i = blockIdx.x; i < nm; i += gridDim.x) {
        const uint32_t b = ws.multi[i];
        const uint32_t K = ws.pbase[b + 1] - ws.pbase[b], base = ws.mbase[b];
        gej acc;
        gej_set_infinity(acc);
#pragma unroll 1
        for (uint32_t k = threadIdx.x; k < K; k += blockDim.x) { gej o, t; load_gej_vec(o, ws.parts + base + k); gej_add(t, acc, o); acc = t; }
        block_sum_gej(acc, scratch);
        if (threadIdx.x == 0) store_gej_vec(ws.buckets + b, acc);
        __syncthreads();
    }
}

// small multi-piece buckets (2 .. MSM_SMALL_MULTI pieces): one thread each
__global__ void __launch_bounds__(128) k_msm_combine_small(msm_ws ws, msm_plan p) {
    const uint32_t nm = ws.counters[1];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nm; i += gridDim.x * blockDim.x) {
        const uint32_t b = ws.multi_s[i];
        const uint32_t K = ws.pbase[b + 1] - ws.pbase[b], base = ws.mbase[b];
        gej acc;
        load_gej_vec(acc, ws.parts + base);
#pragma unroll 1
        for (uint32_t k = 1; k < K; k++) { gej o, t; load_gej_vec(o, ws.parts + base + k); gej_add(t, acc, o); acc = t; }
        store_gej_vec(ws.buckets + b, acc);
    }
}

// ---- (5) running sums over chunks of 8 buckets --------------------------------------------------------------------------------
// (thread form: 158 registers unbounded = 3 blocks per SM and 1.3 waves of blocks at c = 16; bounded to 128 registers: 4 blocks, one wave)
template <bool G4>
__global__ void __launch_bounds__(128, G4 ? 1 : 4) k_msm_r1(msm_ws ws, msm_plan p) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nchunks = p.TB / MSM_CHUNK;
    uint32_t q = G4 ? (t >> 2) : t;
    const bool valid = q < nchunks;
    if (!G4 && !valid) return;
    if (!valid) q = nchunks - 1;                      // G4: the whole warp stays (full-mask shuffles); surplus groups redo the last chunk
    if (!G4 && p.shards > 1) {                        // window sharding: the chunks of another shard's windows are not summed here (k_msm_r2 skips them too)
        msm_win cw;
        msm_window_of_bucket(cw, p, q * MSM_CHUNK);
        if (!msm_window_owned(p, cw.half * p.nw + cw.w)) return;
    }
    const gej_store* s8 = ws.buckets + (size_t)q * MSM_CHUNK;
    gej running, acc;
    gej_set_infinity(running);
    gej_set_infinity(acc);
#pragma unroll 1
    for (int j = MSM_CHUNK - 1; j >= 0; j--) {
        gej s, tt;
        load_gej_vec(s, s8 + j);
        if (G4) { gej_add_g4(tt, running, s); running = tt; gej_add_g4(tt, acc, running); acc = tt; }
        else { gej_add(tt, running, s); running = tt; gej_add(tt, acc, running); acc = tt; }
    }
    if (valid && (!G4 || (threadIdx.x & 3) == 0)) { store_gej_vec(ws.Lq + q, acc); store_gej_vec(ws.Aq + q, running); }
}

// ---- (6) per window: plain sum of the L's, bit sums of the A's; <= MSM_R2_ELEMS chunk sums per block ----------------------------
__global__ void __launch_bounds__(256, 2) k_msm_r2(msm_ws ws, msm_plan p) {      // 2 blocks per SM (unbounded: 146 registers = 1 block, 3 waves)
    __shared__ gej_store scratch[8];
    const uint32_t part = blockIdx.x % MSM_R2_PARTS, ws_i = blockIdx.x / MSM_R2_PARTS;
    const uint32_t wi = ws_i / MSM_MAX_SUMS, s = ws_i % MSM_MAX_SUMS;
    if (!msm_window_owned(p, wi)) return;                               // window sharding: another launch sums this window (block-uniform)
    msm_win w;
    msm_window_by_index(w, p, wi);
    const uint32_t Q = w.Bw / MSM_CHUNK, nbits = msm_ilog2(Q), cb = w.base / MSM_CHUNK;
    const uint32_t count = (s == 0) ? Q : Q / 2;
    const uint32_t lo = part * MSM_R2_ELEMS, hi = (lo + MSM_R2_ELEMS < count) ? lo + MSM_R2_ELEMS : count;
    gej acc;
    gej_set_infinity(acc);
    if (s <= nbits && lo < count) {                                     // block-uniform
#pragma unroll 1
        for (uint32_t k = lo + threadIdx.x; k < hi; k += blockDim.x) {
            gej o, t;
            if (s == 0) load_gej_vec(o, ws.Lq + cb + k);
            else load_gej_vec(o, ws.Aq + cb + msm_bit_index(k, s - 1));
            gej_add(t, acc, o); acc = t;
        }
        block_sum_gej(acc, scratch);
    }
    // the blocks of one sum leave their partial results in T[sum][part]; the last one to finish adds them into T[sum][0]
    const uint32_t nparts = (s <= nbits) ? (count + MSM_R2_ELEMS - 1) / MSM_R2_ELEMS : 0;
    if (threadIdx.x == 0 && part < (nparts ? nparts : 1)) {
        gej_store* slot = ws.T + (size_t)ws_i * MSM_R2_PARTS;
        if (nparts <= 1) { store_gej_vec(slot, acc); return; }
        store_gej_vec(slot + part, acc);
        __threadfence();
        if (atomicAdd(&ws.counters[4 + ws_i], 1u) == nparts - 1) {
            __threadfence();
            gej tot;
            gej_set_infinity(tot);
#pragma unroll 1
            for (uint32_t k = 0; k < nparts; k++) { gej o, t; load_gej_cg(o, slot + k); gej_add(t, tot, o); tot = t; }
            store_gej_vec(slot, tot);
        }
    }
}

// ---- (7) tail: window sums, Horner, join, output ------------------------------------------------------------------------------
// mode 0: out = affine x||y big-endian (64 bytes, zeros = infinity); mode 1: out = gej_store (96 bytes) for a later k_gej_sum
__device__ void write_result(uint8_t* out, const gej& r, int mode) {
    if (mode == 1) { gej_store s; msm_store_gej(s, r); uint32_t* o = reinterpret_cast<uint32_t*>(out); for (int k = 0; k < 24; k++) o[k] = s.w[k]; return; }
    if (r.inf) { for (int k = 0; k < 64; k++) out[k] = 0; return; }
    ge a; ge_set_gej(a, r);
    fe_get_b32(out, a.x); fe_get_b32(out + 32, a.y);
}
// Every chain in here is run by a whole warp on the lane form of ge_w10.cuh: one warp per window for the window sums (16 of them for
// 16-bit windows: 512 threads = one round), one warp per GLV half for the Horner chains, warp 0 for the join and the inversion.
#define MSM_TAIL_THREADS 512
__global__ void __launch_bounds__(MSM_TAIL_THREADS) k_msm_tail(uint8_t* out, int mode, msm_ws ws, msm_plan p) {
    __shared__ gej W[2 * 32];
    __shared__ gej H[2];
    const uint32_t warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (uint32_t wi = warp; wi < 2 * p.nw; wi += nwarps) {
        msm_win w;
        msm_window_by_index(w, p, wi);
        gej r;
        if (msm_window_owned(p, wi)) msm_window_from_sums(r, ws.T + (size_t)wi * MSM_MAX_SUMS * MSM_R2_PARTS, MSM_R2_PARTS, msm_ilog2(w.Bw / MSM_CHUNK));
        else gej_set_infinity(r);                                       // warp-uniform
        if ((threadIdx.x & 31) == 0) W[wi] = r;
    }
    __syncthreads();
    if (warp < 2) {                                                     // warp 0 / warp 1: the two halves, side by side
        gej r;
        msm_horner_half(r, W + warp * p.nw, p);
        if ((threadIdx.x & 31) == 0) H[warp] = r;
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;                                      // the rest is one warp's work
    gej r, a = H[0], b = H[1];
    msm_join_halves(r, a, b);
    if (threadIdx.x == 0) write_result(out, r, mode);
}

// sum of `count` stored Jacobian points (the per-GPU partial results of a sharded MSM) -> affine x||y big-endian
__global__ void __launch_bounds__(128) k_gej_sum(uint8_t* out64, const uint8_t* partials96, uint32_t count) {
    __shared__ gej S[32];
    const uint32_t g = threadIdx.x >> 2, l = threadIdx.x & 3, ng = blockDim.x >> 2;
    gej acc;
    gej_set_infinity(acc);
    for (uint32_t k0 = 0; k0 < count; k0 += ng) {                       // uniform trip count; surplus groups add infinity
        const uint32_t k = k0 + g;
        gej o, t;
        gej_set_infinity(o);
        if (k < count) {
            gej_store s;
            const uint32_t* q = reinterpret_cast<const uint32_t*>(partials96 + (size_t)96 * k);
            for (int j = 0; j < 24; j++) s.w[j] = q[j];
            msm_load_gej(o, s);
        }
        gej_add_g4(t, acc, o); acc = t;
    }
    if (l == 0) S[g] = acc;
    __syncthreads();
    for (uint32_t d = ng >> 1; d >= 1; d >>= 1) {
        gej a = S[g], b, t;
        if (g < d) b = S[g + d]; else gej_set_infinity(b);
        gej_add_g4(t, a, b);
        __syncthreads();
        if (g < d && l == 0) S[g] = t;
        __syncthreads();
    }
    if (threadIdx.x == 0) write_result(out64, S[0], 0);
}

static int sum_grid_blocks() {
    static int cached = 0;
    if (cached) return cached;
    int dev = 0, sms = 148, per = MSM_SUM_MINBLOCKS;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_msm_bucket_sum, 128, 0) != cudaSuccess || per < 1) per = MSM_SUM_MINBLOCKS;
    cached = sms * per;
    return cached;
}

cudaError_t launch_msm(uint8_t* out, int out_mode, const uint8_t* g_scalar32, const uint8_t* scalars32, const uint8_t* points64, size_t n,
                       uint8_t* workspace, size_t workspace_bytes, int window_bits, cudaStream_t s, uint64_t* launches, uint32_t shard, uint32_t shards) {
    if (n + 1 >= ((size_t)1 << 31)) return cudaErrorInvalidValue;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return cudaErrorInvalidValue;
    if (shards == 0 || shard >= shards) return cudaErrorInvalidValue;
    const size_t N = n + 1;
    int c = window_bits ? window_bits : (int)msm_choose_c(N);
    if (!c_allowed(N, c)) return cudaErrorInvalidValue;
    msm_plan p;
    msm_make_plan_c(p, N, (uint32_t)c, shard, shards);
    msm_ws ws;
    if (carve(&ws, workspace, p) > workspace_bytes) return cudaErrorInvalidValue;
    const msm_caps caps = caps_of(p);
    cudaError_t err = cudaMemsetAsync(ws.hist, 0, (size_t)((uint8_t*)ws.counters - (uint8_t*)ws.hist) + 16 + (size_t)2 * p.nw * MSM_MAX_SUMS * 4, s);
    if (err != cudaSuccess) return err;
    err = cudaMemsetAsync(ws.buckets, 0, (size_t)p.TB * sizeof(gej_store), s);
    if (err != cudaSuccess) return err;
    const unsigned tgrid = (unsigned)((N + 255) / 256);
    k_msm_prepare<<<tgrid, 256, 0, s>>>(ws, p, scalars32, points64, g_scalar32, n);
    k_msm_scan_a<<<(unsigned)caps.nblk, 1024, 0, s>>>(ws, p);
    k_msm_scan_b<<<(unsigned)caps.nblk, 1024, 0, s>>>(ws, p);
    k_msm_scatter<<<tgrid, 256, 0, s>>>(ws, p);
    k_msm_units<<<(unsigned)((caps.cap_units + 255) / 256), 256, 0, s>>>(ws, p);
    size_t sum_blocks = (caps.cap_units + 127) / 128;
    if (sum_blocks > (size_t)sum_grid_blocks()) sum_blocks = (size_t)sum_grid_blocks();
    k_msm_bucket_sum<<<(unsigned)sum_blocks, 128, 0, s>>>(ws, p, points64, n);
    size_t comb_blocks = caps.cap_multi < 1184 ? caps.cap_multi : 1184;
    k_msm_combine<<<(unsigned)comb_blocks, 256, 0, s>>>(ws, p);
    size_t combs_blocks = (caps.cap_multi + 127) / 128;
    if (combs_blocks > 2368) combs_blocks = 2368;
    k_msm_combine_small<<<(unsigned)combs_blocks, 128, 0, s>>>(ws, p);
    const size_t chunks = p.TB / MSM_CHUNK;
    if (chunks <= 40960) k_msm_r1<true><<<(unsigned)((chunks * 4 + 127) / 128), 128, 0, s>>>(ws, p);
    else k_msm_r1<false><<<(unsigned)((chunks + 127) / 128), 128, 0, s>>>(ws, p);
    k_msm_r2<<<2 * p.nw * MSM_MAX_SUMS * MSM_R2_PARTS, 256, 0, s>>>(ws, p);
    k_msm_tail<<<1, MSM_TAIL_THREADS, 0, s>>>(out, out_mode, ws, p);
    if (launches) __atomic_fetch_add(launches, (uint64_t)11, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

cudaError_t launch_gej_sum(uint8_t* out64, const uint8_t* partials96, size_t count, cudaStream_t s, uint64_t* launches) {
    if (count >= ((size_t)1 << 31)) return cudaErrorInvalidValue;
    k_gej_sum<<<1, 128, 0, s>>>(out64, partials96, (uint32_t)count);
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

}  // namespace b200
