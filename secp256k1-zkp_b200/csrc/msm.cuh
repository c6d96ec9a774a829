// msm.cuh -- building blocks of the GPU Pippenger multi-scalar multiplication  R = g*G + sum_i s_i * P_i.
//
// Computes what secp256k1_ecmult_multi_var (ecmult_impl.h:823-867) -> secp256k1_ecmult_pippenger_batch (:670-731)
// -> secp256k1_ecmult_pippenger_wnaf (:516-591) compute, re-designed for a GPU (round 2):
//
//   reference                                              here
//   ---------------------------------------------------   ---------------------------------------------------------------------
//   dispatcher: Strauss below 88 points, else a bucket     msm_make_plan(): window width c in [4,16] chosen from n by a work model
//   window from a table (:597-621, :823-867)               (the same trade the reference's table encodes: n*windows + buckets)
//   ecmult_endo_split (:645-658): 2n points, the second    GLV split of the SCALAR only: windows [0,nw) hold the digits of |k1|,
//   half multiplied by lambda (one fe_mul by beta each)    windows [nw,2nw) those of |k2|, both over the SAME point P_i; because the
//                                                          endomorphism is a group homomorphism, sum k2_i*lambda(P_i) =
//                                                          lambda(sum k2_i*P_i): ONE multiplication by beta at the very end
//   wnaf_fixed digits + skew (:437-497)                    signed c-bit digits in [-2^(c-1), 2^(c-1)], the top window of each half
//                                                          unsigned (|k| < 2^128 never carries out of it)
//   serial bucket fill, one window after the other         counting sort by (window, bucket); buckets cut into PIECES of <= Tp entries;
//   (:545-566)                                             one lane per piece, lanes of a warp pull pieces dynamically (no lane waits
//                                                          for a longer bucket); pieces of a hot bucket are combined by one block with a
//                                                          warp-shuffle tree -- 2^20 equal scalars cost the same bucket work as random ones
//   running-sum reduction (:581-588), serial               chunks of 8 buckets (running sum, all chunks in parallel), then per window
//                                                          sum_q L_q + 8*sum_q q*A_q with the second sum taken bit by bit (tree sums)
//   Horner over the windows (:536-589), serial              two Horner chains (one per GLV half) side by side, every doubling / addition
//                                                          spread over a whole warp (ge_w10.cuh: one 26-bit limb per lane), then R = A + lambda(B)
//   one MSM = one machine                                   shards by terms (partial sums, msm.cuh is unchanged) or by bucket windows (msm_window_owned)
// Only the final point has to match the reference; it does because the group element is unique.
#pragma once
#include "gtable.cuh"
#include "ge_g4.cuh"
#include "ge_w10.cuh"
#include "sc.cuh"

namespace B200_NS {

#define MSM_CHUNK 8                  // buckets per running-sum chunk
#define MSM_PIECE_MIN 16             // smallest piece size
#define MSM_PIECE_MAX 4096
#define MSM_SMALL_MULTI 8            // multi-piece buckets with at most this many pieces are combined by one thread, larger ones by a block
#define MSM_R2_ELEMS 2048            // chunk sums per block of the plain / bit sums
#define MSM_R2_PARTS 4               // => at most 4 blocks per sum (the largest window has 8192 chunks)
#define MSM_C_MIN 4
#define MSM_C_MAX 16
#define MSM_MAX_SUMS 15              // per window: 1 plain sum (L) + up to 14 bit sums
#define MSM_SCAN_ELEMS 4096          // buckets per block of the two-level scan

struct msm_plan {
    uint32_t c;        // window bits
    uint32_t nw;       // windows per GLV half = ceil(128 / c)
    uint32_t B;        // buckets of a regular window = 2^(c-1)   (digit magnitudes 1..B)
    uint32_t Btop;     // buckets of the top window of a half = max(2^t, 8), t = 128 - c*(nw-1) (unsigned digit 1..2^t)
    uint32_t HB;       // buckets per half = (nw-1)*B + Btop
    uint32_t TB;       // total buckets = 2*HB (a multiple of 8)
    uint32_t Tp;       // entries per piece
    uint32_t N;        // terms incl. the generator slot (index N-1)
    uint32_t shards;   // window sharding: this launch sums the windows wi (0 <= wi < 2 nw) with wi % shards == shard only;
    uint32_t shard;    // the partial results of all shards add up to the MSM (shards = 1: everything)
};

HD uint32_t msm_ilog2(uint32_t x) { uint32_t r = 0; while (x > 1) { x >>= 1; r++; } return r; }
// Window sharding: shard s sums window j of the first GLV half when j % shards == s, and window nw - 1 - j of the second: the top windows
// (twice the buckets, the longest Horner chain) of the two halves go to different shards.
HD bool msm_window_owned(const msm_plan& p, uint32_t wi) {
    if (p.shards <= 1) return true;
    const uint32_t j = wi < p.nw ? wi : 2 * p.nw - 1 - wi;
    return j % p.shards == p.shard;
}
HD uint32_t msm_windows_owned(const msm_plan& p) {                       // how many of the 2 nw windows this launch sums (needs p.nw)
    return p.shard < p.nw ? 2 * ((p.nw - p.shard + p.shards - 1) / p.shards) : 0;
}

HD void msm_make_plan_c(msm_plan& p, size_t N, uint32_t c, uint32_t shard = 0, uint32_t shards = 1) {
    p.c = c;
    p.shards = shards ? shards : 1;
    p.shard = shard;
    p.nw = (128 + c - 1) / c;
    uint32_t t = 128 - c * (p.nw - 1);
    p.B = 1u << (c - 1);
    p.Btop = (1u << t) < 8u ? 8u : (1u << t);
    p.HB = (p.nw - 1) * p.B + p.Btop;
    p.TB = 2 * p.HB;
    p.N = (uint32_t)N;
    // piece size: small enough that the pieces outnumber the GPU's lanes several times over (~2^19 pieces when there is that much
    // work) and that no lane walks a long bucket alone -- the top window of a half holds only ~0.6 * 2^t distinct digits, so its
    // buckets are several times fuller than the others -- but not smaller than 16 entries (every extra piece costs a Jacobian addition).
#ifndef MSM_PIECE_SHIFT
#define MSM_PIECE_SHIFT 19
#endif
    const uint64_t own = msm_windows_owned(p);
    uint64_t e = ((own ? own : 1) * N) >> MSM_PIECE_SHIFT;
    uint32_t tp = MSM_PIECE_MIN;
    while (tp * 2 <= e && tp < MSM_PIECE_MAX) tp <<= 1;
    p.Tp = tp;
}
// Window width from the number of terms -- the counterpart of secp256k1_pippenger_bucket_window's table (ecmult_impl.h:597-621).
// On the GPU wide windows win much earlier than on a CPU: more buckets = more parallel lanes, and the bucket reduction is parallel
// too.  Measured (tools/msm_bench.py --sweep, profiles/r2*_msm_sweep.jsonl); below 2^12 terms the MSM is latency bound (a ~128-doubling
// chain whatever c is) and a narrow window keeps the reduction short.
//   terms      8     88    1024  4096  16384 32768 65536 131072 262144 2^20      (ms per MSM, profiles/r2d_msm_sweep.jsonl)
//   c = 8    0.50  0.53   0.58  0.60  0.96  1.49  1.83  2.43   3.49   7.35
//   c = 10   0.50  0.54   0.59  0.60  0.72  0.79  1.91  3.41   4.52   8.31
//   c = 11   0.51  0.56   0.63  0.65  0.74  0.79  0.97  2.96   5.89  10.02
//   c = 12   0.51  0.62   0.70  0.70  0.82  0.87  1.06  1.29   4.96  11.90
//   c = 16   0.59  0.72   0.89  0.99  1.03  1.07  1.16  1.34   1.71   3.85
HD uint32_t msm_choose_c(size_t N) {
    if (N < ((size_t)1 << 13)) return 8;
    if (N < ((size_t)1 << 15)) return 10;
    if (N < ((size_t)1 << 17)) return 11;
    if (N < ((size_t)1 << 18)) return 12;
    return 16;
}

// bits [pos, pos + nb) of a 128-bit magnitude (nb <= 16, pos + nb <= 128)
HD uint32_t msm_bits128(const uint32_t* m4, uint32_t pos, uint32_t nb) {
    uint32_t limb = pos >> 5, off = pos & 31;
    uint32_t x = m4[limb] >> off;
    if (off + nb > 32 && limb + 1 < 4) x |= m4[limb + 1] << (32 - off);
    return x & ((nb >= 32) ? 0xFFFFFFFFu : ((1u << nb) - 1u));
}

// signed digits of one GLV half: emit(bucket, sign) for every non-zero digit.  bucket = half*HB + w*B + (|digit| - 1).
template <class Emit>
HD void msm_emit_digits(const msm_plan& p, const uint32_t* mag4, uint32_t neg, uint32_t half, Emit& emit) {
    uint32_t carry = 0;
    const uint32_t base = half * p.HB;
#pragma unroll 1
    for (uint32_t w = 0; w < p.nw; w++) {
        uint32_t pos = w * p.c;
        uint32_t nb = (w + 1 < p.nw) ? p.c : (128 - pos);
        uint32_t v = msm_bits128(mag4, pos, nb) + carry;
        uint32_t sgn = neg;
        if (w + 1 < p.nw && v > p.B) { v = (1u << p.c) - v; sgn ^= 1u; carry = 1; } else carry = 0;
        if (v && msm_window_owned(p, half * p.nw + w)) emit(base + w * p.B + v - 1, sgn);
    }
}

// per-term preparation: GLV split (scalar_impl.h:142-180), magnitudes and signs.  flags: bit0 = k1 negative, bit1 = k2 negative,
// bit2 = term contributes nothing (zero scalar or point at infinity).
HD void msm_prepare_term(uint32_t* k12 /*[8]*/, uint32_t& flags, const sc& k, int point_inf) {
    if (point_inf || sc_is_zero(k)) {
#pragma unroll
        for (int t = 0; t < 8; t++) k12[t] = 0;
        flags = 4;
        return;
    }
    sc k1, k2;
    sc_split_lambda(k1, k2, k);
    uint32_t n1 = (uint32_t)sc_is_high(k1), n2 = (uint32_t)sc_is_high(k2);
    sc_cond_negate(k1, (int)n1);
    sc_cond_negate(k2, (int)n2);
#pragma unroll
    for (int t = 0; t < 4; t++) { k12[t] = k1.v[t]; k12[4 + t] = k2.v[t]; }
    flags = n1 | (n2 << 1);
}

// Jacobian point storage: 24 words x, y, z (LE limbs); z == 0 (all words) <=> infinity, so zeroed memory is a valid "empty".
struct gej_store { uint32_t w[24]; };
HD void msm_store_gej(gej_store& s, const gej& a) {
    if (a.inf) {
#pragma unroll
        for (int k = 0; k < 24; k++) s.w[k] = 0;
        return;
    }
    fe z = a.z;
    fe_normalize(z);                 // a weakly reduced z could be p itself only if it were 0 mod p, which a finite point never has;
#pragma unroll                       // normalising keeps "all words zero" an exact test anyway
    for (int k = 0; k < 8; k++) { s.w[k] = a.x.v[k]; s.w[8 + k] = a.y.v[k]; s.w[16 + k] = z.v[k]; }
}
HD void msm_load_gej(gej& a, const gej_store& s) {
    uint32_t zz = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { a.x.v[k] = s.w[k]; a.y.v[k] = s.w[8 + k]; a.z.v[k] = s.w[16 + k]; zz |= s.w[16 + k]; }
    a.inf = (zz == 0);
}

// a unit of bucket work: <= Tp consecutive sorted entries of one bucket
struct msm_unit { uint32_t begin, count, dst, first; };    // dst: bucket index, or 0x80000000 | index into the parts array; first = entries[begin]
                                                           // (one dependent load less when a lane of k_msm_bucket_sum starts a piece)
#define MSM_DST_PART 0x80000000u

// affine point i of the input (x||y big-endian, raw 256-bit values); index n = the generator slot
HD void msm_load_point(ge& p, const uint8_t* points64, size_t n, const uint8_t* gpoint64, uint32_t idx) {
    const uint8_t* s = (idx < n) ? points64 + (size_t)64 * idx : gpoint64;
#if defined(__CUDA_ARCH__)
    if ((reinterpret_cast<uintptr_t>(s) & 15) == 0) {
        const uint4* q = reinterpret_cast<const uint4*>(s);
        uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
        // big-endian 32 bytes -> LE limbs: limb 7 = first word byte-swapped
        p.x.v[7] = __byte_perm(a.x, 0, 0x0123); p.x.v[6] = __byte_perm(a.y, 0, 0x0123); p.x.v[5] = __byte_perm(a.z, 0, 0x0123); p.x.v[4] = __byte_perm(a.w, 0, 0x0123);
        p.x.v[3] = __byte_perm(b.x, 0, 0x0123); p.x.v[2] = __byte_perm(b.y, 0, 0x0123); p.x.v[1] = __byte_perm(b.z, 0, 0x0123); p.x.v[0] = __byte_perm(b.w, 0, 0x0123);
        p.y.v[7] = __byte_perm(c.x, 0, 0x0123); p.y.v[6] = __byte_perm(c.y, 0, 0x0123); p.y.v[5] = __byte_perm(c.z, 0, 0x0123); p.y.v[4] = __byte_perm(c.w, 0, 0x0123);
        p.y.v[3] = __byte_perm(d.x, 0, 0x0123); p.y.v[2] = __byte_perm(d.y, 0, 0x0123); p.y.v[1] = __byte_perm(d.z, 0, 0x0123); p.y.v[0] = __byte_perm(d.w, 0, 0x0123);
        p.inf = 0;
        return;
    }
#endif
    bn_from_be32(p.x.v, s);
    bn_from_be32(p.y.v, s + 32);
    p.inf = 0;
}

// sum of one unit: acc = sum over its entries of +-P[idx]   (entry = idx | sign << 31; infinity points never become entries)
HD void msm_unit_sum(gej& acc, const msm_unit& u, const uint32_t* entries, const uint8_t* points64, size_t n, const uint8_t* gpoint64) {
    gej_set_infinity(acc);
#pragma unroll 1
    for (uint32_t e = u.begin; e < u.begin + u.count; e++) {
        uint32_t v = entries[e];
        ge p;
        msm_load_point(p, points64, n, gpoint64, v & 0x7FFFFFFFu);
        if (v >> 31) fe_neg(p.y, p.y);
        gej t;
        gej_add_ge(t, acc, p);
        acc = t;
    }
}

// window geometry from a bucket / chunk position
struct msm_win { uint32_t half, w, base, Bw; };             // base = first bucket of the window
HD void msm_window_of_bucket(msm_win& o, const msm_plan& p, uint32_t bucket) {
    o.half = bucket / p.HB;
    uint32_t r = bucket - o.half * p.HB;
    o.w = r / p.B;
    if (o.w > p.nw - 1) o.w = p.nw - 1;
    o.base = o.half * p.HB + o.w * p.B;
    o.Bw = (o.w == p.nw - 1) ? p.Btop : p.B;
}
HD void msm_window_by_index(msm_win& o, const msm_plan& p, uint32_t wi /* 0 .. 2*nw-1 */) {
    o.half = wi / p.nw;
    o.w = wi - o.half * p.nw;
    o.base = o.half * p.HB + o.w * p.B;
    o.Bw = (o.w == p.nw - 1) ? p.Btop : p.B;
}

// running-sum chunk (reference :581-588 on 8 buckets): L = sum_{j<8} (j+1)*S_j, A = sum_j S_j
HD void msm_chunk_r1(gej& L, gej& A, const gej_store* s8) {
    gej running, acc;
    gej_set_infinity(running);
    gej_set_infinity(acc);
#pragma unroll 1
    for (int j = MSM_CHUNK - 1; j >= 0; j--) {
        gej s, t;
        msm_load_gej(s, s8[j]);
        gej_add(t, running, s); running = t;
        gej_add(t, acc, running); acc = t;
    }
    L = acc;
    A = running;
}

// element k of bit-sum j over Q chunk sums: the k-th index q in [0, Q) whose bit j is set
HD uint32_t msm_bit_index(uint32_t k, uint32_t j) { return ((k >> j) << (j + 1)) | (1u << j) | (k & ((1u << j) - 1u)); }

// The three functions below are serial chains of doublings and additions.  On the device each is run by ONE WHOLE WARP, converged, with the
// running point in lane form (ge_w10.cuh: one 26-bit limb of every coordinate per lane); compiled for the host they are plain code.
// window sum from its plain sum T[0] = sum_q L_q and bit sums T[1 + j] = sum_{q: bit j} A_q:
//   W = T[0] + 8 * sum_j 2^j T[1 + j]
HD void msm_window_from_sums(gej& W, const gej_store* T /* [1 + bits] sums, `stride` slots apart */, uint32_t stride, uint32_t nbits) {
    const w10_lane L = w10_lane_init();
    gej_w10 acc;
    w10_set_infinity(acc);
#pragma unroll 1
    for (int j = (int)nbits - 1; j >= -1; j--) {                  // j = -1: the plain sum, after the multiplication by 8
        w10_double_n(acc, j == -1 ? 3u : 1u, L);
        gej s;
        msm_load_gej(s, T[(size_t)(j + 1) * stride]);
        gej_w10 q;
        w10_set_gej(q, s, L);
        w10_add(acc, q, L);
    }
    w10_get_gej(W, acc);
}

// Horner over the nw windows of one half: sum_w 2^(c w) W[w]
HD void msm_horner_half(gej& out, const gej* W, const msm_plan& p) {
    const w10_lane L = w10_lane_init();
    gej_w10 acc;
    w10_set_infinity(acc);
#pragma unroll 1
    for (int w = (int)p.nw - 1; w >= 0; w--) {
        w10_double_n(acc, p.c, L);
        gej_w10 q;
        w10_set_gej(q, W[w], L);
        w10_add(acc, q, L);
    }
    w10_get_gej(out, acc);
}

// R = A + lambda(B), lambda(X, Y, Z) = (beta X, Y, Z)   (group_impl.h ge_mul_lambda on Jacobian coordinates)
HD void msm_join_halves(gej& R, const gej& A, const gej& Bh) {
    gej lb = Bh;
    if (!lb.inf) { fe beta; fe_set_beta(beta); fe_mul(lb.x, lb.x, beta); }
    gej_add_w10(R, A, lb);
}

}  // namespace B200_NS
