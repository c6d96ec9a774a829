// msm.cuh -- building blocks of the GPU Pippenger multi-scalar multiplication  R = sum_i s_i * P_i.
//
// Computes what secp256k1_ecmult_multi_var (ecmult_impl.h:823-867) -> secp256k1_ecmult_pippenger_batch (:670-731)
// -> secp256k1_ecmult_pippenger_wnaf (:516-591) compute, re-designed for a GPU:
//   reference: per window, clear buckets, add every point into its bucket one after the other (:545-566), then the
//              running-sum reduction (:581-588); one thread, scratch space sized by the bucket window (:597-621).
//   here:      c = 16-bit signed windows (16 windows, 2^15 buckets each; a scalar above n/2 is replaced by its negative and the
//              digit signs flipped, so the recoding never carries out of the top window).  (1) every term emits its 16 signed digits and
//              bumps a (window,bucket) histogram; (2) exclusive scan; (3) counting-sort scatter of term indices, so each
//              bucket's points are contiguous; (4) ONE THREAD PER BUCKET sums its points with mixed adds (557 056
//              independent threads at 2^20 terms, ~32 points each); (5) the running-sum reduction is cut into chunks of
//              32 buckets: chunk value = L + b0*A with L = sum j*S_{b0+j}, A = sum S_{b0+j} (b0*A by a 15-bit
//              double-and-add); (6) tree-sum of the chunks per window, Horner over the 16 windows.
// No GLV split (it halves the windows but doubles the points: same number of bucket additions on a GPU).
// Only the final point has to match the reference; it does because the group element is unique.
#pragma once
#include "ge.cuh"
#include "sc.cuh"

namespace B200_NS {

#define MSM_C 16
#define MSM_WINDOWS 16
#define MSM_BUCKETS 32768            // per window: |digit| in 1..32768
#define MSM_CHUNK 32                 // buckets per reduction chunk
#define MSM_CHUNKS (MSM_BUCKETS / MSM_CHUNK)

// signed 16-bit window recoding of a scalar: k = sum d_w 2^(16w) (mod n), d_w in [-32768, 32768], w = 0..15.
// k > n/2 is recoded as -(n - k): then the top window holds < 2^15 and nothing carries out of it.  (Without this every scalar with
// bit 255 set put a digit 1 into a 17th window and ONE bucket thread summed half of all points serially: 256 ms at 131 072 terms.)
HD void msm_digits(int32_t* d /*[16]*/, const sc& k0) {
    int neg = sc_is_high(k0);
    sc k;
    if (neg) sc_negate(k, k0); else k = k0;
    uint32_t carry = 0;
#pragma unroll
    for (int w = 0; w < 16; w++) {
        uint32_t v = ((k.v[w >> 1] >> ((w & 1) * 16)) & 0xFFFFu) + carry;
        int32_t dw;
        if (v > 32768u) { dw = (int32_t)v - 65536; carry = 1; }
        else { dw = (int32_t)v; carry = 0; }
        d[w] = neg ? -dw : dw;
    }
}

// affine point storage used by the MSM kernels: 16 words x[8] y[8] (LE limbs), all-zero = infinity
HD void msm_load_point(ge& p, const uint32_t* s16) {
    uint32_t z = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { p.x.v[k] = s16[k]; p.y.v[k] = s16[8 + k]; z |= s16[k] | s16[8 + k]; }
    p.inf = (z == 0);
}

// Jacobian point storage: 25 words (x, y, z, inf)
struct gej_store { uint32_t w[25]; };
HD void msm_store_gej(gej_store& s, const gej& a) {
#pragma unroll
    for (int k = 0; k < 8; k++) { s.w[k] = a.x.v[k]; s.w[8 + k] = a.y.v[k]; s.w[16 + k] = a.z.v[k]; }
    s.w[24] = (uint32_t)a.inf;
}
HD void msm_load_gej(gej& a, const gej_store& s) {
#pragma unroll
    for (int k = 0; k < 8; k++) { a.x.v[k] = s.w[k]; a.y.v[k] = s.w[8 + k]; a.z.v[k] = s.w[16 + k]; }
    a.inf = (int)s.w[24];
}

// (4) bucket sum: acc = sum over sorted entries [begin,end) of +-P[idx]; entry = idx | sign<<31
HD void msm_bucket_sum(gej& acc, const uint32_t* entries, uint32_t begin, uint32_t end, const uint32_t* points16) {
    gej_set_infinity(acc);
#pragma unroll 1
    for (uint32_t e = begin; e < end; e++) {
        uint32_t v = entries[e];
        ge p;
        msm_load_point(p, points16 + (size_t)(v & 0x7FFFFFFFu) * 16);
        if (p.inf) continue;
        if (v >> 31) fe_neg(p.y, p.y);
        gej t;
        gej_add_ge(t, acc, p);
        acc = t;
    }
}

// (5) chunk reduction over buckets b0+1 .. b0+MSM_CHUNK (bucket index b holds multiplier b+1):
//     out = sum_{j=0}^{CHUNK-1} (b0 + j + 1) * S[b0 + j]
HD void msm_chunk_reduce(gej& out, const gej_store* bucket_sums /* window base */, uint32_t b0) {
    gej running, acc;
    gej_set_infinity(running);
    gej_set_infinity(acc);
#pragma unroll 1
    for (int j = MSM_CHUNK - 1; j >= 0; j--) {
        gej s, t;
        msm_load_gej(s, bucket_sums[b0 + j]);
        gej_add(t, running, s); running = t;       // running = sum_{i>=j} S_i
        gej_add(t, acc, running); acc = t;         // acc = sum_j (j+1) S_j   (local multipliers 1..CHUNK)
    }
    // + b0 * A, A = running (plain sum of the chunk)
    if (b0 != 0 && !running.inf) {
        gej m;
        gej_set_infinity(m);
#pragma unroll 1
        for (int bit = 14; bit >= 0; bit--) {
            gej t;
            if (!m.inf) { gej_double(t, m); m = t; }
            if ((b0 >> bit) & 1u) { gej_add(t, m, running); m = t; }
        }
        gej t;
        gej_add(t, acc, m); acc = t;
    }
    out = acc;
}

// sum of `count` stored Jacobian points
HD void msm_sum_range(gej& out, const gej_store* v, uint32_t count) {
    gej acc;
    gej_set_infinity(acc);
#pragma unroll 1
    for (uint32_t i = 0; i < count; i++) {
        gej s, t;
        msm_load_gej(s, v[i]);
        gej_add(t, acc, s); acc = t;
    }
    out = acc;
}

// (6) Horner over windows: R = sum_w 2^(16 w) * W[w]
HD void msm_horner(gej& out, const gej_store* window_sums) {
    gej acc;
    gej_set_infinity(acc);
#pragma unroll 1
    for (int w = MSM_WINDOWS - 1; w >= 0; w--) {
        if (!acc.inf) {
#pragma unroll 1
            for (int k = 0; k < MSM_C; k++) { gej t; gej_double(t, acc); acc = t; }
        }
        gej s, t;
        msm_load_gej(s, window_sums[w]);
        gej_add(t, acc, s); acc = t;
    }
    out = acc;
}

}  // namespace B200_NS
