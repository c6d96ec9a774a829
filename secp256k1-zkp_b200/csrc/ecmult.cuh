// ecmult.cuh -- R = na*A + ng*G for one item per GPU thread.
//
// Computes what secp256k1_ecmult (ecmult_impl.h:365) / secp256k1_ecmult_strauss_wnaf (:252-363)
// compute, re-designed for SIMT:
//  * wNAF is pointless on a GPU (32 lanes have 32 different sparse digit patterns, so the warp
//    executes the add at every position anyway).  Instead both GLV halves (scalar_impl.h:142) use a
//    FIXED 4-bit window with signed ODD digits: every lane adds at the same 33 positions, no
//    zero digits, no divergence.  The 8-entry table of odd multiples is the one ecmult_impl.h:73-115
//    builds (same isomorphism / global-z trick, group_impl.h:289-320) so all adds are mixed adds.
//  * the table lives in per-thread LOCAL memory (tab_local: the hardware interleaves it [word][thread] across the warp, so any per-lane
//    entry index is a coalesced access) since r1h; before that, and still in k_bppp.cu, in shared memory laid out [entry][word][thread]
//    (tab_shared: bank = thread, conflict free).  Never a dynamically indexed REGISTER array.
//  * ng*G needs no doublings at all: a comb table T[j][d] = d * 2^(gw*j) * G in HBM/L2 (gw = 16:
//    16 windows x 65536 affine points = 64 MiB, built on the GPU at engine init) replaces the
//    reference's 2 x 8192-entry wNAF tables (precomputed_ecmult.c); 16 mixed adds per item.
// Any correct algorithm yields the same group element; parity is checked on serialized outputs.
#pragma once
#include "ge.cuh"
#include "sc.cuh"

namespace B200_NS {

// affine point as stored in HBM: x[8], y[8] LE limbs == secp256k1_ge_storage (group.h:36-41)
struct ge_storage { uint32_t x[8]; uint32_t y[8]; };

struct ecmult_gtable {
    const uint32_t* data;   // [windows][1 << gw][16] words; entry d=0 unused
    int gw;                 // window bits
    int windows;            // ceil(256 / gw)
};

HD size_t ecmult_gtable_words(int gw) {
    int windows = (256 + gw - 1) / gw;
    return (size_t)windows * ((size_t)1 << gw) * 16;
}

// beta: cube root of unity in F_p with lambda*(x,y) = (beta*x, y)   (field.h:69-72)
HD void fe_set_beta(fe& b) {
    const uint32_t v[8] = {0x719501eeu, 0xc1396c28u, 0x12f58995u, 0x9cf04975u, 0xac3434e9u, 0x6e64479eu, 0x657c0710u, 0x7ae96a2bu};
#pragma unroll
    for (int i = 0; i < 8; i++) b.v[i] = v[i];
}

// per-thread table storage: plain local array (host emulation, generic kernels)
struct tab_local {
    uint32_t w[8][16];
    HD void put(int i, const fe& x, const fe& y) {
#pragma unroll
        for (int k = 0; k < 8; k++) { w[i][k] = x.v[k]; w[i][8 + k] = y.v[k]; }
    }
    HD void get(int i, fe& x, fe& y) const {
#pragma unroll
        for (int k = 0; k < 8; k++) { x.v[k] = w[i][k]; y.v[k] = w[i][8 + k]; }
    }
};
// per-thread table storage in shared memory: word (i,k) of thread t at base[(i*16+k)*stride + t]
struct tab_shared {
    uint32_t* base;   // smem + threadIdx.x
    int stride;       // blockDim.x
    HD void put(int i, const fe& x, const fe& y) {
#pragma unroll
        for (int k = 0; k < 8; k++) { base[(i * 16 + k) * stride] = x.v[k]; base[(i * 16 + 8 + k) * stride] = y.v[k]; }
    }
    HD void get(int i, fe& x, fe& y) const {
#pragma unroll
        for (int k = 0; k < 8; k++) { x.v[k] = base[(i * 16 + k) * stride]; y.v[k] = base[(i * 16 + 8 + k) * stride]; }
    }
};

// signed-odd-digit recoding of a 128-bit magnitude.  k odd, < 2^128.
//   k = sum_{i=0}^{32} (2*t_i - 15) * 16^i   with  t_i = 4-bit window i of v = (k + 2^132 - 1) / 2
struct glv_digits { uint32_t v[5]; };   // 132 bits
HD void glv_recode(glv_digits& d, const uint32_t* k4 /*4 limbs, odd*/) {
    // v = (k + 2^132 - 1) >> 1
    uint64_t acc = 0;
    uint32_t s[5];
    const uint32_t add[5] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xfu};
#pragma unroll
    for (int i = 0; i < 5; i++) {
        acc += (uint64_t)(i < 4 ? k4[i] : 0u) + add[i];
        s[i] = (uint32_t)acc;
        acc >>= 32;
    }
    // k + 2^132 - 1 < 2^133 : fits 5 limbs (bit 132 may be set in s[4] bit 4)
#pragma unroll
    for (int i = 0; i < 4; i++) d.v[i] = (s[i] >> 1) | (s[i + 1] << 31);
    d.v[4] = s[4] >> 1;
}
HD uint32_t glv_digit(const glv_digits& d, int i) {   // i in [0,32]
    return (d.v[i >> 3] >> ((i & 7) * 4)) & 0xFu;
}

// Build the table of odd multiples 1,3,..,15 of `a` as affine points on the curve isomorphic by the
// returned global Z (ecmult_impl.h:73-115 + group_impl.h:289-320).  a must not be infinity.
template <class Tab>
HD void ecmult_odd_multiples(Tab& tab, fe& globalz, const gej& a) {
    gej d;
    gej_double(d, a);
    fe dz2, dz3;
    fe_sqr(dz2, d.z);
    fe_mul(dz3, dz2, d.z);
    ge daff; daff.x = d.x; daff.y = d.y; daff.inf = 0;   // d seen from the curve scaled by d.z
    gej p;
    fe_mul(p.x, a.x, dz2);
    fe_mul(p.y, a.y, dz3);
    p.z = a.z;
    p.inf = 0;
    tab.put(0, p.x, p.y);
    fe zr[8];
#pragma unroll 1
    for (int i = 1; i < 8; i++) {
        gej q;
        gej_add_ge(q, p, daff, &zr[i]);
        p = q;
        tab.put(i, p.x, p.y);
    }
    // bring all entries to the z of the last one
    fe_mul(globalz, p.z, d.z);
    fe zs = zr[7];
#pragma unroll 1
    for (int i = 6; i >= 0; i--) {
        fe x, y, zs2, zs3;
        tab.get(i, x, y);
        fe_sqr(zs2, zs);
        fe_mul(zs3, zs2, zs);
        fe_mul(x, x, zs2);
        fe_mul(y, y, zs3);
        tab.put(i, x, y);
        if (i > 0) fe_mul(zs, zs, zr[i]);
    }
}

HD void ecmult_load_gentry(ge& p, const ecmult_gtable& gt, int j, uint32_t d) {
    const uint32_t* e = gt.data + (((size_t)j << gt.gw) + d) * 16;
#if defined(__CUDA_ARCH__)
    const uint4* e4 = reinterpret_cast<const uint4*>(e);
    uint4 q0 = __ldg(e4), q1 = __ldg(e4 + 1), q2 = __ldg(e4 + 2), q3 = __ldg(e4 + 3);
    p.x.v[0] = q0.x; p.x.v[1] = q0.y; p.x.v[2] = q0.z; p.x.v[3] = q0.w;
    p.x.v[4] = q1.x; p.x.v[5] = q1.y; p.x.v[6] = q1.z; p.x.v[7] = q1.w;
    p.y.v[0] = q2.x; p.y.v[1] = q2.y; p.y.v[2] = q2.z; p.y.v[3] = q2.w;
    p.y.v[4] = q3.x; p.y.v[5] = q3.y; p.y.v[6] = q3.z; p.y.v[7] = q3.w;
#else
    for (int k = 0; k < 8; k++) { p.x.v[k] = e[k]; p.y.v[k] = e[8 + k]; }
#endif
    p.inf = 0;
}

// window j (gw bits) of a 256-bit scalar
HD uint32_t sc_window(const sc& s, int pos, int bits) {
    int limb = pos >> 5, off = pos & 31;
    uint64_t w = s.v[limb];
    if (limb + 1 < 8) w |= (uint64_t)s.v[limb + 1] << 32;
    return (uint32_t)(w >> off) & ((1u << bits) - 1u);
}

// r += ng*G using the comb table
HD void ecmult_gen_accumulate(gej& r, const sc& ng, const ecmult_gtable& gt) {
#pragma unroll 1
    for (int j = 0; j < gt.windows; j++) {
        int pos = j * gt.gw;
        int bits = (256 - pos) < gt.gw ? (256 - pos) : gt.gw;
        uint32_t d = sc_window(ng, pos, bits);
        if (d) {
            ge e;
            ecmult_load_gentry(e, gt, j, d);
            gej t;
            gej_add_ge(t, r, e);
            r = t;
        }
    }
}

// r = na*a + ng*G.   a may be infinity; na / ng may be zero.
template <class Tab>
HD void ecmult(gej& r, const gej& a, const sc& na, const sc& ng, const ecmult_gtable& gt, Tab& tab) {
    gej_set_infinity(r);
    if (!a.inf && !sc_is_zero(na)) {
        sc k1, k2;
        sc_split_lambda(k1, k2, na);
        int neg1 = sc_is_high(k1), neg2 = sc_is_high(k2);
        sc_cond_negate(k1, neg1);
        sc_cond_negate(k2, neg2);
        // make both odd; remember the +1 ("skew") to take it out at the end
        uint32_t skew1 = (k1.v[0] & 1u) ^ 1u, skew2 = (k2.v[0] & 1u) ^ 1u;
        k1.v[0] |= 1u; k2.v[0] |= 1u;     // even -> +1 (no carry since the low bit was 0)
        glv_digits d1, d2;
        glv_recode(d1, k1.v);
        glv_recode(d2, k2.v);
        fe globalz, beta;
        ecmult_odd_multiples(tab, globalz, a);
        fe_set_beta(beta);
#pragma unroll 1
        for (int i = 32; i >= 0; i--) {
            if (!r.inf) {
#pragma unroll 1
                for (int k = 0; k < 4; k++) { gej t; gej_double_hot(t, r); r = t; }
            }
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
                uint32_t t = half ? glv_digit(d2, i) : glv_digit(d1, i);
                int neg = (int)((t >> 3) ^ 1u) ^ (half ? neg2 : neg1);
                int idx = (int)((t & 8u) ? (t & 7u) : ((t ^ 7u) & 7u));
                ge e;
                tab.get(idx, e.x, e.y);
                e.inf = 0;
                if (half) fe_mul(e.x, e.x, beta);
                if (neg) fe_neg(e.y, e.y);
                gej q;
                gej_add_ge_hot(q, r, e);
                r = q;
            }
        }
        // remove the skews: subtract (sign)*a resp. (sign)*lambda*a
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            uint32_t sk = half ? skew2 : skew1;
            if (sk) {
                ge e;
                tab.get(0, e.x, e.y);
                e.inf = 0;
                if (half) fe_mul(e.x, e.x, beta);
                int neg = (half ? neg2 : neg1) ^ 1;
                if (neg) fe_neg(e.y, e.y);
                gej q;
                gej_add_ge(q, r, e);
                r = q;
            }
        }
        if (!r.inf) fe_mul(r.z, r.z, globalz);
    }
    if (!sc_is_zero(ng)) ecmult_gen_accumulate(r, ng, gt);
}

}  // namespace B200_NS
