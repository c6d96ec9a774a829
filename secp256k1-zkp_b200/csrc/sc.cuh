// sc.cuh -- scalars modulo the group order n (reference: scalar_4x64_impl.h, scalar_impl.h).
//
// Same 8x32 little-endian limb form as the reference's 4x64 `secp256k1_scalar`.  Always kept
// fully reduced (< n), like the reference.  Scalar work is a few percent of a verification, so
// this file is plain C on top of bn_mul_wide; only the field layer has hand-scheduled PTX.
#pragma once
#include "bn.cuh"

namespace B200_NS {

struct sc { uint32_t v[8]; };

// n and 2^256 - n
#define SC_N_LIMBS {0xD0364141u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}
#define SC_NC_LIMBS {0x2FC9BEBFu, 0x402DA173u, 0x50B75FC4u, 0x45512319u, 1u, 0u, 0u, 0u}
// (n-1)/2
#define SC_NH_LIMBS {0x681B20A0u, 0xDFE92F46u, 0x57A4501Du, 0x5D576E73u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x7FFFFFFFu}

HD void sc_set_zero(sc& r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
}
HD void sc_set_int(sc& r, uint32_t a) { sc_set_zero(r); r.v[0] = a; }
HD void sc_set_u64(sc& r, uint64_t a) { sc_set_zero(r); r.v[0] = (uint32_t)a; r.v[1] = (uint32_t)(a >> 32); }
HD int sc_is_zero(const sc& a) { return bn_is_zero8(a.v); }
HD int sc_eq(const sc& a, const sc& b) { return bn_eq8(a.v, b.v); }

// subtract n once if r >= n; returns whether it did (scalar_4x64_impl.h:72-102)
HD int sc_reduce_once(sc& r) {
    const uint32_t nc[8] = SC_NC_LIMBS;
    uint32_t t[8];
    uint32_t c = bn_add8(t, r.v, nc);   // carry <=> r + (2^256 - n) >= 2^256 <=> r >= n
    uint32_t m = 0u - c;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (t[i] & m) | (r.v[i] & ~m);
    return (int)c;
}
// big-endian bytes -> scalar, reduced mod n; returns overflow flag (scalar_4x64_impl.h:155)
HD int sc_set_b32(sc& r, const uint8_t* b) {
    bn_from_be32(r.v, b);
    return sc_reduce_once(r);
}
HD void sc_get_b32(uint8_t* b, const sc& a) { bn_to_be32(b, a.v); }

HD void sc_add(sc& r, const sc& a, const sc& b) {
    const uint32_t nc[8] = SC_NC_LIMBS;
    uint32_t t[8], u[8];
    uint32_t c = bn_add8(t, a.v, b.v);
    // a+b < 2n: subtract n if carry or t >= n
    uint32_t c2 = bn_add8(u, t, nc);
    uint32_t m = 0u - (c | c2);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (u[i] & m) | (t[i] & ~m);
}
HD void sc_negate(sc& r, const sc& a) {
    const uint32_t n[8] = SC_N_LIMBS;
    uint32_t t[8];
    bn_sub8(t, n, a.v);
    uint32_t m = 0u - (uint32_t)(!bn_is_zero8(a.v));
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i] & m;
}
HD void sc_cond_negate(sc& r, int flag) {
    sc t; sc_negate(t, r);
    if (flag) r = t;
}
// a > (n-1)/2  (scalar_4x64_impl.h:252)
HD int sc_is_high(const sc& a) {
    const uint32_t nh[8] = SC_NH_LIMBS;
    uint32_t t[8];
    return bn_sub8(t, nh, a.v) != 0;  // borrow <=> a > nh
}

// out[0..nout) = lo[0..nlo) + a[0..na) * b[0..nb)   (caller guarantees it fits)
template <int NOUT, int NLO, int NA, int NB>
HD void bn_muladd(uint32_t* out, const uint32_t* lo, const uint32_t* a, const uint32_t* b) {
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
        uint64_t l = carry & 0xffffffffu, h = carry >> 32;
        if (k < NLO) l += lo[k];
#pragma unroll
        for (int i = 0; i < NA; i++) {
            int j = k - i;
            if (j < 0 || j >= NB) continue;
            uint64_t p = (uint64_t)a[i] * b[j];
            l += p & 0xffffffffu;
            h += p >> 32;
        }
        out[k] = (uint32_t)l;
        carry = h + (l >> 32);
    }
}

// r = t[0..15] mod n (three folds by 2^256 = NC (mod n); scalar_4x64_impl.h:636-711 does the same in 64-bit limbs)
HD void sc_reduce512(sc& r, const uint32_t* t) {
    const uint32_t nc[8] = SC_NC_LIMBS;
    uint32_t m1[13], m2[9], m3[9];
    bn_muladd<13, 8, 8, 5>(m1, t, t + 8, nc);        // < 2^386
    bn_muladd<9, 8, 5, 5>(m2, m1, m1 + 8, nc);       // < 2^260
    bn_muladd<9, 8, 1, 5>(m3, m2, m2 + 8, nc);       // < 2^256 + 2^134
    uint32_t c = m3[8];
    uint32_t add[8];
#pragma unroll
    for (int i = 0; i < 8; i++) add[i] = nc[i] & (0u - c);
    bn_add8(r.v, m3, add);
    sc_reduce_once(r);
}
HD void sc_mul_inl(sc& r, const sc& a, const sc& b) {
    uint32_t t[16];
    bn_mul_wide(t, a.v, b.v);
    sc_reduce512(r, t);
}
#if defined(__CUDACC__)
static __device__ __noinline__ sc sc_mul_nv(sc a, sc b) { sc r; sc_mul_inl(r, a, b); return r; }
#endif
HD void sc_mul(sc& r, const sc& a, const sc& b) {
#if defined(__CUDA_ARCH__)
    r = sc_mul_nv(a, b);      // one shared body per kernel instead of ~300 inlined instructions per call site
#else
    sc_mul_inl(r, a, b);
#endif
}
HD void sc_sqr(sc& r, const sc& a) { sc_mul(r, a, a); }

// r = round(a*b / 2^384)   (scalar_4x64_impl.h:1071 with shift = 384)
HD void sc_mul_shift_384(sc& r, const sc& a, const sc& b) {
    uint32_t t[16];
    bn_mul_wide(t, a.v, b.v);
    uint32_t rnd = (t[11] >> 31) & 1u;
    uint64_t acc = (uint64_t)t[12] + rnd;
    r.v[0] = (uint32_t)acc; acc >>= 32;
    acc += t[13]; r.v[1] = (uint32_t)acc; acc >>= 32;
    acc += t[14]; r.v[2] = (uint32_t)acc; acc >>= 32;
    acc += t[15]; r.v[3] = (uint32_t)acc; acc >>= 32;
    r.v[4] = (uint32_t)acc;
    r.v[5] = r.v[6] = r.v[7] = 0;
}

// k = r1 + r2*lambda (mod n) with r1, r2 (or their negations) < 2^128  (scalar_impl.h:142-180)
HD void sc_split_lambda(sc& r1, sc& r2, const sc& k) {
    const sc minus_b1 = {{0x0ABFE4C3u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u, 0, 0, 0, 0}};
    const sc minus_b2 = {{0x3DB1562Cu, 0xD765CDA8u, 0x0774346Du, 0x8A280AC5u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}};
    const sc g1 = {{0x45DBB031u, 0xE893209Au, 0x71E8CA7Fu, 0x3DAA8A14u, 0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u}};
    const sc g2 = {{0x8AC47F71u, 0x1571B4AEu, 0x9DF506C6u, 0x221208ACu, 0x0ABFE4C4u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u}};
    const sc lambda = {{0x1B23BD72u, 0xDF02967Cu, 0x20816678u, 0x122E22EAu, 0x8812645Au, 0xA5261C02u, 0xC05C30E0u, 0x5363AD4Cu}};
    sc c1, c2, t;
    sc_mul_shift_384(c1, k, g1);
    sc_mul_shift_384(c2, k, g2);
    sc_mul(c1, c1, minus_b1);
    sc_mul(c2, c2, minus_b2);
    sc_add(r2, c1, c2);
    sc_mul(t, r2, lambda);
    sc_negate(t, t);
    sc_add(r1, t, k);
}

// r = a^(n-2) (a = 0 -> 0); value-identical to the reference's safegcd inverse (scalar_4x64_impl.h:1162)
// NOTE: deliberately NOT an outlined (__noinline__) function.  As a separate function with its 16-entry dynamically
// indexed local table, one sm_100a build of k_bppp_norm_verify computed a wrong inverse deterministically (no sanitizer
// finding; inlining this function, or any unrelated change to the kernel, made it disappear -- A/B libraries on the GPU,
// round 1).  Inlined, the table lives in the kernel's own frame like every other local array.
HD void sc_inverse_fermat(sc& r, const sc& a) {
    // a^(n-2) without any table (r1g: the former 16-entry window table -- a dynamically indexed local array -- was the common factor of two
    // wrong-result builds: the outlined copy in k_bppp_norm_verify (see above) and, inlined into k_rangeproof_rewind, a build whose
    // 512-byte table shared its stack slot with the first 512 bytes of a live 4 KB array of the caller; both clean under
    // compute-sanitizer, both correct when compiled for the host).
    // n - 2 = (2^127 - 1) * 2^129 + L with L = the low 129 bits (bit 128 is 0).  f(k) := a^(2^k - 1), f(i+j) = f(i)^(2^j) * f(j).
    sc f3, f7, f15, f31, f63, t, acc;
    sc_sqr(t, a); sc_mul(t, t, a);                                  // f2
    sc_sqr(t, t); sc_mul(f3, t, a);                                 // f3
    t = f3;
    for (int k = 0; k < 3; k++) sc_sqr(t, t);
    sc_mul(t, t, f3);                                               // f6
    sc_sqr(t, t); sc_mul(f7, t, a);                                 // f7
    t = f7;
    for (int k = 0; k < 7; k++) sc_sqr(t, t);
    sc_mul(t, t, f7);                                               // f14
    sc_sqr(t, t); sc_mul(f15, t, a);                                // f15
    t = f15;
#pragma unroll 1
    for (int k = 0; k < 15; k++) sc_sqr(t, t);
    sc_mul(t, t, f15);                                              // f30
    sc_sqr(t, t); sc_mul(f31, t, a);                                // f31
    t = f31;
#pragma unroll 1
    for (int k = 0; k < 31; k++) sc_sqr(t, t);
    sc_mul(t, t, f31);                                              // f62
    sc_sqr(t, t); sc_mul(f63, t, a);                                // f63
    t = f63;
#pragma unroll 1
    for (int k = 0; k < 63; k++) sc_sqr(t, t);
    sc_mul(t, t, f63);                                              // f126
    sc_sqr(t, t); sc_mul(acc, t, a);                                // f127
    sc_sqr(acc, acc);                                               // bit 128 of the exponent is 0
    const uint32_t lo[4] = {0xD036413Fu, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u};      // low 128 bits of n - 2
#pragma unroll 1
    for (int i = 127; i >= 0; i--) {
        sc_sqr(acc, acc);
        if ((lo[i >> 5] >> (i & 31)) & 1u) sc_mul(acc, acc, a);
    }
    r = acc;
}

}  // namespace B200_NS
