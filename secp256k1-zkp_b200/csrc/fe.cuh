// fe.cuh -- arithmetic in F_p, p = 2^256 - 2^32 - 977 (secp256k1 base field).
//
// Replaces, on the B200, what the reference does in field_5x52_int128_impl.h:18
// (fe_mul_inner), :154 (fe_sqr_inner), field_5x52_impl.h (normalize, negate, add,
// half, set_b32/get_b32) and field_impl.h:37 (sqrt) / field_5x52_impl.h:491 (inv).
//
// Invariant ("weakly reduced"): an fe holds ANY 256-bit value v in [0, 2^256) with
// v = x (mod p).  Because 2^256 = C (mod p) with C = 2^32 + 977, every operation
// folds its overflow back with one multiply by C; the only non-canonical values
// are [p, 2^256), i.e. 2^32+977 of them.  fe_normalize() makes the value canonical.
// This replaces the reference's magnitude bookkeeping (field.h:24, group_impl.h
// magnitude annotations) by a representation that needs none.
#pragma once
#include "bn.cuh"

namespace B200_NS {

struct fe { uint32_t v[8]; };

#define FE_C0 0x000003D1u  /* C = 2^32 + 977: low limb 977, limb1 = 1 */
#define FE_P0 0xFFFFFC2Fu
#define FE_P1 0xFFFFFFFEu

HD void fe_set_zero(fe& r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
}
HD void fe_set_int(fe& r, uint32_t a) { fe_set_zero(r); r.v[0] = a; }

// r = t[0..15] mod p (weakly reduced).  t < 2^512.
HD void fe_reduce512(fe& r, const uint32_t* t) {
#if BN_PTX
    uint32_t lo[10];
#pragma unroll
    for (int i = 0; i < 8; i++) lo[i] = t[i];
    lo[8] = 0; lo[9] = 0;
    const uint32_t* h = t + 8;
    // lo[0..9] = t_lo + hi*977 + (hi << 32)      (< 2^289, so lo[9] <= 1)
    bn_cmad4(lo, h[0], h[2], h[4], h[6], 977u);
    bn_cmad4<false>(lo + 1, h[1], h[3], h[5], h[7], 977u);      // top word: lo[8] (0/1) + hi(h7*977) (< 977) + carry: cannot overflow
    asm("add.cc.u32  %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.cc.u32 %7, %7, %16;\n\t"
        "addc.u32    %8, %8, 0;\n\t"
        : "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7]), "+r"(lo[8]), "+r"(lo[9])
        : "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]));
    // second fold: Th = lo[8] + lo[9]*2^32 ; u = Th*C = Th*977 + (Th<<32) as limbs (v0,v1,v2)
    uint32_t v0, v1, v2, c;
    asm("mul.lo.u32  %0, %3, 977;\n\t"
        "mul.hi.u32  %1, %3, 977;\n\t"
        "mad.lo.u32  %1, %4, 977, %1;\n\t"   // lo[9] in {0,1}: no overflow (m1 < 977 + 977)
        "add.cc.u32  %1, %1, %3;\n\t"
        "addc.u32    %2, %4, 0;\n\t"
        : "=r"(v0), "=r"(v1), "=r"(v2)
        : "r"(lo[8]), "r"(lo[9]));
    asm("add.cc.u32  %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32    %8, 0, 0;\n\t"
        : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7]), "=r"(c)
        : "r"(v0), "r"(v1), "r"(v2));
    // third fold: carry c means value = 2^256 + (small < 2^67): add C once, touches limbs 0..2 only
    asm("mad.lo.cc.u32 %0, %3, 977, %0;\n\t"
        "addc.cc.u32   %1, %1, %3;\n\t"
        "addc.u32      %2, %2, 0;\n\t"
        : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2])
        : "r"(c));
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = lo[i];
#else
    uint32_t o[8];
    uint64_t carry = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t acc = (uint64_t)t[i] + (uint64_t)t[8 + i] * 977u + (i > 0 ? (uint64_t)t[8 + i - 1] : 0) + carry;
        o[i] = (uint32_t)acc;
        carry = acc >> 32;
    }
    uint64_t th = carry + t[15];            // < 2^33 + 2^11
    uint64_t m = th * 977u;                 // < 2^44
    uint64_t acc = (uint64_t)o[0] + (m & 0xffffffffu);
    o[0] = (uint32_t)acc; carry = acc >> 32;
    acc = (uint64_t)o[1] + (m >> 32) + (th & 0xffffffffu) + carry;
    o[1] = (uint32_t)acc; carry = acc >> 32;
    acc = (uint64_t)o[2] + (th >> 32) + carry;
    o[2] = (uint32_t)acc; carry = acc >> 32;
    for (int i = 3; i < 8; i++) {
        acc = (uint64_t)o[i] + carry;
        o[i] = (uint32_t)acc; carry = acc >> 32;
    }
    uint32_t c = (uint32_t)carry;           // 0/1
    acc = (uint64_t)o[0] + (uint64_t)c * 977u;
    o[0] = (uint32_t)acc; carry = acc >> 32;
    acc = (uint64_t)o[1] + c + carry;
    o[1] = (uint32_t)acc; carry = acc >> 32;
    o[2] += (uint32_t)carry;
    for (int i = 0; i < 8; i++) r.v[i] = o[i];
#endif
}

HD void fe_mul_inl(fe& r, const fe& a, const fe& b) {
    uint32_t t[16];
    bn_mul_wide(t, a.v, b.v);
    fe_reduce512(r, t);
}
HD void fe_sqr_inl(fe& r, const fe& a) {
    uint32_t t[16];
    bn_sqr_wide(t, a.v);
    fe_reduce512(r, t);
}
// In the large kernels fe_mul / fe_sqr are real (non-inlined) functions: operands travel BY VALUE in registers
// (checked in SASS: CALL.REL.NOINC, no local-memory traffic).  Fully inlined, a verify kernel is several hundred KB
// of code and ncu shows it instruction-fetch bound (stall_no_instruction dominant); with ~2 KB bodies shared by all
// call sites the hot loops sit in the instruction cache.  The field micro-kernel (k_fe.cu) keeps the inlined form.
#if defined(__CUDACC__)
static __device__ __noinline__ fe fe_mul_nv(fe a, fe b) { fe r; fe_mul_inl(r, a, b); return r; }
static __device__ __noinline__ fe fe_sqr_nv(fe a) { fe r; fe_sqr_inl(r, a); return r; }
#endif
HD void fe_mul(fe& r, const fe& a, const fe& b) {
#if defined(__CUDA_ARCH__) && !defined(B200_INLINE_FE)
    r = fe_mul_nv(a, b);
#else
    fe_mul_inl(r, a, b);
#endif
}
HD void fe_sqr(fe& r, const fe& a) {
#if defined(__CUDA_ARCH__) && !defined(B200_INLINE_FE)
    r = fe_sqr_nv(a);
#else
    fe_sqr_inl(r, a);
#endif
}
// INL = true: always the inlined body (used inside the outlined group operations, ge.cuh); false: the default policy
template <bool INL> HD void fe_mul_t(fe& r, const fe& a, const fe& b) { if (INL) fe_mul_inl(r, a, b); else fe_mul(r, a, b); }
template <bool INL> HD void fe_sqr_t(fe& r, const fe& a) { if (INL) fe_sqr_inl(r, a); else fe_sqr(r, a); }

// r = a + b
HD void fe_add(fe& r, const fe& a, const fe& b) {
    uint32_t t[8];
    uint32_t c = bn_add8(t, a.v, b.v);
    // fold carry: += c*C ; a second carry can only happen from a tiny value, touching limbs 0..1
    uint32_t cc[8] = {c * FE_C0, c, 0, 0, 0, 0, 0, 0};
    uint32_t c2 = bn_add8(r.v, t, cc);
    uint64_t acc = (uint64_t)r.v[0] + (uint64_t)c2 * FE_C0;
    r.v[0] = (uint32_t)acc;
    r.v[1] += c2 + (uint32_t)(acc >> 32);
}
// r = a - b
HD void fe_sub(fe& r, const fe& a, const fe& b) {
    uint32_t t[8];
    uint32_t bw = bn_sub8(t, a.v, b.v);
    // borrow: true value = t - 2^256 = t - C (mod p)
    uint32_t cc[8] = {bw * FE_C0, bw, 0, 0, 0, 0, 0, 0};
    uint32_t b2 = bn_sub8(r.v, t, cc);
    // second borrow: r = 2^256 - k (k <= C); subtract C again: affects limbs 0..1 only
    uint64_t lo = ((uint64_t)r.v[1] << 32) | r.v[0];
    lo -= (uint64_t)b2 * (((uint64_t)1 << 32) + FE_C0);
    r.v[0] = (uint32_t)lo;
    r.v[1] = (uint32_t)(lo >> 32);
}
HD void fe_neg(fe& r, const fe& a) {
    fe z; fe_set_zero(z);
    fe_sub(r, z, a);
}
HD void fe_dbl(fe& r, const fe& a) { fe_add(r, a, a); }

// r = a * k for small k (k < 2^20)
HD void fe_mul_small(fe& r, const fe& a, uint32_t k) {
    uint32_t o[8];
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t acc = (uint64_t)a.v[i] * k + carry;
        o[i] = (uint32_t)acc;
        carry = acc >> 32;
    }
    // top = carry < 2^20 ; top*C < 2^53
    uint64_t u = carry * ((((uint64_t)1) << 32) + FE_C0);
    uint32_t cc[8] = {(uint32_t)u, (uint32_t)(u >> 32), 0, 0, 0, 0, 0, 0};
    uint32_t c2 = bn_add8(r.v, o, cc);
    uint64_t acc = (uint64_t)r.v[0] + (uint64_t)c2 * FE_C0;
    r.v[0] = (uint32_t)acc;
    r.v[1] += c2 + (uint32_t)(acc >> 32);
}

// r = a / 2
HD void fe_half(fe& r, const fe& a) {
    uint32_t mask = 0u - (a.v[0] & 1u);
    uint32_t pm[8] = {FE_P0 & mask, FE_P1 & mask, mask, mask, mask, mask, mask, mask};
    uint32_t t[8];
    uint32_t c = bn_add8(t, a.v, pm);
#pragma unroll
    for (int i = 0; i < 7; i++) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
    r.v[7] = (t[7] >> 1) | (c << 31);
}

// canonical representative in [0, p)
HD void fe_normalize(fe& r) {
    uint32_t cc[8] = {FE_C0, 1, 0, 0, 0, 0, 0, 0};
    uint32_t t[8];
    uint32_t c = bn_add8(t, r.v, cc);   // carry <=> r >= p ; then t = r - p
    uint32_t m = 0u - c;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (t[i] & m) | (r.v[i] & ~m);
}
HD int fe_is_zero(const fe& a) {
    uint32_t z = a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7];
    uint32_t q = (a.v[0] ^ FE_P0) | (a.v[1] ^ FE_P1) | ~(a.v[2] & a.v[3] & a.v[4] & a.v[5] & a.v[6] & a.v[7]);
    return (z == 0) | (q == 0);
}
HD int fe_equal(const fe& a, const fe& b) {
    fe d; fe_sub(d, a, b);
    return fe_is_zero(d);
}
// parity of the canonical value
HD int fe_is_odd(const fe& a) {
    fe t = a; fe_normalize(t);
    return (int)(t.v[0] & 1u);
}
HD void fe_cmov(fe& r, const fe& a, int flag) {
    uint32_t m = 0u - (uint32_t)(flag != 0);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (a.v[i] & m) | (r.v[i] & ~m);
}

// 32 big-endian bytes -> fe.  returns 1 if the value was < p ("set_b32_limit", field_5x52_impl.h:265);
// the stored value is the raw 256-bit integer either way ("set_b32_mod" semantics, :228).
HD int fe_set_b32(fe& r, const uint8_t* b) {
    bn_from_be32(r.v, b);
    uint32_t cc[8] = {FE_C0, 1, 0, 0, 0, 0, 0, 0};
    uint32_t t[8];
    return bn_add8(t, r.v, cc) == 0;
}
// canonical 32 big-endian bytes (field_5x52_impl.h:271)
HD void fe_get_b32(uint8_t* b, const fe& a) {
    fe t = a; fe_normalize(t);
    bn_to_be32(b, t.v);
}

HD void fe_sqr_n(fe& r, const fe& a, int n) {
    r = a;
    for (int i = 0; i < n; i++) fe_sqr(r, r);
}

// shared prefix of the inverse / sqrt exponent chains: x2 = a^3, x22 = a^(2^22-1), x223 = a^(2^223-1)
HD void fe_pow_chain(fe& x2, fe& x22, fe& x223, const fe& a) {
    fe x3, x6, x9, x11, x44, x88, x176, x220, t;
    fe_sqr(t, a); fe_mul(x2, t, a);
    fe_sqr(t, x2); fe_mul(x3, t, a);
    fe_sqr_n(t, x3, 3); fe_mul(x6, t, x3);
    fe_sqr_n(t, x6, 3); fe_mul(x9, t, x3);
    fe_sqr_n(t, x9, 2); fe_mul(x11, t, x2);
    fe_sqr_n(t, x11, 11); fe_mul(x22, t, x11);
    fe_sqr_n(t, x22, 22); fe_mul(x44, t, x22);
    fe_sqr_n(t, x44, 44); fe_mul(x88, t, x44);
    fe_sqr_n(t, x88, 88); fe_mul(x176, t, x88);
    fe_sqr_n(t, x176, 44); fe_mul(x220, t, x44);
    fe_sqr_n(t, x220, 3); fe_mul(x223, t, x3);
}

// r = a^(p-2)  (a = 0 -> 0).  Any exact inverse gives the reference's fe_inv_var value (field_5x52_impl.h:491 uses safegcd; results are
// unique mod p).  Round-1 form, kept as the cross-check of fe_inv (modinv.cuh), which is what the kernels call now.
HDNI void fe_inv_fermat(fe& r, const fe& a) {
    fe x2, x22, x223, t;
    fe_pow_chain(x2, x22, x223, a);
    // p - 2 = (2^223-1) << 33 | 2^32... : bits: 223 ones, 0, 22 ones, 0000, 1, 0, 11, 0, 1  (binary of p-2 tail ...FFFFFC2D)
    fe_sqr_n(t, x223, 23); fe_mul(t, t, x22);
    fe_sqr_n(t, t, 5); fe_mul(t, t, a);
    fe_sqr_n(t, t, 3); fe_mul(t, t, x2);
    fe_sqr_n(t, t, 2); fe_mul(r, t, a);
}

// r = a^((p+1)/4); returns 1 iff r^2 == a (field_impl.h:37-146 semantics)
HDNI int fe_sqrt(fe& r, const fe& a) {
    fe x2, x22, x223, t;
    fe_pow_chain(x2, x22, x223, a);
    fe_sqr_n(t, x223, 23); fe_mul(t, t, x22);
    fe_sqr_n(t, t, 6); fe_mul(t, t, x2);
    fe_sqr_n(t, t, 2);
    fe chk; fe_sqr(chk, t);
    r = t;
    return fe_equal(chk, a);
}

// is_square(a) <=> sqrt succeeds (round-1 form; fe_is_square in modinv.cuh computes the Jacobi symbol instead)
HD int fe_is_square_sqrt(const fe& a) {
    fe r;
    return fe_sqrt(r, a);
}

}  // namespace B200_NS
