// ge.cuh -- secp256k1 group law: affine `ge`, Jacobian `gej` (reference: group.h:16-41, group_impl.h).
//
// Complete "_var" semantics: every add handles infinity, equal inputs (-> doubling) and opposite
// inputs (-> infinity) exactly like gej_add_var / gej_add_ge_var (group_impl.h:534-659), because
// adversarial verification inputs can reach those branches.  Formulas are the textbook a=0 Jacobian
// ones; only serialized results have to match the reference bit for bit, and those are unique.
#pragma once
#include "modinv.cuh"

namespace B200_NS {

// -DB200_DBL_2M5S selects the 2M + 5S doubling (dbl-2009-l); measured SLOWER than the 3M + 4S form on B200 (42.2 vs 43.2 M verify/s,
// r1i): the squaring it saves costs less than the four extra field additions it needs.
struct ge { fe x, y; int inf; };
struct gej { fe x, y, z; int inf; };

HD void ge_set_infinity(ge& r) { fe_set_zero(r.x); fe_set_zero(r.y); r.inf = 1; }
HD void gej_set_infinity(gej& r) { fe_set_zero(r.x); fe_set_zero(r.y); fe_set_zero(r.z); r.inf = 1; }
HD void gej_set_ge(gej& r, const ge& a) { r.x = a.x; r.y = a.y; fe_set_int(r.z, 1); r.inf = a.inf; }
HD void ge_neg(ge& r, const ge& a) { r.x = a.x; fe_neg(r.y, a.y); r.inf = a.inf; }
HD void gej_neg(gej& r, const gej& a) { r.x = a.x; fe_neg(r.y, a.y); r.z = a.z; r.inf = a.inf; }

// y^2 == x^3 + 7  (group_impl.h ge_is_valid_var)
HD int ge_is_valid(const ge& a) {
    if (a.inf) return 0;
    fe y2, x3, seven;
    fe_sqr(y2, a.y);
    fe_sqr(x3, a.x); fe_mul(x3, x3, a.x);
    fe_set_int(seven, 7);
    fe_add(x3, x3, seven);
    return fe_equal(y2, x3);
}

// r = 2a.  L = 3/2 X^2; X3 = L^2 - 2XY^2; Y3 = L(XY^2 - X3) - Y^4; Z3 = YZ   (3M + 4S)
// (same operation count as group_impl.h:468-501)
template <bool INL>
HD void gej_double_t(gej& r, const gej& a) {
    // on secp256k1 (odd order, no 2-torsion) 2a = inf only for a = inf
#ifdef B200_DBL_2M5S
    // dbl-2009-l (a = 0): A = X^2, B = Y^2, C = B^2, D = 2((X+B)^2 - A - C), E = 3A, X3 = E^2 - 2D, Y3 = E(D - X3) - 8C, Z3 = 2YZ
    // 2M + 5S: one multiplication of the 3M + 4S form below becomes a squaring (a squaring costs ~0.68 of a multiplication here)
    fe A, B, C, D, E, t, u;
    int inf = a.inf;
    fe_mul_t<INL>(u, a.y, a.z);
    fe_dbl(u, u);                        // Z3
    fe_sqr_t<INL>(A, a.x);
    fe_sqr_t<INL>(B, a.y);
    fe_sqr_t<INL>(C, B);
    fe_add(t, a.x, B);
    fe_sqr_t<INL>(D, t);
    fe_sub(D, D, A);
    fe_sub(D, D, C);
    fe_dbl(D, D);
    fe_mul_small(E, A, 3);
    fe x3;
    fe_sqr_t<INL>(x3, E);
    fe_sub(x3, x3, D);
    fe_sub(x3, x3, D);
    fe_sub(t, D, x3);
    fe_mul_t<INL>(t, t, E);
    fe_mul_small(C, C, 8);
    fe_sub(r.y, t, C);
    r.x = x3;
    r.z = u;
    r.inf = inf;
#else
    fe l, s, t, u;
    int inf = a.inf;
    fe_mul_t<INL>(u, a.y, a.z);          // Z3
    fe_sqr_t<INL>(s, a.y);               // Y^2
    fe_sqr_t<INL>(l, a.x);
    fe_mul_small(l, l, 3);
    fe_half(l, l);                // L = 3/2 X^2
    fe_mul_t<INL>(t, a.x, s);            // X Y^2
    fe_sqr_t<INL>(s, s);                 // Y^4
    fe x3;
    fe_sqr_t<INL>(x3, l);
    fe_sub(x3, x3, t);
    fe_sub(x3, x3, t);
    fe_sub(t, t, x3);
    fe_mul_t<INL>(t, t, l);
    fe_sub(r.y, t, s);
    r.x = x3;
    r.z = u;
    r.inf = inf;
#endif
}

template <bool INL> HD void gej_double_dispatch(gej& r, const gej& a);
// r = a + b, b affine.  If rzr != nullptr and the generic branch is taken, *rzr = r.z / a.z.
// (group_impl.h:598-659; 8M + 3S)
template <bool INL>
HD void gej_add_ge_t(gej& r, const gej& a, const ge& b, fe* rzr = nullptr) {
    if (a.inf) {
        // rzr undefined in this case, as in the reference
        gej_set_ge(r, b);
        return;
    }
    if (b.inf) {
        if (rzr) fe_set_int(*rzr, 1);
        r = a;
        return;
    }
    fe z12, u2, s2, h, rr;
    fe_sqr_t<INL>(z12, a.z);
    fe_mul_t<INL>(u2, b.x, z12);
    fe_mul_t<INL>(s2, b.y, a.z);
    fe_mul_t<INL>(s2, s2, z12);
    fe_sub(h, u2, a.x);
    fe_sub(rr, s2, a.y);
    if (fe_is_zero(h)) {
        if (fe_is_zero(rr)) {
            // a == b: double.  r.z = 2*... relation: Z3 = Y*Z so rzr = a.y (Jacobian y of a)
#ifdef B200_DBL_2M5S
            if (rzr) fe_dbl(*rzr, a.y);      // Z3 = 2*Y*Z
#else
            if (rzr) *rzr = a.y;
#endif
            gej_double_dispatch<INL>(r, a);
        } else {
            if (rzr) fe_set_zero(*rzr);
            gej_set_infinity(r);
        }
        return;
    }
    fe hh, hhh, v, t;
    fe_sqr_t<INL>(hh, h);
    fe_mul_t<INL>(hhh, hh, h);
    fe_mul_t<INL>(v, a.x, hh);
    fe_sqr_t<INL>(t, rr);
    fe_sub(t, t, hhh);
    fe_sub(t, t, v);
    fe_sub(t, t, v);              // X3
    fe_sub(v, v, t);
    fe_mul_t<INL>(v, v, rr);
    fe_mul_t<INL>(hhh, hhh, a.y);
    if (rzr) *rzr = h;
    fe_mul_t<INL>(r.z, a.z, h);
    fe_sub(r.y, v, hhh);
    r.x = t;
    r.inf = 0;
}

HD void gej_double(gej& r, const gej& a) { gej_double_t<false>(r, a); }
HD void gej_add_ge(gej& r, const gej& a, const ge& b, fe* rzr = nullptr) { gej_add_ge_t<false>(r, a, b, rzr); }

// Outlined group operations for the hot loops (ecmult main loop, comb accumulation): ONE shared body each, with the field
// multiplications inlined INSIDE it.  Compared with calling the outlined fe_mul 7 / 11 times this removes the
// argument-marshalling moves around every call (ncu r1d: 14 % of all executed instructions were IMAD.MOV on the multiply
// pipe); the point travels through local memory (L1) instead, ~50 accesses per ~1000-instruction operation.
#if defined(__CUDACC__)
static __device__ __noinline__ void gej_double_nv(gej* r, const gej* a) { gej t; gej_double_t<true>(t, *a); *r = t; }
static __device__ __noinline__ void gej_add_ge_nv(gej* r, const gej* a, const ge* b) { gej t; gej_add_ge_t<true>(t, *a, *b, nullptr); *r = t; }
#endif
template <bool INL> HD void gej_double_dispatch(gej& r, const gej& a) {
#if defined(__CUDA_ARCH__)
    if (INL) { gej_double_nv(&r, &a); return; }      // inside an outlined add: do not inline a second doubling body
#endif
    gej_double_t<false>(r, a);
}
// r = 2a / r = a + b for the hot loops (r must not alias a)
HD void gej_double_hot(gej& r, const gej& a) {
#if defined(__CUDA_ARCH__) && defined(B200_HOT_INLINE)
    gej_double_t<true>(r, a);                        // field ops inlined right here: no call, point stays in registers
#elif defined(__CUDA_ARCH__) && defined(B200_OUTLINE_GE)
    gej_double_nv(&r, &a);
#else
    gej_double(r, a);
#endif
}
HD void gej_add_ge_hot(gej& r, const gej& a, const ge& b) {
#if defined(__CUDA_ARCH__) && defined(B200_HOT_INLINE)
    gej_add_ge_t<true>(r, a, b, nullptr);
#elif defined(__CUDA_ARCH__) && defined(B200_OUTLINE_GE)
    gej_add_ge_nv(&r, &a, &b);
#else
    gej_add_ge(r, a, b);
#endif
}

// r = a + b, both Jacobian (group_impl.h:534-596; 12M + 4S)
HD void gej_add(gej& r, const gej& a, const gej& b) {
    if (a.inf) { r = b; return; }
    if (b.inf) { r = a; return; }
    fe z22, z12, u1, u2, s1, s2, h, rr;
    fe_sqr(z22, b.z);
    fe_sqr(z12, a.z);
    fe_mul(u1, a.x, z22);
    fe_mul(u2, b.x, z12);
    fe_mul(s1, a.y, z22); fe_mul(s1, s1, b.z);
    fe_mul(s2, b.y, z12); fe_mul(s2, s2, a.z);
    fe_sub(h, u2, u1);
    fe_sub(rr, s2, s1);
    if (fe_is_zero(h)) {
        if (fe_is_zero(rr)) gej_double(r, a);
        else gej_set_infinity(r);
        return;
    }
    fe hh, hhh, v, t;
    fe_sqr(hh, h);
    fe_mul(hhh, hh, h);
    fe_mul(v, u1, hh);
    fe_sqr(t, rr);
    fe_sub(t, t, hhh);
    fe_sub(t, t, v);
    fe_sub(t, t, v);
    fe_sub(v, v, t);
    fe_mul(v, v, rr);
    fe_mul(hhh, hhh, s1);
    fe_mul(u1, a.z, b.z);
    fe_mul(r.z, u1, h);
    fe_sub(r.y, v, hhh);
    r.x = t;
    r.inf = 0;
}

// affine <- Jacobian with a known inverse of z
HD void ge_set_gej_zinv(ge& r, const gej& a, const fe& zi) {
    fe zi2, zi3;
    fe_sqr(zi2, zi);
    fe_mul(zi3, zi2, zi);
    fe_mul(r.x, a.x, zi2);
    fe_mul(r.y, a.y, zi3);
    r.inf = a.inf;
}
// affine <- Jacobian (one field inversion; group_impl.h:177-196)
HD void ge_set_gej(ge& r, const gej& a) {
    if (a.inf) { ge_set_infinity(r); return; }
    fe zi;
    fe_inv(zi, a.z);
    ge_set_gej_zinv(r, a, zi);
}

// y from x: returns 0 if x^3+7 is not a square.  The root returned is THE square root that is itself
// a square ("xquad", group_impl.h:347-361) -- a^((p+1)/4) always is.
HD int ge_set_xquad(ge& r, const fe& x) {
    fe c, seven;
    fe_sqr(c, x); fe_mul(c, c, x);
    fe_set_int(seven, 7);
    fe_add(c, c, seven);
    r.x = x;
    r.inf = 0;
    return fe_sqrt(r.y, c);
}
// y from x with chosen parity (group_impl.h:363-373)
HD int ge_set_xo(ge& r, const fe& x, int odd) {
    if (!ge_set_xquad(r, x)) return 0;
    if (fe_is_odd(r.y) != odd) fe_neg(r.y, r.y);
    return 1;
}

// x of the Jacobian point equals the affine field value x?  (inversion-free, group_impl.h:425-433)
HD int gej_eq_x(const fe& x, const gej& a) {
    fe z2, t;
    fe_sqr(z2, a.z);
    fe_mul(t, x, z2);
    return fe_equal(t, a.x);
}

}  // namespace B200_NS
