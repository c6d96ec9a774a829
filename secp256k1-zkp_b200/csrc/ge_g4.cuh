// ge_g4.cuh -- group operations spread over a GROUP OF 4 LANES, for the latency-bound serial chains of the MSM tail
// (window Horner: ~128 dependent doublings; running sums of a few buckets; sums of a handful of partial results).
//
// One thread executing gej_double runs its 7 field multiplications one after the other (~0.3 us each: a single warp can issue
// one IMAD.WIDE per 4 cycles whether 1 or 32 of its lanes are active).  Here the 4 lanes of a group hold IDENTICAL copies of the
// operands, each lane multiplies a different pair of field elements at every step, and the products are exchanged with
// __shfl_sync: a doubling is 3 multiplication steps deep instead of 7, a Jacobian addition 6 instead of 16, a mixed addition 6
// instead of 11.  Squarings are executed as multiplications so that all lanes run the same instruction stream.
// Contract: the 4 lanes of a group (lane & ~3 .. +3) call with identical arguments and receive identical results; different
// groups of a warp may take different paths (the shuffles are masked to the group).
// Host build (tests/hostemu): the same names fall back to the plain one-thread functions.
#pragma once
#include "ge.cuh"

namespace B200_NS {

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t g4_mask() { return 0xFu << ((threadIdx.x & 31u) & ~3u); }
__device__ __forceinline__ fe g4_bcast(const fe& a, int src) {
    fe r;
    const uint32_t m = g4_mask();
    const int lane = (int)((threadIdx.x & 31u) & ~3u) + src;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = __shfl_sync(m, a.v[k], lane);
    return r;
}
__device__ __forceinline__ fe g4_sel(int l, const fe& a0, const fe& a1, const fe& a2, const fe& a3) {
    fe r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = (l == 0) ? a0.v[k] : (l == 1) ? a1.v[k] : (l == 2) ? a2.v[k] : a3.v[k];
    return r;
}
__device__ __forceinline__ fe g4_sel2(int l, const fe& a0, const fe& a1) {      // lanes 0,2 -> a0 ; lanes 1,3 -> a1
    fe r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = (l & 1) ? a1.v[k] : a0.v[k];
    return r;
}

static __device__ __noinline__ void gej_double_g4(gej& r, const gej& a) {
    if (a.inf) { r = a; return; }
    const int l = (int)(threadIdx.x & 3u);
    fe p, q, o;
    // step 1: Y*Z | Y*Y | X*X
    p = g4_sel(l, a.y, a.y, a.x, a.x);
    q = g4_sel(l, a.z, a.y, a.x, a.x);
    fe_mul(o, p, q);
    fe u = g4_bcast(o, 0), s = g4_bcast(o, 1), lq = g4_bcast(o, 2);
    fe_mul_small(lq, lq, 3);
    fe_half(lq, lq);                                  // L = 3/2 X^2
    // step 2: X*s | s*s | L*L
    p = g4_sel(l, a.x, s, lq, lq);
    q = g4_sel(l, s, s, lq, lq);
    fe_mul(o, p, q);
    fe t = g4_bcast(o, 0), ss = g4_bcast(o, 1), x3 = g4_bcast(o, 2);
    fe_sub(x3, x3, t);
    fe_sub(x3, x3, t);
    fe_sub(t, t, x3);
    // step 3 (every lane, no exchange): (t - x3) * L
    fe_mul(o, t, lq);
    fe_sub(r.y, o, ss);
    r.x = x3;
    r.z = u;
    r.inf = 0;
}

static __device__ __noinline__ void gej_add_g4(gej& r, const gej& a, const gej& b) {
    if (a.inf) { r = b; return; }
    if (b.inf) { r = a; return; }
    const int l = (int)(threadIdx.x & 3u);
    fe p, q, o;
    // step 1: z1^2 | z2^2 | z1*z2
    p = g4_sel(l, a.z, b.z, a.z, a.z);
    q = g4_sel(l, a.z, b.z, b.z, b.z);
    fe_mul(o, p, q);
    fe z12 = g4_bcast(o, 0), z22 = g4_bcast(o, 1), zz = g4_bcast(o, 2);
    // step 2: x1*z2^2 | x2*z1^2 | y1*z2 | y2*z1
    p = g4_sel(l, a.x, b.x, a.y, b.y);
    q = g4_sel(l, z22, z12, b.z, a.z);
    fe_mul(o, p, q);
    fe u1 = g4_bcast(o, 0), u2 = g4_bcast(o, 1), t1 = g4_bcast(o, 2), t2 = g4_bcast(o, 3);
    // step 3: t1*z2^2 | t2*z1^2
    p = g4_sel2(l, t1, t2);
    q = g4_sel2(l, z22, z12);
    fe_mul(o, p, q);
    fe s1 = g4_bcast(o, 0), s2 = g4_bcast(o, 1);
    fe h, rr;
    fe_sub(h, u2, u1);
    fe_sub(rr, s2, s1);
    if (fe_is_zero(h)) {
        if (fe_is_zero(rr)) gej_double_g4(r, a);
        else gej_set_infinity(r);
        return;
    }
    // step 4: h^2 | rr^2 | zz*h
    p = g4_sel(l, h, rr, zz, zz);
    q = g4_sel(l, h, rr, h, h);
    fe_mul(o, p, q);
    fe hh = g4_bcast(o, 0), x3 = g4_bcast(o, 1), z3 = g4_bcast(o, 2);
    // step 5: hh*h | u1*hh
    p = g4_sel2(l, hh, u1);
    q = g4_sel2(l, h, hh);
    fe_mul(o, p, q);
    fe hhh = g4_bcast(o, 0), v = g4_bcast(o, 1);
    fe_sub(x3, x3, hhh);
    fe_sub(x3, x3, v);
    fe_sub(x3, x3, v);
    fe_sub(v, v, x3);
    // step 6: (v - x3)*rr | hhh*s1
    p = g4_sel2(l, v, hhh);
    q = g4_sel2(l, rr, s1);
    fe_mul(o, p, q);
    fe m = g4_bcast(o, 0), w = g4_bcast(o, 1);
    fe_sub(r.y, m, w);
    r.x = x3;
    r.z = z3;
    r.inf = 0;
}

// r = a + b, b affine
static __device__ __noinline__ void gej_add_ge_g4(gej& r, const gej& a, const ge& b) {
    if (a.inf) { gej_set_ge(r, b); return; }
    if (b.inf) { r = a; return; }
    const int l = (int)(threadIdx.x & 3u);
    fe p, q, o, z12;
    fe_mul(z12, a.z, a.z);                                      // step 1 (every lane)
    // step 2: x2*z1^2 | y2*z1
    p = g4_sel2(l, b.x, b.y);
    q = g4_sel2(l, z12, a.z);
    fe_mul(o, p, q);
    fe u2 = g4_bcast(o, 0), t2 = g4_bcast(o, 1), s2;
    fe_mul(s2, t2, z12);                                        // step 3 (every lane)
    fe h, rr;
    fe_sub(h, u2, a.x);
    fe_sub(rr, s2, a.y);
    if (fe_is_zero(h)) {
        if (fe_is_zero(rr)) gej_double_g4(r, a);
        else gej_set_infinity(r);
        return;
    }
    // step 4: h^2 | rr^2 | z1*h
    p = g4_sel(l, h, rr, a.z, a.z);
    q = g4_sel(l, h, rr, h, h);
    fe_mul(o, p, q);
    fe hh = g4_bcast(o, 0), x3 = g4_bcast(o, 1), z3 = g4_bcast(o, 2);
    // step 5: hh*h | x1*hh
    p = g4_sel2(l, hh, a.x);
    q = g4_sel2(l, h, hh);
    fe_mul(o, p, q);
    fe hhh = g4_bcast(o, 0), v = g4_bcast(o, 1);
    fe_sub(x3, x3, hhh);
    fe_sub(x3, x3, v);
    fe_sub(x3, x3, v);
    fe_sub(v, v, x3);
    // step 6: (v - x3)*rr | hhh*y1
    p = g4_sel2(l, v, hhh);
    q = g4_sel2(l, rr, a.y);
    fe_mul(o, p, q);
    fe m = g4_bcast(o, 0), w = g4_bcast(o, 1);
    fe_sub(r.y, m, w);
    r.x = x3;
    r.z = z3;
    r.inf = 0;
}
#else
HD void gej_double_g4(gej& r, const gej& a) { gej t; gej_double(t, a); r = t; }
HD void gej_add_g4(gej& r, const gej& a, const gej& b) { gej t; gej_add(t, a, b); r = t; }
HD void gej_add_ge_g4(gej& r, const gej& a, const ge& b) { gej t; gej_add_ge(t, a, b); r = t; }
#endif

}  // namespace B200_NS
