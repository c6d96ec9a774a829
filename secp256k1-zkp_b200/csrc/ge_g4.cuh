// ge_g4.cuh -- group operations spread over a GROUP OF 4 LANES, for the latency-bound serial chains of the MSM tail
// (window Horner: ~128 dependent doublings; running sums of a few buckets; sums of a handful of partial results).
//
// One thread executing gej_double runs its 7 field multiplications one after the other (~0.3 us each: a single warp can issue
// one IMAD.WIDE per 4 cycles whether 1 or 32 of its lanes are active).  Here the 4 lanes of a group hold IDENTICAL copies of the
// operands, each lane multiplies a different pair of field elements at every step, and the products are exchanged with
// __shfl_sync: a doubling is 3 multiplication steps deep instead of 7, a Jacobian addition 6 instead of 16, a mixed addition 6
// instead of 11.  Squarings are executed as multiplications so that all lanes run the same instruction stream.
//
// Contract (r2b): ALL 32 LANES OF THE WARP call together, converged; the 4 lanes of a group (lane & ~3 .. +3) pass identical arguments
// and receive identical results; the 8 groups of a warp work on 8 different items.  Control flow inside is warp-uniform (the
// degenerate cases are resolved by selects, or by branches taken on a warp vote), so every shuffle is a plain full-mask SHFL.  The
// first version masked the shuffles to the group and let groups branch independently: ptxas then wraps every exchange in a
// WARPSYNC.COLLECTIVE region and a doubling took ~3 us instead of ~1 (profiles/r2a: k_msm_tail 600 us for ~130 doublings).
// Host build (tests/hostemu): the same names fall back to the plain one-thread functions.
#pragma once
#include "ge.cuh"

namespace B200_NS {

#if defined(__CUDA_ARCH__)
#define G4_FULL 0xffffffffu
// -DB200_G4_INLINE: the operations (and the field multiplications inside them) are inlined at their call sites -- for kernels that are
// nothing but a latency-bound chain (k_msm_tail, k_msm_r1<true>): no argument copies through local memory, no call overhead.
#ifdef B200_G4_INLINE
#define G4_FN static __device__ __forceinline__
#define G4_MUL(o, p, q) fe_mul_inl(o, p, q)
#else
#define G4_FN static __device__ __noinline__
#define G4_MUL(o, p, q) fe_mul(o, p, q)
#endif
__device__ __forceinline__ fe g4_bcast(const fe& a, int src) {
    fe r;
    const int lane = (int)((threadIdx.x & 31u) & ~3u) + src;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = __shfl_sync(G4_FULL, a.v[k], lane);
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
__device__ __forceinline__ void g4_gej_select(gej& r, const gej& a, bool take) {   // r = take ? a : r
    if (take) r = a;
}

// r = 2a.  Branch free: an infinite input just carries its flag through (the arithmetic runs on whatever the coordinates hold).
G4_FN void gej_double_g4(gej& r, const gej& a) {
    const int l = (int)(threadIdx.x & 3u);
    fe p, q, o;
    // step 1: Y*Z | Y*Y | X*X
    p = g4_sel(l, a.y, a.y, a.x, a.x);
    q = g4_sel(l, a.z, a.y, a.x, a.x);
    G4_MUL(o, p, q);
    fe u = g4_bcast(o, 0), s = g4_bcast(o, 1), lq = g4_bcast(o, 2);
    fe_mul_small(lq, lq, 3);
    fe_half(lq, lq);                                  // L = 3/2 X^2
    // step 2: X*s | s*s | L*L
    p = g4_sel(l, a.x, s, lq, lq);
    q = g4_sel(l, s, s, lq, lq);
    G4_MUL(o, p, q);
    fe t = g4_bcast(o, 0), ss = g4_bcast(o, 1), x3 = g4_bcast(o, 2);
    fe_sub(x3, x3, t);
    fe_sub(x3, x3, t);
    fe_sub(t, t, x3);
    // step 3 (every lane, no exchange): (t - x3) * L
    G4_MUL(o, t, lq);
    const int inf = a.inf;
    fe_sub(r.y, o, ss);
    r.x = x3;
    r.z = u;
    r.inf = inf;
}

// steps 4..6 shared by the two additions: given h, rr, (x1, s1 = "y1 * z2^3") and zz = z1*z2 (or z1)
__device__ __forceinline__ void g4_add_finish(gej& g, int l, const fe& h, const fe& rr, const fe& u1, const fe& s1, const fe& zz) {
    fe p, q, o;
    // step 4: h^2 | rr^2 | zz*h
    p = g4_sel(l, h, rr, zz, zz);
    q = g4_sel(l, h, rr, h, h);
    G4_MUL(o, p, q);
    fe hh = g4_bcast(o, 0), x3 = g4_bcast(o, 1), z3 = g4_bcast(o, 2);
    // step 5: hh*h | u1*hh
    p = g4_sel2(l, hh, u1);
    q = g4_sel2(l, h, hh);
    G4_MUL(o, p, q);
    fe hhh = g4_bcast(o, 0), v = g4_bcast(o, 1);
    fe_sub(x3, x3, hhh);
    fe_sub(x3, x3, v);
    fe_sub(x3, x3, v);
    fe_sub(v, v, x3);
    // step 6: (v - x3)*rr | hhh*s1
    p = g4_sel2(l, v, hhh);
    q = g4_sel2(l, rr, s1);
    G4_MUL(o, p, q);
    fe m = g4_bcast(o, 0), w = g4_bcast(o, 1);
    fe_sub(g.y, m, w);
    g.x = x3;
    g.z = z3;
    g.inf = 0;
}

G4_FN void gej_add_g4(gej& r, const gej& a, const gej& b) {
    const bool triv = a.inf || b.inf;
    if (__all_sync(G4_FULL, triv)) { gej t = a.inf ? b : a; r = t; return; }      // warp-uniform shortcut (sparse buckets)
    const int l = (int)(threadIdx.x & 3u);
    fe p, q, o;
    // step 1: z1^2 | z2^2 | z1*z2
    p = g4_sel(l, a.z, b.z, a.z, a.z);
    q = g4_sel(l, a.z, b.z, b.z, b.z);
    G4_MUL(o, p, q);
    fe z12 = g4_bcast(o, 0), z22 = g4_bcast(o, 1), zz = g4_bcast(o, 2);
    // step 2: x1*z2^2 | x2*z1^2 | y1*z2 | y2*z1
    p = g4_sel(l, a.x, b.x, a.y, b.y);
    q = g4_sel(l, z22, z12, b.z, a.z);
    G4_MUL(o, p, q);
    fe u1 = g4_bcast(o, 0), u2 = g4_bcast(o, 1), t1 = g4_bcast(o, 2), t2 = g4_bcast(o, 3);
    // step 3: t1*z2^2 | t2*z1^2
    p = g4_sel2(l, t1, t2);
    q = g4_sel2(l, z22, z12);
    G4_MUL(o, p, q);
    fe s1 = g4_bcast(o, 0), s2 = g4_bcast(o, 1);
    fe h, rr;
    fe_sub(h, u2, u1);
    fe_sub(rr, s2, s1);
    const bool hz = fe_is_zero(h), rz = fe_is_zero(rr);
    gej g;
    g4_add_finish(g, l, h, rr, u1, s1, zz);
    const bool need_dbl = !triv && hz && rz;
    gej d;
    if (__any_sync(G4_FULL, need_dbl)) gej_double_g4(d, a);                        // branch on a warp vote
    gej out;
    if (a.inf) out = b;
    else if (b.inf) out = a;
    else if (hz) { if (rz) out = d; else gej_set_infinity(out); }
    else out = g;
    r = out;
}

// r = a + b, b affine
G4_FN void gej_add_ge_g4(gej& r, const gej& a, const ge& b) {
    const bool triv = a.inf || b.inf;
    if (__all_sync(G4_FULL, triv)) { gej t; if (a.inf) gej_set_ge(t, b); else t = a; r = t; return; }
    const int l = (int)(threadIdx.x & 3u);
    fe p, q, o, z12;
    G4_MUL(z12, a.z, a.z);                                      // step 1 (every lane)
    // step 2: x2*z1^2 | y2*z1
    p = g4_sel2(l, b.x, b.y);
    q = g4_sel2(l, z12, a.z);
    G4_MUL(o, p, q);
    fe u2 = g4_bcast(o, 0), t2 = g4_bcast(o, 1), s2;
    G4_MUL(s2, t2, z12);                                        // step 3 (every lane)
    fe h, rr;
    fe_sub(h, u2, a.x);
    fe_sub(rr, s2, a.y);
    const bool hz = fe_is_zero(h), rz = fe_is_zero(rr);
    gej g;
    g4_add_finish(g, l, h, rr, a.x, a.y, a.z);
    const bool need_dbl = !triv && hz && rz;
    gej d;
    if (__any_sync(G4_FULL, need_dbl)) gej_double_g4(d, a);
    gej out;
    if (a.inf) gej_set_ge(out, b);
    else if (b.inf) out = a;
    else if (hz) { if (rz) out = d; else gej_set_infinity(out); }
    else out = g;
    r = out;
}
#else
HD void gej_double_g4(gej& r, const gej& a) { gej t; gej_double(t, a); r = t; }
HD void gej_add_g4(gej& r, const gej& a, const gej& b) { gej t; gej_add(t, a, b); r = t; }
HD void gej_add_ge_g4(gej& r, const gej& a, const ge& b) { gej t; gej_add_ge(t, a, b); r = t; }
#endif

}  // namespace B200_NS
