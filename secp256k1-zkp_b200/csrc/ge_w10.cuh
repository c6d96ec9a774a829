// ge_w10.cuh -- a chain of point doublings spread over a WHOLE WARP, one 26-bit limb of every field element per lane.
//
// Why: sum_w 2^(c w) W_w needs ~128 - c doublings one after the other whatever the size of the multi-scalar multiplication; this chain is
// the latency floor of the MSM (k_msm_tail) and the whole cost of a small one.  A warp executing fe_mul issues its 72 IMAD.WIDE one per 4
// cycles whether 1 or 32 lanes are active, and each waits for the carry of the one before: ~0.3 us per multiplication, 2.0 us per doubling
// in the 4-lane form of ge_g4.cuh (3 multiplications deep).  Here a field element is 10 limbs of 26 bits (the reference's 10x26 form,
// field_10x26_impl.h: sums of ten 52-bit products fit 64 bits, additions need no carries), limb k of every element lives in lane k of a
// group of 10 lanes, and a multiplication is: 20 shuffles that bring a_i and b_(k-i) to lane k, 20 independent multiply-adds per lane into the
// two column sums it owns (k and k+10), and a reduction whose carries travel between lanes in four shuffle rounds -- about a third of the
// dependent instructions.  Three groups (lanes 0-9, 10-19, 20-29) work on three different products of the same doubling; lanes 30, 31 shadow
// lanes 20, 21.
//
// Doubling (dbl-2009-l, a = 0; a different Jacobian representative than gej_double's, the same point):
//   step 1   A = X^2 | B = Y^2 | YZ = Y Z            step 2   XB = X B | C = B^2 | A2 = A^2
//   X3 = 9 A2 - 8 XB     N = 12 XB - 9 A2     E = 3 A       step 3   M = E N (all groups)       Y3 = M - 8 C     Z3 = 2 YZ
// Magnitudes: a product's limbs are < 2^26 + 2^21 (top limb < 2^22 + 2^7); w10_mul needs a_i b_j < 2^54.1, i.e. magnitudes 1 x 4 or 2 x 2;
// X3, N, Y3 go through one carry round (w10_norm), Z3 keeps magnitude 2, E magnitude 3.  Differences are taken as a + 16 p - b.
// The bit widths of every intermediate were checked with an exact integer model of the lane program (random and extreme inputs).
//
// Contract: all 32 lanes call together, converged, with identical arguments; all lanes receive the identical result.  Infinity (z = 0)
// stays infinity: every product with Z is exactly zero.  Host build: plain gej_double, n times.
#pragma once
#include "ge.cuh"

namespace B200_NS {

#if defined(__CUDA_ARCH__)
#define W10_FULL 0xffffffffu
#define W10_M26 0x3FFFFFFu
#define W10_M22 0x03FFFFFu

struct w10_lane { int k, base, prev, top, g; uint32_t p16; };     // limb index, first lane of the group, lanes of limb k-1 / limb 9, group, 16 p_k

__device__ __forceinline__ w10_lane w10_lane_init() {
    w10_lane L;
    const int lane = (int)(threadIdx.x & 31u);
    L.g = lane >= 20 ? 2 : lane / 10;
    L.base = 10 * L.g;
    L.k = lane >= 30 ? lane - 30 : lane - L.base;
    L.prev = L.base + (L.k == 0 ? 9 : L.k - 1);
    L.top = L.base + 9;
    L.p16 = 16u * (L.k == 0 ? 0x3FFFC2Fu : L.k == 1 ? 0x3FFFFBFu : L.k == 9 ? 0x03FFFFFu : 0x3FFFFFFu);
    return L;
}

// limb k of a 256-bit value held (identically) by every lane
__device__ __forceinline__ uint32_t w10_from_fe(const fe& a, const w10_lane& L) {
    const int o = 26 * L.k, w = o >> 5;
    uint32_t lo = a.v[0], hi = a.v[1];
#pragma unroll
    for (int j = 1; j < 8; j++) {
        if (w == j) { lo = a.v[j]; hi = j < 7 ? a.v[j < 7 ? j + 1 : 7] : 0u; }
    }
    return __funnelshift_r(lo, hi, (uint32_t)(o & 31)) & W10_M26;
}
// the value whose limbs the lanes of group 0 hold (limbs < 2^32), as 8 words in every lane
__device__ __forceinline__ fe w10_to_fe(uint32_t v) {
    uint32_t l[10];
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 10; j++) {
        c += __shfl_sync(W10_FULL, v, j);
        l[j] = (uint32_t)c & W10_M26;
        c >>= 26;
    }
    uint32_t w[8];
    w[0] = l[0] | (l[1] << 26);
    w[1] = (l[1] >> 6) | (l[2] << 20);
    w[2] = (l[2] >> 12) | (l[3] << 14);
    w[3] = (l[3] >> 18) | (l[4] << 8);
    w[4] = (l[4] >> 24) | (l[5] << 2) | (l[6] << 28);
    w[5] = (l[6] >> 4) | (l[7] << 22);
    w[6] = (l[7] >> 10) | (l[8] << 16);
    w[7] = (l[8] >> 16) | (l[9] << 10);
    uint64_t extra = (uint64_t)(l[9] >> 22) + (c << 4);              // multiples of 2^256 = 2^32 + 977 (mod p); small
#pragma unroll 1
    for (int round = 0; round < 2; round++) {                        // the second round takes the (rare) carry of the first
        uint64_t t = (uint64_t)w[0] + extra * 977u;
        w[0] = (uint32_t)t; t >>= 32;
        t += (uint64_t)w[1] + extra;
        w[1] = (uint32_t)t; t >>= 32;
#pragma unroll
        for (int j = 2; j < 8; j++) { t += w[j]; w[j] = (uint32_t)t; t >>= 32; }
        extra = t;
    }
    fe r;
#pragma unroll
    for (int j = 0; j < 8; j++) r.v[j] = w[j];
    return r;
}

// one carry round: limbs < 2^32 in, limbs < 2^26 + 2^21 out (top limb < 2^22 + 2^7); the bits above 2^256 come back through 2^256 = 2^32 + 977
__device__ __forceinline__ uint32_t w10_norm(uint32_t v, const w10_lane& L) {
    const bool top = L.k == 9;
    const uint32_t l = v & (top ? W10_M22 : W10_M26);
    const uint32_t c = v >> (top ? 22 : 26);
    const uint32_t cp = __shfl_sync(W10_FULL, c, L.prev), c9 = __shfl_sync(W10_FULL, c, L.top);
    return l + (L.k == 0 ? cp * 0x3D1u : cp) + (L.k == 1 ? c9 * 0x40u : 0u);
}

// limb k of a * b mod p for the group's a, b (each lane passes ITS limb of a and of b)
__device__ __forceinline__ uint32_t w10_mul(uint32_t a, uint32_t b, const w10_lane& L) {
    uint64_t lo = 0, hi = 0;                                          // columns k and k + 10
    int j = L.k;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint32_t ai = __shfl_sync(W10_FULL, a, L.base + i);
        const uint32_t bj = __shfl_sync(W10_FULL, b, L.base + j);    // j = (k - i) mod 10
        const bool low = i <= L.k;
        lo += (uint64_t)ai * (low ? bj : 0u);
        hi += (uint64_t)ai * (low ? 0u : bj);
        j = j == 0 ? 9 : j - 1;
    }
    // columns 10..18 -> 26-bit pieces e_k of weight 2^(260 + 26 k); 2^260 = 2^36 + 0x3D10 (mod p), i.e. 0x3D10 at limb k, 0x400 at limb k + 1
    const uint32_t hl = (uint32_t)hi & W10_M26, hh = (uint32_t)(hi >> 26);
    const uint32_t e = hl + __shfl_sync(W10_FULL, hh, L.prev);       // lane 0 receives hh_9 = 0 (column 19 is empty)
    const uint32_t ep = __shfl_sync(W10_FULL, e, L.prev), e9 = __shfl_sync(W10_FULL, e, L.top);
    uint64_t d = lo + (uint64_t)e * 0x3D10u + (uint64_t)(L.k ? ep : 0u) * 0x400u;
    d += (uint64_t)e9 * (L.k == 0 ? 0xF44000u : L.k == 1 ? 0x100000u : 0u);   // 0x400 e_9 sits at limb 10: once more through 2^260
    // carry round A (26 bits everywhere; the carry of limb 9 wraps through 2^260)
    const uint32_t l = (uint32_t)d & W10_M26, h = (uint32_t)(d >> 26);
    const uint32_t hp = __shfl_sync(W10_FULL, h, L.prev), h9 = __shfl_sync(W10_FULL, h, L.top);
    const uint64_t r = (uint64_t)l + (L.k == 0 ? (uint64_t)hp * 0x3D10u : (uint64_t)hp) + (L.k == 1 ? (uint64_t)h9 * 0x400u : 0ull);
    // carry round B (the top limb keeps 22 bits; what is above 2^256 wraps through 2^256 = 2^32 + 977)
    const bool top = L.k == 9;
    const uint32_t l2 = (uint32_t)r & (top ? W10_M22 : W10_M26);
    const uint32_t c = (uint32_t)(r >> (top ? 22 : 26));
    const uint32_t cp = __shfl_sync(W10_FULL, c, L.prev), c9 = __shfl_sync(W10_FULL, c, L.top);
    return l2 + (L.k == 0 ? cp * 0x3D1u : cp) + (L.k == 1 ? c9 * 0x40u : 0u);
}

// A point in lane form: lane k of each group holds limb k of X, Y, Z (all three groups hold the same values); inf is uniform.
// Invariant between operations: X, Y magnitude 1, Z magnitude <= 2.
struct gej_w10 { uint32_t x, y, z; int inf; };

__device__ __forceinline__ void w10_set_gej(gej_w10& P, const gej& a, const w10_lane& L) {
    P.x = w10_from_fe(a.x, L); P.y = w10_from_fe(a.y, L); P.z = w10_from_fe(a.z, L); P.inf = a.inf;
}
__device__ __forceinline__ void w10_set_infinity(gej_w10& P) { P.x = 0; P.y = 0; P.z = 0; P.inf = 1; }
__device__ __forceinline__ void w10_get_gej(gej& r, const gej_w10& P) {
    if (P.inf) { gej_set_infinity(r); return; }                        // uniform
    r.x = w10_to_fe(P.x); r.y = w10_to_fe(P.y); r.z = w10_to_fe(P.z); r.inf = 0;
}
// the three products of one step, each group's result copied to every group
#define W10_MUL3(o0, o1, o2, p, q)                                                   \
    { const uint32_t o_ = w10_mul(p, q, L);                                         \
      o0 = __shfl_sync(W10_FULL, o_, L.k); o1 = __shfl_sync(W10_FULL, o_, 10 + L.k); o2 = __shfl_sync(W10_FULL, o_, 20 + L.k); }
#define W10_SEL(a0, a1, a2) (L.g == 0 ? (a0) : (L.g == 1 ? (a1) : (a2)))

// P = 2^n P
static __device__ __forceinline__ void w10_double_n(gej_w10& P, uint32_t n, const w10_lane& L) {
    if (P.inf) return;                                                  // uniform
    uint32_t X = P.x, Y = P.y, Z = P.z;
#pragma unroll 1
    for (uint32_t it = 0; it < n; it++) {
        uint32_t A, B, YZ, XB, C, A2;
        W10_MUL3(A, B, YZ, W10_SEL(X, Y, Y), W10_SEL(X, Y, Z));         // step 1: X*X | Y*Y | Y*Z
        W10_MUL3(XB, C, A2, W10_SEL(X, B, A), W10_SEL(B, B, A));        // step 2: X*B | B*B | A*A
        X = w10_norm(9u * A2 + L.p16 - 8u * XB, L);
        const uint32_t N = w10_norm(12u * XB + L.p16 - 9u * A2, L);
        const uint32_t M = w10_mul(3u * A, N, L);                       // step 3: (3 A) * N in every group
        Y = w10_norm(M + L.p16 - 8u * C, L);
        Z = 2u * YZ;
    }
    P.x = X; P.y = Y; P.z = Z;
}

// P = P + Q (complete: infinity on either side, Q = P, Q = -P).  add-2007-bl without the doubling shortcuts:
//   step 1  Z2^2 | Z1^2 | Z1 Z2        step 2  U1 = X1 Z2^2 | U2 = X2 Z1^2 | Z2^3        H = U2 - U1
//   step 3  Z1^3 | S1 = Y1 Z2^3 | H^2   step 4  S2 = Y2 Z1^3 | H^3 | V = U1 H^2          R = S2 - S1
//   step 5  R^2 | S1 H^3 | Z3 = Z1 Z2 H      X3 = R^2 - H^3 - 2V      step 6  R (V - X3)      Y3 = R (V - X3) - S1 H^3
static __device__ __forceinline__ void w10_add(gej_w10& P, const gej_w10& Q, const w10_lane& L) {
    if (Q.inf) return;                                                  // uniform branches: all lanes hold the same points
    if (P.inf) { P = Q; return; }
    uint32_t z22, z12, zab, u1, u2, t1, t2, S1, HH, S2, HHH, V, RR, S1H, Z3;
    W10_MUL3(z22, z12, zab, W10_SEL(Q.z, P.z, P.z), W10_SEL(Q.z, P.z, Q.z));
    W10_MUL3(u1, u2, t1, W10_SEL(P.x, Q.x, Q.z), W10_SEL(z22, z12, z22));
    const uint32_t H = w10_norm(u2 + L.p16 - u1, L);
    W10_MUL3(t2, S1, HH, W10_SEL(P.z, P.y, H), W10_SEL(z12, t1, H));
    W10_MUL3(S2, HHH, V, W10_SEL(Q.y, H, u1), W10_SEL(t2, HH, HH));
    const uint32_t R = w10_norm(S2 + L.p16 - S1, L);
    const fe hf = w10_to_fe(H), rf = w10_to_fe(R);                      // exact zero tests on the 8 x 32 form
    if (fe_is_zero(hf)) {                                               // same x: Q = P or Q = -P (uniform)
        if (fe_is_zero(rf)) w10_double_n(P, 1, L); else P.inf = 1;
        return;
    }
    W10_MUL3(RR, S1H, Z3, W10_SEL(R, S1, zab), W10_SEL(R, HHH, H));
    const uint32_t X3 = w10_norm(RR + L.p16 - HHH - 2u * V, L);
    const uint32_t D = w10_norm(V + L.p16 - X3, L);
    const uint32_t M = w10_mul(R, D, L);
    P.x = X3; P.y = w10_norm(M + L.p16 - S1H, L); P.z = Z3;
}

// r = 2^n a on uniform operands (all lanes pass and receive the same point)
static __device__ __forceinline__ void gej_double_n_w10(gej& r, const gej& a, uint32_t n) {
    if (a.inf || n == 0) { r = a; return; }
    const w10_lane L = w10_lane_init();
    gej_w10 P;
    w10_set_gej(P, a, L);
    w10_double_n(P, n, L);
    w10_get_gej(r, P);
}
// r = a + b on uniform operands
static __device__ __forceinline__ void gej_add_w10(gej& r, const gej& a, const gej& b) {
    const w10_lane L = w10_lane_init();
    gej_w10 P, Q;
    w10_set_gej(P, a, L); w10_set_gej(Q, b, L);
    w10_add(P, Q, L);
    if (P.inf) gej_set_infinity(r); else w10_get_gej(r, P);
}

#else   // host build (tests/hostemu): the same names on plain points
struct w10_lane { int unused; };
struct gej_w10 { gej g; int inf; };
inline w10_lane w10_lane_init() { return w10_lane{0}; }
inline void w10_set_gej(gej_w10& P, const gej& a, const w10_lane&) { P.g = a; P.inf = a.inf; }
inline void w10_set_infinity(gej_w10& P) { gej_set_infinity(P.g); P.inf = 1; }
inline void w10_get_gej(gej& r, const gej_w10& P) { r = P.g; }
inline void w10_double_n(gej_w10& P, uint32_t n, const w10_lane&) {
    for (uint32_t i = 0; i < n; i++) { gej u; gej_double(u, P.g); P.g = u; }
    P.inf = P.g.inf;
}
inline void w10_add(gej_w10& P, const gej_w10& Q, const w10_lane&) { gej u; gej_add(u, P.g, Q.g); P.g = u; P.inf = u.inf; }
inline void gej_double_n_w10(gej& r, const gej& a, uint32_t n) {
    gej t = a;
    for (uint32_t i = 0; i < n; i++) { gej u; gej_double(u, t); t = u; }
    r = t;
}
inline void gej_add_w10(gej& r, const gej& a, const gej& b) { gej_add(r, a, b); }
#endif

}  // namespace B200_NS
