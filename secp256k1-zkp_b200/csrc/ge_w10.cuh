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

// r = 2^n a
static __device__ __forceinline__ void gej_double_n_w10(gej& r, const gej& a, uint32_t n) {
    if (a.inf || n == 0) { r = a; return; }                            // uniform
    const w10_lane L = w10_lane_init();
    const int r0 = L.k, r1 = 10 + L.k, r2 = 20 + L.k;                 // the lanes that hold limb k of group 0 / 1 / 2's product
    uint32_t X = w10_from_fe(a.x, L), Y = w10_from_fe(a.y, L), Z = w10_from_fe(a.z, L);
#pragma unroll 1
    for (uint32_t it = 0; it < n; it++) {
        // step 1: X*X | Y*Y | Y*Z
        uint32_t o = w10_mul(L.g == 0 ? X : Y, L.g == 0 ? X : (L.g == 1 ? Y : Z), L);
        const uint32_t A = __shfl_sync(W10_FULL, o, r0), B = __shfl_sync(W10_FULL, o, r1), YZ = __shfl_sync(W10_FULL, o, r2);
        // step 2: X*B | B*B | A*A
        o = w10_mul(L.g == 0 ? X : (L.g == 1 ? B : A), L.g == 2 ? A : B, L);
        const uint32_t XB = __shfl_sync(W10_FULL, o, r0), C = __shfl_sync(W10_FULL, o, r1), A2 = __shfl_sync(W10_FULL, o, r2);
        X = w10_norm(9u * A2 + L.p16 - 8u * XB, L);
        const uint32_t N = w10_norm(12u * XB + L.p16 - 9u * A2, L);
        // step 3: (3 A) * N in every group
        o = w10_mul(3u * A, N, L);
        Y = w10_norm(o + L.p16 - 8u * C, L);
        Z = 2u * YZ;
    }
    r.x = w10_to_fe(X); r.y = w10_to_fe(Y); r.z = w10_to_fe(Z);
    r.inf = 0;
}

#else   // host build: the plain function, n times
inline void gej_double_n_w10(gej& r, const gej& a, uint32_t n) {
    gej t = a;
    for (uint32_t i = 0; i < n; i++) { gej u; gej_double(u, t); t = u; }
    r = t;
}
#endif

}  // namespace B200_NS
