// gtable.cuh -- construction of the generator comb table T[j][d] = d * 2^(gw*j) * G.
// Replaces the reference's link-time tables secp256k1_pre_g / secp256k1_pre_g_128
// (precomputed_ecmult.c:16,:8236; generator precompute_ecmult.c:36-44): a different table shape (no
// doublings are shared with the variable-point part on the GPU), built at engine init on the device.
#pragma once
#include "ecmult.cuh"

namespace B200_NS {

HD void ge_set_generator(ge& g) {
    const uint32_t gx[8] = {0x16F81798u, 0x59F2815Bu, 0x2DCE28D9u, 0x029BFCDBu, 0xCE870B07u, 0x55A06295u, 0xF9DCBBACu, 0x79BE667Eu};
    const uint32_t gy[8] = {0xFB10D4B8u, 0x9C47D08Fu, 0xA6855419u, 0xFD17B448u, 0x0E1108A8u, 0x5DA4FBFCu, 0x26A3C465u, 0x483ADA77u};
#pragma unroll
    for (int i = 0; i < 8; i++) { g.x.v[i] = gx[i]; g.y.v[i] = gy[i]; }
    g.inf = 0;
}

HD void ge_to_storage(uint32_t* out16, const ge& a) {
    fe x = a.x, y = a.y;
    fe_normalize(x); fe_normalize(y);
#pragma unroll
    for (int k = 0; k < 8; k++) { out16[k] = x.v[k]; out16[8 + k] = y.v[k]; }
}
HD void ge_from_storage(ge& r, const uint32_t* in16) {
#pragma unroll
    for (int k = 0; k < 8; k++) { r.x.v[k] = in16[k]; r.y.v[k] = in16[8 + k]; }
    r.inf = 0;
}

// bases[j] = 2^(gw*j) * G as affine storage (serial: 256 doublings + `windows` inversions)
HD void gtable_compute_bases(uint32_t* bases /*[windows][16]*/, int gw, int windows) {
    ge g; ge_set_generator(g);
    gej a; gej_set_ge(a, g);
    for (int j = 0; j < windows; j++) {
        ge b; ge_set_gej(b, a);
        ge_to_storage(bases + 16 * j, b);
        for (int k = 0; k < gw; k++) { gej t; gej_double(t, a); a = t; }
    }
}
// out = d * base (d in [1, 2^gw)) as affine storage
HD void gtable_compute_entry(uint32_t* out16, const uint32_t* base16, uint32_t d, int gw) {
    ge b; ge_from_storage(b, base16);
    gej acc; gej_set_infinity(acc);
#pragma unroll 1
    for (int k = gw - 1; k >= 0; k--) {
        gej t;
        if (!acc.inf) { gej_double(t, acc); acc = t; }
        if ((d >> k) & 1u) { gej_add_ge(t, acc, b); acc = t; }
    }
    ge r; ge_set_gej(r, acc);
    ge_to_storage(out16, r);
}

}  // namespace B200_NS
