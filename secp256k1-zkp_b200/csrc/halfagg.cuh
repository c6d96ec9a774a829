// halfagg.cuh -- BIP-340 half-aggregation verification (secp256k1_schnorrsig_aggverify,
// src/modules/schnorrsig_halfagg/main_impl.h:108-196), re-cut for a GPU:
//
//   reference (serial, per signature):  T_i = R_i + e_i*P_i ; rhs += z_i*T_i     (two variable-base ecmults per signature)
//   here:                               sum_i z_i*R_i + (z_i*e_i)*P_i - s*G == infinity   as ONE Pippenger MSM over 2n+1 terms
//
// Pieces: (1) the randomizer chain z_i (a running SHA-256 over r_0||pk_0||m_0 .. r_i||pk_i||m_i: inherently serial, :144-155),
// (2) one thread per signature: lift R_i and P_i, challenge e_i, scalars z_i and z_i*e_i, (3) the MSM of msm.cuh with
// generator scalar -s, (4) accept iff nothing failed to parse and the sum is the point at infinity.
#pragma once
#include "verify.cuh"

namespace B200_NS {

struct halfagg_midstate { uint32_t s[8]; };   // SHA256 state after SHA256(tag)||SHA256(tag), tag = "HalfAgg/randomizer" (:9-17)

// z32[i] = 32 raw hash bytes (reduced mod n by the consumer); z_0 = 1 by definition (:187 "if i == 0, zi = 1 implicitly").
// The hash still absorbs triple 0.  One caller walks the whole chain.
HD void halfagg_z_chain(uint8_t* z32, const halfagg_midstate& mid, const uint8_t* aggsig, const uint8_t* pk32, const uint8_t* msgs32, size_t n) {
    sha256 h;
    for (int k = 0; k < 8; k++) h.s[k] = mid.s[k];
    h.bytes = 64;
    for (size_t i = 0; i < n; i++) {
        sha256_write(h, aggsig + 32 * i, 32);
        sha256_write(h, pk32 + 32 * i, 32);      // = fe_get_b32(pp.x) of the parsed key; an unparsable key fails the item in halfagg_term
        sha256_write(h, msgs32 + 32 * i, 32);
        if (i == 0) {
            for (int k = 0; k < 32; k++) z32[k] = (k == 31);
        } else {
            sha256 c = h;                        // hashcopy (:150-153)
            sha256_finalize(c, z32 + 32 * i);
        }
    }
}

// One signature: MSM inputs (2 scalars, 2 affine points, big-endian).  Returns 0 (and emits neutral terms) if r_i >= p, R_i or P_i
// is not on the curve (:160-166, and the caller-side secp256k1_xonly_pubkey_parse of the reference API).
HD int halfagg_term(uint8_t* sc_out64, uint8_t* pt_out128, const uint8_t* r32, const uint8_t* pk32, const uint8_t* msg32, const uint8_t* z32,
                    const bip340_midstate& mid) {
    fe rx;
    ge R, P;
    int ok = fe_set_b32(rx, r32) && ge_set_xo(R, rx, 0);            // fe_set_b32_limit + ge_set_xo_var(.., 0)
    ok = ok && pubkey_load(P, pk32, PK_XONLY32);
    if (!ok) {
        for (int k = 0; k < 64; k++) sc_out64[k] = 0;
        for (int k = 0; k < 128; k++) pt_out128[k] = 0;
        return 0;
    }
    sc e, z, ze;
    {
        sha256 h;
#pragma unroll
        for (int i = 0; i < 8; i++) h.s[i] = mid.s[i];
        h.bytes = 64;
        uint8_t eb[32];
        sha256_write(h, r32, 32);
        sha256_write(h, pk32, 32);                                   // pk32 is canonical here (x < p): identical to fe_get_b32(P.x)
        sha256_write(h, msg32, 32);
        sha256_finalize(h, eb);
        sc_set_b32(e, eb);                                           // secp256k1_schnorrsig_challenge (:169)
    }
    sc_set_b32(z, z32);
    sc_mul(ze, z, e);
    sc_get_b32(sc_out64, z);
    sc_get_b32(sc_out64 + 32, ze);
    fe_get_b32(pt_out128, R.x); fe_get_b32(pt_out128 + 32, R.y);
    fe_get_b32(pt_out128 + 64, P.x); fe_get_b32(pt_out128 + 96, P.y);
    return 1;
}

// generator scalar of the MSM: -s mod n; returns 0 if s >= n (:183-186)
HD int halfagg_neg_s(uint8_t* out32, const uint8_t* s32) {
    sc s;
    int overflow = sc_set_b32(s, s32);
    sc_negate(s, s);
    sc_get_b32(out32, s);
    return !overflow;
}

// whole verification on one thread (tests/hostemu and tiny inputs): same pieces, plain double-and-add instead of the MSM
template <class Tab>
HD int halfagg_verify_serial(const halfagg_midstate& hmid, const bip340_midstate& mid, const uint8_t* pk32, const uint8_t* msgs32, size_t n,
                             const uint8_t* aggsig, size_t aggsig_len, uint8_t* z_scratch /* 32*n bytes */, const ecmult_gtable& gt, Tab& tab) {
    if (aggsig_len / 32 == 0 || aggsig_len / 32 - 1 != n || aggsig_len % 32 != 0) return 0;     // (:124-126)
    halfagg_z_chain(z_scratch, hmid, aggsig, pk32, msgs32, n);
    gej acc; gej_set_infinity(acc);
    for (size_t i = 0; i < n; i++) {
        uint8_t scb[64], ptb[128];
        if (!halfagg_term(scb, ptb, aggsig + 32 * i, pk32 + 32 * i, msgs32 + 32 * i, z_scratch + 32 * i, mid)) return 0;
        for (int t = 0; t < 2; t++) {
            ge A; fe_set_b32(A.x, ptb + 64 * t); fe_set_b32(A.y, ptb + 64 * t + 32); A.inf = 0;
            sc k, zero; sc_set_b32(k, scb + 32 * t); sc_set_zero(zero);
            gej Aj, Rj, S; gej_set_ge(Aj, A);
            ecmult(Rj, Aj, k, zero, gt, tab);
            gej_add(S, acc, Rj); acc = S;
        }
    }
    uint8_t ns[32];
    if (!halfagg_neg_s(ns, aggsig + 32 * n)) return 0;
    sc k, zero; sc_set_b32(k, ns); sc_set_zero(zero);
    gej inf, Rj, S; gej_set_infinity(inf);
    ecmult(Rj, inf, zero, k, gt, tab);
    gej_add(S, acc, Rj);
    return S.inf;
}

}  // namespace B200_NS
