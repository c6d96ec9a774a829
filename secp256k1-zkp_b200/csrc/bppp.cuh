// bppp.cuh -- Bulletproofs++ norm-product argument verification, one warp per proof.
//
// Replaces static secp256k1_bppp_rangeproof_norm_product_verify (modules/bppp/bppp_norm_product_impl.h:425-552)
// with its helpers: challenge scalar (bppp_transcript_impl.h:25-33), point-pair codec (bppp_util.h:28-45),
// s_g / s_h vectors (:499-517), v (:518-523) and the two ecmult_multi_var calls (:531, :543).
//
// Re-design: the reference checks  C + sum gamma_i X_i + (gamma_i^2-1) R_i  ==  v G + sum s_g[i] G_i + sum s_h[i] H_i
// with two Strauss MSMs.  Here the equation is moved to one side -- a single sum of 2*rounds + g_len + h_len
// independent a*P terms (+ C, - v*G) must be the point at infinity -- and the terms are dealt to the 32 lanes of a warp
// (exactly one term per lane at the 64-bit geometry: 8 proof points + 16 + 8 generators).  s_g[i] and s_h[i] are computed
// per lane in closed form, s[i] = s[0] * prod_{b in bits(i)} f_b, instead of the reference's sequential fill.
#pragma once
#include "ecmult.cuh"
#include "sha256.cuh"
#include "verify.cuh"

namespace B200_NS {

#define BPPP_MAX_LOG 6       // g_len, h_len <= 64 as in the reference tests (bppp/tests_impl.h:655-663)

struct bppp_ctx {
    uint32_t g_len, h_len, log_g, log_h, n_rounds;
    sc gammas[BPPP_MAX_LOG];
    sc gfac[BPPP_MAX_LOG];   // gamma_b * rho_inv^(2^b)
    sc sg0;                  // n * rho^(2^log_g) * rho^-1
    sc n, l, rho_f;
};

HD uint32_t bppp_log2(uint32_t n) { uint32_t k = 0; while (n >>= 1) k++; return k; }

// shape checks, n/l parsing, rho powers, Fiat-Shamir challenges.  transcript: a SHA-256 object in the state the parent
// protocol left it (reference argument `secp256k1_sha256 *transcript`), passed by value and advanced locally.
HD int bppp_prepare(bppp_ctx& c, const uint8_t* proof, size_t plen, sha256 transcript, const sc& rho, uint32_t g_len, uint32_t h_len) {
    if (g_len == 0 || h_len == 0) return 0;
    if (g_len > 64 || h_len > 64) return 0;
    c.g_len = g_len; c.h_len = h_len;
    c.log_g = bppp_log2(g_len); c.log_h = bppp_log2(h_len);
    c.n_rounds = c.log_g > c.log_h ? c.log_g : c.log_h;
    if (plen != 65 * (size_t)c.n_rounds + 64) return 0;
    if ((g_len & (g_len - 1)) || (h_len & (h_len - 1))) return 0;
    if (sc_set_b32(c.n, proof + 65 * c.n_rounds)) return 0;
    if (sc_set_b32(c.l, proof + 65 * c.n_rounds + 32)) return 0;
    if (sc_is_zero(rho)) return 0;
    sc rho_inv;
    sc_inverse(rho_inv, rho);
    sc rip = rho_inv;                       // rho_inv^(2^b)
    c.rho_f = rho;
    for (uint32_t b = 0; b < c.n_rounds; b++) {
        // gamma_b = H(transcript || X_b R_b pair) with a little-endian 0 index appended (bppp_transcript_impl.h:25-33)
        sha256_write(transcript, proof + 65 * b, 65);
        sha256 t2 = transcript;
        const uint8_t zero8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint8_t hb[32];
        sha256_write(t2, zero8, 8);
        sha256_finalize(t2, hb);
        sc_set_b32(c.gammas[b], hb);
        if (b < c.log_g) {
            sc_mul(c.gfac[b], c.gammas[b], rip);
            sc_sqr(rip, rip);
            sc_sqr(c.rho_f, c.rho_f);
        }
    }
    sc_mul(c.sg0, c.n, c.rho_f);
    sc_mul(c.sg0, c.sg0, rho_inv);
    return 1;
}

// s_g[i] (:499-510) and s_h[i] (:511-517) in closed form
HD void bppp_sg(sc& r, const bppp_ctx& c, uint32_t i) {
    r = c.sg0;
    for (uint32_t b = 0; b < c.log_g; b++) if ((i >> b) & 1u) sc_mul(r, r, c.gfac[b]);
}
HD void bppp_sh(sc& r, const bppp_ctx& c, uint32_t i) {
    r = c.l;
    for (uint32_t b = 0; b < c.log_h; b++) if ((i >> b) & 1u) sc_mul(r, r, c.gammas[b]);
}
// v = n^2 * rho_f^2 + <c_vec, s_h>   (:518-523)
HD void bppp_v(sc& v, const bppp_ctx& c, const uint8_t* c_vec32) {
    sc hc; sc_set_zero(hc);
    for (uint32_t i = 0; i < c.h_len; i++) {
        sc ci, sh, t;
        sc_set_b32(ci, c_vec32 + 32 * i);
        bppp_sh(sh, c, i);
        sc_mul(t, ci, sh);
        sc_add(hc, hc, t);
    }
    sc mu_f;
    sc_sqr(mu_f, c.rho_f);
    sc_mul(v, c.n, c.n);
    sc_mul(v, v, mu_f);
    sc_add(v, v, hc);
}

// one point of a 65-byte pair (bppp_util.h:28-45): returns 0 on a malformed encoding; infinity allowed (all-zero x, sign 0)
HD int bppp_parse_one_of_points(ge& pt, const uint8_t* in65, int idx) {
    if (in65[0] > 3) return 0;
    const uint8_t* x = in65 + 1 + 32 * idx;
    uint8_t z = 0;
    for (int k = 0; k < 32; k++) z |= x[k];
    uint8_t sign = (uint8_t)((in65[0] & (2 - idx)) >> (1 - idx));
    if (z == 0) {
        if (sign) return 0;
        ge_set_infinity(pt);
        return 1;
    }
    fe fx;
    if (!fe_set_b32(fx, x)) return 0;
    return ge_set_xo(pt, fx, sign);
}

HD uint32_t bppp_num_terms(const bppp_ctx& c) { return 2 * c.n_rounds + c.g_len + c.h_len; }

// term t of the single-sided equation: point (affine, may be infinity) and its scalar.
//   t <  rounds            : X_t,  gamma_t
//   t <  2*rounds          : R_i,  gamma_i^2 - 1
//   then generators        : G_i, -s_g[i]   /  H_i, -s_h[i]
HD int bppp_term(ge& pt, sc& s, const bppp_ctx& c, uint32_t t, const uint8_t* proof, const uint8_t* gens64) {
    if (t < c.n_rounds) {
        s = c.gammas[t];
        return bppp_parse_one_of_points(pt, proof + 65 * t, 0);
    }
    t -= c.n_rounds;
    if (t < c.n_rounds) {
        sc one; sc_set_int(one, 1); sc_negate(one, one);
        sc_sqr(s, c.gammas[t]);
        sc_add(s, s, one);
        return bppp_parse_one_of_points(pt, proof + 65 * t, 1);
    }
    t -= c.n_rounds;
    if (t < c.g_len) bppp_sg(s, c, t); else bppp_sh(s, c, t - c.g_len);
    sc_negate(s, s);
    fe_set_b32(pt.x, gens64 + 64 * t);
    fe_set_b32(pt.y, gens64 + 64 * t + 32);
    pt.inf = 0;
    return 1;
}

}  // namespace B200_NS
