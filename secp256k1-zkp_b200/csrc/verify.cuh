// verify.cuh -- per-item verification routines (one GPU thread each).
//   schnorr_verify_item : secp256k1_schnorrsig_verify (modules/schnorrsig/main_impl.h:215-261)
//                         + secp256k1_xonly_pubkey_parse (modules/extrakeys/main_impl.h:22-42)
//   ecdsa_verify_item   : secp256k1_ecdsa_verify (secp256k1.c:498-512) -> secp256k1_ecdsa_sig_verify
//                         (ecdsa_impl.h:195-272) + secp256k1_ec_pubkey_parse (eckey_impl.h:18-36)
#pragma once
#include "ecmult.cuh"
#include "sha256.cuh"

namespace B200_NS {

enum pk_format { PK_OPAQUE64 = 0, PK_XONLY32 = 1, PK_COMPRESSED33 = 2 };
enum sig_format { SIG_OPAQUE64 = 0, SIG_COMPACT64 = 1 };

// 64-byte opaque pubkey = ge_storage: x then y, each a 256-bit little-endian integer
// (group_impl.h:999-1017 "ge_to_bytes" is a memcpy of ge_storage on this platform)
HD void ge_from_opaque64(ge& r, const uint8_t* b) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.x.v[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
        r.y.v[i] = (uint32_t)b[32 + 4 * i] | ((uint32_t)b[32 + 4 * i + 1] << 8) | ((uint32_t)b[32 + 4 * i + 2] << 16) | ((uint32_t)b[32 + 4 * i + 3] << 24);
    }
    r.inf = 0;
}
HD void ge_to_opaque64(uint8_t* b, const ge& a) {
    fe x = a.x, y = a.y;
    fe_normalize(x); fe_normalize(y);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        b[4 * i] = (uint8_t)x.v[i]; b[4 * i + 1] = (uint8_t)(x.v[i] >> 8); b[4 * i + 2] = (uint8_t)(x.v[i] >> 16); b[4 * i + 3] = (uint8_t)(x.v[i] >> 24);
        b[32 + 4 * i] = (uint8_t)y.v[i]; b[32 + 4 * i + 1] = (uint8_t)(y.v[i] >> 8); b[32 + 4 * i + 2] = (uint8_t)(y.v[i] >> 16); b[32 + 4 * i + 3] = (uint8_t)(y.v[i] >> 24);
    }
}
HD void sc_from_le32(sc& r, const uint8_t* b) {
#pragma unroll
    for (int i = 0; i < 8; i++)
        r.v[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
}

// parse a public key in one of the accepted formats; returns 0 if invalid
HD int pubkey_load(ge& p, const uint8_t* pk, int fmt) {
    if (fmt == PK_OPAQUE64) {
        ge_from_opaque64(p, pk);
        // secp256k1_pubkey_load ARG_CHECKs x != 0 (secp256k1.c:282): an all-zero (unparsed) key is invalid.  The reference trusts the
        // rest of the object (it can only come from its own parser); a batch API cannot: an off-curve point such as (x, 0) would double
        // to Z = 0 and poison the inversion its thread block shares.  Off-curve opaque keys are therefore rejected (3 field operations).
        return !fe_is_zero(p.x) && ge_is_valid(p);
    } else if (fmt == PK_XONLY32) {
        fe x;
        if (!fe_set_b32(x, pk)) return 0;
        return ge_set_xo(p, x, 0);
    } else {
        if (pk[0] != 0x02 && pk[0] != 0x03) return 0;
        fe x;
        if (!fe_set_b32(x, pk + 1)) return 0;
        return ge_set_xo(p, x, pk[0] == 0x03);
    }
}

struct bip340_midstate { uint32_t s[8]; };   // SHA256 state after SHA256(tag)||SHA256(tag), tag = "BIP0340/challenge"

// BIP-340 verification, split so that the kernel can share ONE field inversion per thread block between all its
// signatures (Montgomery's trick as a shared-memory tree, see fe_block_inverse in k_verify.cu):
//   prepare : parse, challenge hash, R = s*G - e*P (Jacobian).  returns 0 if the signature is already rejected.
//   finish  : given 1/R.z, check y(R) even and x(R) == r.
template <class Tab>
HD int schnorr_verify_prepare(gej& R, fe& rx, const uint8_t* sig64, const uint8_t* msg, uint32_t msglen, const uint8_t* pk, int pkfmt,
                              const bip340_midstate& mid, const ecmult_gtable& gt, Tab& tab) {
    sc s, e;
    ge P;
    if (!fe_set_b32(rx, sig64)) return 0;            // r >= p           (main_impl.h:231)
    if (sc_set_b32(s, sig64 + 32)) return 0;         // s >= n           (:235-238)
    if (!pubkey_load(P, pk, pkfmt)) return 0;
    // e = H_tag(r || pk.x || msg) mod n               (:106-120)
    {
        sha256 h;
#pragma unroll
        for (int i = 0; i < 8; i++) h.s[i] = mid.s[i];
        h.bytes = 64;
        uint8_t px[32], eb[32];
        fe_get_b32(px, P.x);
        sha256_write(h, sig64, 32);
        sha256_write(h, px, 32);
        sha256_write(h, msg, msglen);
        sha256_finalize(h, eb);
        sc_set_b32(e, eb);
    }
    sc_negate(e, e);
    gej Pj;
    gej_set_ge(Pj, P);
    ecmult(R, Pj, e, s, gt, tab);                    // R = s*G - e*P    (:249-251)
    return !R.inf;                                   // (:254)
}
HD int schnorr_verify_finish(const gej& R, const fe& rx, const fe& zinv) {
    ge Ra;
    ge_set_gej_zinv(Ra, R, zinv);
    if (fe_is_odd(Ra.y)) return 0;                   // (:258-259)
    return fe_equal(rx, Ra.x);
}
template <class Tab>
HD int schnorr_verify_item(const uint8_t* sig64, const uint8_t* msg, uint32_t msglen, const uint8_t* pk, int pkfmt,
                           const bip340_midstate& mid, const ecmult_gtable& gt, Tab& tab) {
    gej R; fe rx, zi;
    if (!schnorr_verify_prepare(R, rx, sig64, msg, msglen, pk, pkfmt, mid, gt, tab)) return 0;
    fe_inv(zi, R.z);
    return schnorr_verify_finish(R, rx, zi);
}

// ECDSA verification split around the scalar inversion so that a thread block can share ONE inversion mod n
// (sc_block_inverse in k_verify.cu): parse -> [s^-1] -> finish.
struct ecdsa_state { sc r, s, m; ge P; };
HD int ecdsa_verify_parse(ecdsa_state& st, const uint8_t* sig, int sigfmt, const uint8_t* msghash32, const uint8_t* pk, int pkfmt) {
    if (sigfmt == SIG_COMPACT64) {
        // secp256k1_ecdsa_signature_parse_compact: overflow in r or s => parse failure (secp256k1.c:372-391)
        if (sc_set_b32(st.r, sig)) return 0;
        if (sc_set_b32(st.s, sig + 32)) return 0;
    } else {
        sc_from_le32(st.r, sig);
        sc_from_le32(st.s, sig + 32);
    }
    sc_set_b32(st.m, msghash32);                     // reduced silently (secp256k1.c:507)
    if (sc_is_high(st.s)) return 0;                  // (:509)
    if (!pubkey_load(st.P, pk, pkfmt)) return 0;
    if (sc_is_zero(st.r) || sc_is_zero(st.s)) return 0;    // (ecdsa_impl.h:205-207)
    return 1;
}
template <class Tab>
HD int ecdsa_verify_finish(const ecdsa_state& st, const sc& sn /* s^-1 */, const ecmult_gtable& gt, Tab& tab) {
    sc u1, u2;
    sc_mul(u1, sn, st.m);
    sc_mul(u2, sn, st.r);
    gej Pj, R;
    gej_set_ge(Pj, st.P);
    ecmult(R, Pj, u2, u1, gt, tab);
    if (R.inf) return 0;
    // inversion-free comparison x(R) mod n == r        (ecdsa_impl.h:231-270)
    fe xr;
#pragma unroll
    for (int i = 0; i < 8; i++) xr.v[i] = st.r.v[i];   // r < n < p
    if (gej_eq_x(xr, R)) return 1;
    // second chance only if r + n < p, i.e. r < p - n
    const uint32_t pmn[8] = {0x2FC9BAEEu, 0x402DA172u, 0x50B75FC4u, 0x45512319u, 1u, 0u, 0u, 0u};  // p - n
    if (bn_gte8(st.r.v, pmn)) return 0;
    const uint32_t n[8] = SC_N_LIMBS;
    bn_add8(xr.v, st.r.v, n);
    return gej_eq_x(xr, R);
}
// ECDSA public-key recovery (secp256k1_ecdsa_sig_recover, src/modules/recovery/main_impl.h:87-121), split around r^-1 mod n like
// the verification above.  recid bit 0 = parity of R.y, bit 1 = "R.x is r + n".
struct ecdsa_recover_state { sc r, s, m; ge X; };
HD int ecdsa_recover_parse(ecdsa_recover_state& st, const uint8_t* sig64, int recid, const uint8_t* msghash32) {
    // secp256k1_ecdsa_recoverable_signature_parse_compact: overflow in r or s => parse failure (:38-59); recid is range-checked by the caller
    if (sc_set_b32(st.r, sig64)) return 0;
    if (sc_set_b32(st.s, sig64 + 32)) return 0;
    sc_set_b32(st.m, msghash32);
    if (sc_is_zero(st.r) || sc_is_zero(st.s)) return 0;                    // (:96-98)
    fe fx;
#pragma unroll
    for (int i = 0; i < 8; i++) fx.v[i] = st.r.v[i];                       // r < n < p
    if (recid & 2) {
        const uint32_t pmn[8] = {0x2FC9BAEEu, 0x402DA172u, 0x50B75FC4u, 0x45512319u, 1u, 0u, 0u, 0u};  // p - n
        if (bn_gte8(st.r.v, pmn)) return 0;                                // (:104-107)
        const uint32_t n[8] = SC_N_LIMBS;
        bn_add8(fx.v, st.r.v, n);
    }
    return ge_set_xo(st.X, fx, recid & 1);                                 // (:110-112)
}
template <class Tab>
HD int ecdsa_recover_finish(gej& Q, const ecdsa_recover_state& st, const sc& rn /* r^-1 */, const ecmult_gtable& gt, Tab& tab) {
    sc u1, u2;
    sc_mul(u1, rn, st.m);
    sc_negate(u1, u1);
    sc_mul(u2, rn, st.s);
    gej Xj;
    gej_set_ge(Xj, st.X);
    ecmult(Q, Xj, u2, u1, gt, tab);                                        // Q = u2*X + u1*G   (:117)
    return !Q.inf;
}
template <class Tab>
HD int ecdsa_recover_item(uint8_t* out64, const uint8_t* sig64, int recid, const uint8_t* msghash32, const ecmult_gtable& gt, Tab& tab) {
    ecdsa_recover_state st;
    gej Q;
    int ok = ecdsa_recover_parse(st, sig64, recid, msghash32);
    if (ok) { sc rn; sc_inverse(rn, st.r); ok = ecdsa_recover_finish(Q, st, rn, gt, tab); }
    if (!ok) { for (int k = 0; k < 64; k++) out64[k] = 0; return 0; }
    fe zi; fe_inv(zi, Q.z);
    ge Qa; ge_set_gej_zinv(Qa, Q, zi);
    fe_get_b32(out64, Qa.x); fe_get_b32(out64 + 32, Qa.y);
    return 1;
}

template <class Tab>
HD int ecdsa_verify_item(const uint8_t* sig, int sigfmt, const uint8_t* msghash32, const uint8_t* pk, int pkfmt,
                         const ecmult_gtable& gt, Tab& tab) {
    ecdsa_state st;
    if (!ecdsa_verify_parse(st, sig, sigfmt, msghash32, pk, pkfmt)) return 0;
    sc sn;
    sc_inverse(sn, st.s);
    return ecdsa_verify_finish(st, sn, gt, tab);
}

}  // namespace B200_NS
