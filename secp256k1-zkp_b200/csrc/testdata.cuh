// testdata.cuh -- SYNTHETIC INPUT GENERATORS for tests and benchmarks (not constant time, never to be
// used with real secrets).  They exist so that a 2^20-item batch of VALID inputs can be produced on the
// GPU box in milliseconds instead of minutes of host signing.
//   schnorr_sign_item: BIP-340 signing with the default nonce function and aux_rand = NULL, byte-for-byte
//   what secp256k1_schnorrsig_sign_custom(ctx, sig, msg, 32, keypair, NULL) emits
//   (modules/schnorrsig/main_impl.h:38-87 nonce, :125-190 signing; bench_impl.h:60-76 uses the same call).
#pragma once
#include "verify.cuh"

namespace B200_NS {

struct bip340_sign_consts {
    uint32_t mid_challenge[8];   // "BIP0340/challenge"
    uint32_t mid_nonce[8];       // "BIP0340/nonce"
    uint8_t zero_mask[32];       // TaggedHash("BIP0340/aux", 0^32)
};

HD void sha256_from_mid(sha256& h, const uint32_t* mid) {
#pragma unroll
    for (int i = 0; i < 8; i++) h.s[i] = mid[i];
    h.bytes = 64;
}

// returns 0 if sk is 0 or >= n (then sig/pk are zeroed)
HD int schnorr_sign_item(uint8_t* sig64, uint8_t* pk32, const uint8_t* sk32, const uint8_t* msg, uint32_t msglen,
                         const bip340_sign_consts& c, const ecmult_gtable& gt) {
    sc d;
    int ov = sc_set_b32(d, sk32);
    if (ov || sc_is_zero(d)) {
        for (int i = 0; i < 64; i++) sig64[i] = 0;
        for (int i = 0; i < 32; i++) pk32[i] = 0;
        return 0;
    }
    gej Pj; gej_set_infinity(Pj);
    ecmult_gen_accumulate(Pj, d, gt);
    ge P; ge_set_gej(P, Pj);
    if (fe_is_odd(P.y)) sc_negate(d, d);
    uint8_t px[32], dk[32], t[32], rnd[32];
    fe_get_b32(px, P.x);
    sc_get_b32(dk, d);
    for (int i = 0; i < 32; i++) t[i] = dk[i] ^ c.zero_mask[i];
    sha256 h;
    sha256_from_mid(h, c.mid_nonce);
    sha256_write(h, t, 32); sha256_write(h, px, 32); sha256_write(h, msg, msglen);
    sha256_finalize(h, rnd);
    sc k; sc_set_b32(k, rnd);
    if (sc_is_zero(k)) sc_set_int(k, 1);
    gej Rj; gej_set_infinity(Rj);
    ecmult_gen_accumulate(Rj, k, gt);
    ge R; ge_set_gej(R, Rj);
    if (fe_is_odd(R.y)) sc_negate(k, k);
    uint8_t rx[32], eb[32];
    fe_get_b32(rx, R.x);
    sha256_from_mid(h, c.mid_challenge);
    sha256_write(h, rx, 32); sha256_write(h, px, 32); sha256_write(h, msg, msglen);
    sha256_finalize(h, eb);
    sc e, s; sc_set_b32(e, eb);
    sc_mul(s, e, d);
    sc_add(s, s, k);
    for (int i = 0; i < 32; i++) { sig64[i] = rx[i]; pk32[i] = px[i]; }
    sc_get_b32(sig64 + 32, s);
    return 1;
}

}  // namespace B200_NS
