// codec.cuh -- per-item point (de)serialization helpers that need field arithmetic (square roots), so that
// the single-call drop-in API never computes on the host:
//   secp256k1_ec_pubkey_parse (secp256k1.c:290-306, eckey_impl.h:18-36), secp256k1_xonly_pubkey_parse
//   (extrakeys/main_impl.h:22-42), secp256k1_pedersen_commitment_parse (generator/main_impl.h:281-297),
//   secp256k1_generator_parse / _serialize (generator/main_impl.h:59-92).
#pragma once
#include "verify.cuh"
#include "rangeproof.cuh"

namespace B200_NS {

enum codec_fmt { CODEC_XONLY32 = 1, CODEC_EC_PUBKEY = 2, CODEC_COMMIT33 = 3, CODEC_GENERATOR33 = 4, CODEC_GENERATOR_SER = 5 };

// returns ok; out64 layout depends on fmt: pubkeys -> 64-byte opaque (x,y little-endian); commitment/generator -> x||y big-endian;
// CODEC_GENERATOR_SER: in = 64-byte generator, out = 33-byte serialization (in the first 33 bytes of out64)
HD int codec_item(uint8_t* out64, const uint8_t* in, uint32_t in_len, int fmt) {
    ge p;
    int ok = 0;
    for (int k = 0; k < 64; k++) out64[k] = 0;
    if (fmt == CODEC_XONLY32) {
        ok = pubkey_load(p, in, PK_XONLY32);
        if (ok) ge_to_opaque64(out64, p);
    } else if (fmt == CODEC_EC_PUBKEY) {
        if (in_len == 33) {
            ok = pubkey_load(p, in, PK_COMPRESSED33);
        } else if (in_len == 65 && (in[0] == 4 || in[0] == 6 || in[0] == 7)) {
            fe x, y;
            if (fe_set_b32(x, in + 1) && fe_set_b32(y, in + 33)) {
                p.x = x; p.y = y; p.inf = 0;
                ok = 1;
                if ((in[0] == 6 || in[0] == 7) && fe_is_odd(y) != (in[0] == 7)) ok = 0;
                if (ok) ok = ge_is_valid(p);
            }
        }
        if (ok) ge_to_opaque64(out64, p);
    } else if (fmt == CODEC_COMMIT33) {
        ok = rp_commitment_load(p, in);
        if (ok) { fe_get_b32(out64, p.x); fe_get_b32(out64 + 32, p.y); }
    } else if (fmt == CODEC_GENERATOR33) {
        fe x;
        if ((in[0] & 0xFE) == 10 && fe_set_b32(x, in + 1) && ge_set_xquad(p, x)) {
            if (in[0] & 1) fe_neg(p.y, p.y);
            fe_get_b32(out64, p.x); fe_get_b32(out64 + 32, p.y);
            ok = 1;
        }
    } else if (fmt == CODEC_GENERATOR_SER) {
        rp_generator_load(p, in);
        out64[0] = (uint8_t)(11 ^ fe_is_square(p.y));
        fe_get_b32(out64 + 1, p.x);
        ok = 1;
    }
    return ok;
}

}  // namespace B200_NS
