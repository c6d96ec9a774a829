// rewind.cuh -- range-proof rewinding (secp256k1_rangeproof_rewind, modules/rangeproof/main_impl.h:31-52): after a proof has
// verified, the holder of the nonce replays the prover's deterministic randomness and recovers value, blinding factor and the
// embedded message.  Follows secp256k1_rangeproof_genrand (rangeproof_impl.h:61-112), secp256k1_rangeproof_rewind_inner (:354-471),
// the tail of secp256k1_rangeproof_verify_impl (:653-681) and the RFC 6979 HMAC-SHA256 generator (hash_impl.h:211-314).
// One GPU thread per proof: the generator is a serial hash chain (about 2300 compressions for a 64-bit proof); the ring walk that
// produced the challenges e_{i,j} ran before, one warp per proof (rangeproof.cuh), and left them in global memory.
#pragma once
#include "rangeproof.cuh"

namespace B200_NS {

// ---- HMAC-SHA256 with a 32-byte key (hash_impl.h:211-249; longer keys never occur on this path) ---------------------------
struct hmac_sha256 { sha256 inner, outer; };
HD void hmac_init32(hmac_sha256& h, const uint8_t* key32) {
    uint8_t pad[64];
    for (int i = 0; i < 64; i++) pad[i] = (uint8_t)((i < 32 ? key32[i] : 0) ^ 0x5c);
    sha256_init(h.outer); sha256_write(h.outer, pad, 64);
    for (int i = 0; i < 64; i++) pad[i] ^= 0x5c ^ 0x36;
    sha256_init(h.inner); sha256_write(h.inner, pad, 64);
}
HD void hmac_write(hmac_sha256& h, const uint8_t* d, size_t n) { sha256_write(h.inner, d, n); }
HD void hmac_final(hmac_sha256& h, uint8_t* out32) {
    uint8_t t[32];
    sha256_finalize(h.inner, t);
    sha256_write(h.outer, t, 32);
    sha256_finalize(h.outer, out32);
}

// ---- RFC 6979 generator (hash_impl.h:255-314) -------------------------------------------------------------------------------
struct rfc6979 { uint8_t v[32], k[32]; int retry; };
HD void rfc6979_update(rfc6979& r, uint8_t sep, const uint8_t* data, size_t len) {      // K = HMAC_K(V || sep || data); V = HMAC_K(V)
    hmac_sha256 h;
    hmac_init32(h, r.k); hmac_write(h, r.v, 32); hmac_write(h, &sep, 1);
    if (len) hmac_write(h, data, len);
    hmac_final(h, r.k);
    hmac_init32(h, r.k); hmac_write(h, r.v, 32); hmac_final(h, r.v);
}
HD void rfc6979_init(rfc6979& r, const uint8_t* key, size_t keylen) {
    for (int i = 0; i < 32; i++) { r.v[i] = 1; r.k[i] = 0; }
    rfc6979_update(r, 0, key, keylen);
    rfc6979_update(r, 1, key, keylen);
    r.retry = 0;
}
HD void rfc6979_generate32(rfc6979& r, uint8_t* out32) {
    if (r.retry) rfc6979_update(r, 0, nullptr, 0);
    hmac_sha256 h;
    hmac_init32(h, r.k); hmac_write(h, r.v, 32); hmac_final(h, r.v);
    for (int i = 0; i < 32; i++) out32[i] = r.v[i];
    r.retry = 1;
}

// ---- the prover's random values (rangeproof_impl.h:61-112 with message = 4096 zero bytes, as rewind_inner calls it) ------------
// sec[i]: per-ring blinding factors (the last one = -(sum of the others)); prep[(4i+j)*32..]: the 32 bytes generated for position
// (i,j), which are both the forged s value (as a scalar) and the XOR pad of the embedded message.
HD void rp_genrand(sc* sec /*[32]*/, uint8_t* prep /*[4096]*/, const rp_header& h, const uint8_t* nonce32, const uint8_t* commit33,
                   const uint8_t* gen64, int gen_y_is_square, const uint8_t* proof) {
    uint8_t seed[32 + 33 + 33 + 10], tmp[32];
    const uint32_t len = h.off_signs;                                 // header bytes (<= 10)
    for (int i = 0; i < 32; i++) seed[i] = nonce32[i];
    seed[32] = commit33[0] & 1;                                       // rangeproof_serialize_point: [!is_square(y)] || x  (:53-59)
    for (int i = 0; i < 32; i++) seed[33 + i] = commit33[1 + i];
    seed[65] = (uint8_t)!gen_y_is_square;
    for (int i = 0; i < 32; i++) seed[66 + i] = gen64[i];
    for (uint32_t i = 0; i < len; i++) seed[98 + i] = proof[i];
    rfc6979 rng;
    rfc6979_init(rng, seed, 98 + len);
    sc acc; sc_set_zero(acc);
    for (uint32_t i = 0; i < h.rings; i++) {
        if (i + 1 < h.rings) {
            rfc6979_generate32(rng, tmp);
            int overflow;
            do {
                rfc6979_generate32(rng, tmp);
                overflow = sc_set_b32(sec[i], tmp);
            } while (overflow || sc_is_zero(sec[i]));
            sc t; sc_add(t, acc, sec[i]); acc = t;
        } else {
            sc_negate(sec[i], acc);
        }
        const uint32_t rs = rp_rsize(h, i);
        for (uint32_t j = 0; j < rs; j++) rfc6979_generate32(rng, prep + (i * 4 + j) * 32);
    }
}

// x = (k - s) / e   (rangeproof_impl.h:334-341)
HD void rp_recover_x(sc& x, const sc& k, const sc& e, const sc& s) {
    sc t, ei;
    sc_negate(t, s);
    sc_add(x, t, k);
    sc_inverse(ei, e);
    sc_mul(t, x, ei);
    x = t;
}

// One verified proof.  ev = the challenges of the ring walk, 32 bytes big-endian per position, index 4*ring + j.
// msg_cap = capacity of msg_out (0 or msg_out == nullptr: no message wanted).  Returns 1 and fills blind32 / value / msg_len on success.
// sec / prep: caller-provided scratch (32 scalars, 4096 bytes).  On the device they live in the global-memory workspace, not on the
// thread's stack: with the 4 KB array as a local, two sm_100a builds of k_rangeproof_rewind overwrote its first 512 / 64 bytes with
// the temporaries of the scalar inversion further down (wrong message bytes, correct blind and value; clean under the host build).
HD int rp_rewind_item(uint8_t* blind32, uint64_t& value_out, uint8_t* msg_out, size_t msg_cap, size_t& msg_len, const rp_header& h,
                      const uint8_t* proof, const uint8_t* commit33, const uint8_t* gen64, int gen_y_is_square, const uint8_t* nonce32,
                      const uint8_t* ev, const ecmult_gtable& gt, sc* sec, uint8_t* prep) {
    msg_len = 0;
    rp_genrand(sec, prep, h, nonce32, commit33, gen64, gen_y_is_square, proof);
    const uint8_t* sb = proof + h.off_s;
    sc blind, stmp, k, e, s;
    uint64_t value = 0;
    if (h.rings == 1 && h.last_rsize == 1) {                          // single-value proof: only the blinding factor (:386-396)
        sc_set_b32(k, prep); sc_set_b32(e, ev); sc_set_b32(s, sb);
        rp_recover_x(blind, k, e, s);
    } else {
        const uint32_t base = (h.rings - 1) << 2;
        uint32_t j;
        uint8_t tmp[32];
        for (j = 0; j < 2; j++) {                                     // the value is encoded in one of the last two slots (:398-416)
            const uint32_t idx = base + h.last_rsize - 1 - j;
            for (int b = 0; b < 32; b++) tmp[b] = sb[32 * idx + b] ^ prep[32 * idx + b];
            int same = (tmp[0] & 128) != 0;
            for (int b = 0; b < 8; b++) same &= (tmp[16 + b] == tmp[24 + b]) & (tmp[8 + b] == tmp[16 + b]);
            if (same) {
                value = 0;
                for (int b = 0; b < 8; b++) value = (value << 8) + tmp[24 + b];
                for (int b = 0; b < 32; b++) prep[32 * idx + b] = tmp[b];
                break;
            }
        }
        if (j > 1) return 0;
        uint32_t skip1 = h.last_rsize - 1 - j;
        uint32_t skip2 = (uint32_t)((value >> ((h.rings - 1) << 1)) & 3);
        if (skip1 == skip2) return 0;                                 // value in the wrong position (:426-432)
        skip1 += base; skip2 += base;
        sc_set_b32(k, prep + 32 * skip2); sc_set_b32(e, ev + 32 * skip2); sc_set_b32(s, sb + 32 * skip2);
        rp_recover_x(stmp, k, e, s);
        sc_negate(sec[h.rings - 1], sec[h.rings - 1]);                // stays negated for the message loop below, as in the reference (:437)
        sc_add(blind, stmp, sec[h.rings - 1]);
        if (msg_out != nullptr && msg_cap != 0) {
            size_t offset = 0;
            uint32_t npub = 0;
            for (uint32_t i = 0; i < h.rings; i++) {
                const uint32_t idx = (uint32_t)((value >> (i << 1)) & 3), rs = rp_rsize(h, i);
                for (uint32_t jj = 0; jj < rs; jj++, npub++) {
                    if (npub == skip1 || npub == skip2) continue;
                    sc_set_b32(s, sb + 32 * npub);
                    if (idx == jj) {                                  // a real signature: the message sits in the nonce k = s + x*e (:343-350)
                        sc_set_b32(e, ev + 32 * npub);
                        sc_mul(k, sec[i], e);
                        sc_add(stmp, s, k);
                    } else {
                        stmp = s;
                    }
                    sc_get_b32(tmp, stmp);
                    for (int b = 0; b < 32 && offset < msg_cap; b++) msg_out[offset++] = tmp[b] ^ prep[32 * npub + b];
                }
            }
            msg_len = offset;
        }
    }
    // can the commitment be rebuilt?  (:663-674)  vv wraps like the reference's uint64_t arithmetic
    const uint64_t vv = value * h.scale + h.min_value;
    uint8_t bb[32], c33[33];
    sc_get_b32(bb, blind);
    if (!pedersen_commit_item(c33, bb, vv, gen64, gt)) return 0;      // blind*G + vv*H == infinity (or blind out of range: cannot happen)
    int diff = 0;
    for (int b = 0; b < 33; b++) diff |= c33[b] ^ commit33[b];
    if (diff) return 0;
    for (int b = 0; b < 32; b++) blind32[b] = bb[b];
    value_out = vv;
    return 1;
}

}  // namespace B200_NS
