// surjection.cuh -- asset surjection proof verification and Pedersen tally check, one thread per proof / per tally.
// together with range proofs this completes Confidential-Assets output validation.
//   surjection_verify_item : secp256k1_surjectionproof_parse (modules/surjection/main_impl.h:46-83) +
//                            secp256k1_surjectionproof_verify (:360-402): message hash (surjection_impl.h:18-36), ring keys
//                            output_tag - input_tag_i (:66-92), one-ring Borromean verification (borromean_impl.h:53-104)
//   pedersen_tally_item    : secp256k1_pedersen_verify_tally (modules/generator/main_impl.h:371-396)
// The ring of a surjection proof is ONE hash chain (each challenge depends on the previous R), so a proof is strictly
// sequential; the batch supplies the parallelism.
#pragma once
#include "rangeproof.cuh"

namespace B200_NS {

#define SJ_MAX_INPUTS 256

HD int sj_popcount8(uint8_t v) { int c = 0; for (int k = 0; k < 8; k++) c += (v >> k) & 1; return c; }

// in_tags64: n_tags generators (64 bytes x||y big-endian each); out_tag64: the output generator
template <class Tab>
HD int surjection_verify_item(const uint8_t* proof, size_t plen, const uint8_t* in_tags64, uint32_t n_tags, const uint8_t* out_tag64,
                              const ecmult_gtable& gt, Tab& tab) {
    // ---- parse (:46-83)
    if (plen < 2) return 0;
    uint32_t n_inputs = (uint32_t)proof[0] | ((uint32_t)proof[1] << 8);
    if (n_inputs > SJ_MAX_INPUTS) return 0;
    uint32_t bm = (n_inputs + 7) / 8;
    if (plen < 2 + (size_t)bm) return 0;
    if (n_inputs % 8 != 0) {
        uint8_t mask = (uint8_t)(0xFFu << (n_inputs % 8));
        if (proof[2 + bm - 1] & mask) return 0;
    }
    uint32_t n_used = 0;
    for (uint32_t k = 0; k < bm; k++) n_used += (uint32_t)sj_popcount8(proof[2 + k]);
    if (plen != 2 + (size_t)bm + 32 * (1 + (size_t)n_used)) return 0;
    // ---- verify preconditions (:377-385)
    if (n_used == 0 || n_used > n_inputs || n_inputs != n_tags) return 0;
    const uint8_t* e0 = proof + 2 + bm;
    const uint8_t* sdata = e0 + 32;
    for (uint32_t j = 0; j < n_used; j++) { sc t; if (sc_set_b32(t, sdata + 32 * j)) return 0; }   // (:391-397)
    // ---- message (surjection_impl.h:18-36)
    uint8_t m32[32];
    {
        sha256 h; sha256_init(h);
        uint8_t ser[33];
        for (uint32_t i = 0; i <= n_inputs; i++) {
            const uint8_t* t = i < n_inputs ? in_tags64 + 64 * (size_t)i : out_tag64;
            ser[0] = (uint8_t)(2 + (t[63] & 1));
            for (int k = 0; k < 32; k++) ser[1 + k] = t[k];
            sha256_write(h, ser, 33);
        }
        sha256_finalize(h, m32);
    }
    // ---- one-ring Borromean verification
    ge outp; rp_generator_load(outp, out_tag64);
    sc ens;
    int overflow = rp_borromean_challenge(ens, e0, 32, m32, 0, 0);
    uint8_t tmp[33];
    uint32_t j = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < n_inputs; i++) {
        if (!((proof[2 + (i >> 3)] >> (i & 7)) & 1)) continue;
        ge inp; rp_generator_load(inp, in_tags64 + 64 * (size_t)i);
        fe_neg(inp.y, inp.y);
        gej pub, t0;
        gej_set_ge(t0, inp);
        gej_add_ge(pub, t0, outp);                              // output_tag - input_tag_i
        sc s; sc_set_b32(s, sdata + 32 * j);
        if (overflow || sc_is_zero(s) || sc_is_zero(ens) || pub.inf) return 0;
        gej R;
        ecmult(R, pub, ens, s, gt, tab);
        if (R.inf) return 0;
        ge Ra; ge_set_gej(Ra, R);
        tmp[0] = fe_is_odd(Ra.y) ? 3 : 2;
        fe_get_b32(tmp + 1, Ra.x);
        j++;
        if (j != n_used) overflow = rp_borromean_challenge(ens, tmp, 33, m32, 0, j);
    }
    sha256 h; sha256_init(h);
    uint8_t out[32];
    sha256_write(h, tmp, 33); sha256_write(h, m32, 32);
    sha256_finalize(h, out);
    int diff = 0;
    for (int k = 0; k < 32; k++) diff |= out[k] ^ e0[k];
    return diff == 0;
}

// sum(pos) - sum(neg) == infinity; any commitment that does not parse rejects the tally
HD int pedersen_tally_item(const uint8_t* pos33, uint32_t pcnt, const uint8_t* neg33, uint32_t ncnt) {
    gej acc; gej_set_infinity(acc);
#pragma unroll 1
    for (uint32_t i = 0; i < ncnt; i++) {
        ge c; gej t;
        if (!rp_commitment_load(c, neg33 + 33 * (size_t)i)) return 0;
        gej_add_ge(t, acc, c); acc = t;
    }
    { gej t; gej_neg(t, acc); acc = t; }
#pragma unroll 1
    for (uint32_t i = 0; i < pcnt; i++) {
        ge c; gej t;
        if (!rp_commitment_load(c, pos33 + 33 * (size_t)i)) return 0;
        gej_add_ge(t, acc, c); acc = t;
    }
    return acc.inf;
}

}  // namespace B200_NS
