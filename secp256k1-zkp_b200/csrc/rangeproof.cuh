// rangeproof.cuh -- Back-Maxwell range proof verification, one WARP per proof, one lane per ring.
//
// Replaces secp256k1_rangeproof_verify (modules/rangeproof/main_impl.h:54-70) ->
// secp256k1_rangeproof_verify_impl (rangeproof_impl.h:541-683) -> secp256k1_borromean_verify
// (borromean_impl.h:53-104), including header parsing (rangeproof_impl.h:487-538), commitment loading
// (generator/main_impl.h:266-297) and the ring public-key expansion (rangeproof_impl.h:20-51).
//
// Parallel structure (the reference is strictly sequential): a 64-bit proof has 32 rings of 4 keys; the
// 4 ecmults of a ring are a hash chain, the rings are independent.  Lane i of the warp owns ring i:
//   P1  lanes 0..rings-2 decompress the digit commitments C_i (one sqrt each); lane rings-1 decompresses the
//       value commitment; lane 0 then hashes the message m (sha256_m of the reference)
//   P2  tree-sum of the C_i (+ min_value*H) in shared memory -> the last ring's first key = commit - sum
//   P3  every lane derives its ring's step point -(4^i * 10^exp) * H
//   P4  every lane walks its ring: e -> R = s*G + e*P -> affine -> serialize -> next e   (<= 4 rounds)
//   P5  lane 0 hashes the 32 ring results and compares with e0
// The pieces are HD functions so that the CPU tier can run the identical logic lane by lane.
#pragma once
#include "rp_getheader.cuh"
#include "ecmult.cuh"
#include "sha256.cuh"

namespace B200_NS {

#define RP_MAX_RINGS 32

struct rp_header {
    int exp, mantissa;
    uint64_t scale, min_value, max_value;
    uint32_t rings, last_rsize, npub;
    uint32_t off_signs;     // offset of the sign bitfield (= header length)
    uint32_t off_points;    // offset of the first 32-byte digit commitment
    uint32_t off_e0;        // offset of e0
    uint32_t off_s;         // offset of the first s value
};

// rangeproof_impl.h:487-538 + ring layout / length / padding checks (:566-603, :634-646 length part).
// min/max are written exactly when the reference writes them (also on later failure).
HD int rp_parse_header(rp_header& h, uint64_t& min_value, uint64_t& max_value, const uint8_t* proof, size_t plen) {
    size_t offset;
    if (!rp_getheader(offset, h.exp, h.mantissa, h.scale, min_value, max_value, proof, plen)) return 0;
    h.min_value = min_value; h.max_value = max_value;
    h.rings = 1; h.last_rsize = 1; h.npub = 1;
    if (h.mantissa != 0) {
        h.rings = (uint32_t)(h.mantissa >> 1);
        h.last_rsize = 4;
        h.npub = (uint32_t)((h.mantissa >> 1) << 2);
        if (h.mantissa & 1) { h.last_rsize = 2; h.npub += 2; h.rings++; }
    }
    if (plen - offset < 32 * ((size_t)h.npub + h.rings - 1) + 32 + ((h.rings + 6) >> 3)) return 0;
    h.off_signs = (uint32_t)offset;
    offset += (h.rings + 6) >> 3;
    if ((h.rings - 1) & 7) {
        if ((proof[offset - 1] >> ((h.rings - 1) & 7)) != 0) return 0;   // unused sign bits must be zero
    }
    h.off_points = (uint32_t)offset;
    h.off_e0 = h.off_points + 32 * (h.rings - 1);
    h.off_s = h.off_e0 + 32;
    if ((size_t)h.off_s + 32 * (size_t)h.npub != plen) return 0;           // trailing / missing bytes (:643-646)
    return 1;
}
HD uint32_t rp_rsize(const rp_header& h, uint32_t ring) { return ring + 1 == h.rings ? h.last_rsize : 4u; }
HD int rp_sign(const rp_header& h, const uint8_t* proof, uint32_t i) { return (proof[h.off_signs + (i >> 3)] >> (i & 7)) & 1; }

// 64-byte opaque generator = x||y big-endian (generator/main_impl.h:40-49)
HD void rp_generator_load(ge& g, const uint8_t* gen64) { fe_set_b32(g.x, gen64); fe_set_b32(g.y, gen64 + 32); g.inf = 0; }

// 33-byte commitment (tag 8/9 + x): parse (generator/main_impl.h:281-297) + load (:266-273).  0 if invalid.
HD int rp_commitment_load(ge& c, const uint8_t* commit33) {
    fe x;
    if ((commit33[0] & 0xFE) != 8) return 0;
    if (!fe_set_b32(x, commit33 + 1)) return 0;
    if (!ge_set_xquad(c, x)) return 0;
    if (commit33[0] & 1) fe_neg(c.y, c.y);
    return 1;
}

// P1, lanes < rings-1: digit commitment i (rangeproof_impl.h:609-618)
HD int rp_digit_commitment_load(ge& c, const rp_header& h, const uint8_t* proof, uint32_t i) {
    fe x;
    if (!fe_set_b32(x, proof + h.off_points + 32 * i)) return 0;
    if (!ge_set_xquad(c, x)) return 0;
    if (rp_sign(h, proof, i)) fe_neg(c.y, c.y);
    return 1;
}

// message hash m = SHA256(ser(commit) || ser(gen) || header || (sign_i || x_i)* || extra)   (rangeproof_impl.h:588-651)
// ser(P) = [!is_square(P.y)] || x (rangeproof_impl.h:53-59).  For the commitment, y is the square root of x^3+7
// negated iff tag&1, and -1 is a non-residue, so !is_square(y) == tag & 1: no field work needed.
HD void rp_message_hash(uint8_t* m32, const rp_header& h, const uint8_t* proof, const uint8_t* commit33, const uint8_t* gen64,
                        int gen_y_is_square, const uint8_t* extra, size_t extra_len) {
    sha256 s;
    sha256_init(s);
    uint8_t b = commit33[0] & 1;
    sha256_write(s, &b, 1); sha256_write(s, commit33 + 1, 32);
    b = (uint8_t)!gen_y_is_square;
    sha256_write(s, &b, 1); sha256_write(s, gen64, 32);
    sha256_write(s, proof, h.off_signs);
    for (uint32_t i = 0; i + 1 < h.rings; i++) {
        b = (uint8_t)rp_sign(h, proof, i);
        sha256_write(s, &b, 1);
        sha256_write(s, proof + h.off_points + 32 * i, 32);
    }
    if (extra != nullptr && extra_len) sha256_write(s, extra, extra_len);
    sha256_finalize(s, m32);
}

// min_value * H  (pedersen_impl.h:34-39 "ecmult_small"; plain double-and-add over the 64-bit value)
HD void rp_small_mult(gej& r, uint64_t v, const ge& genp) {
    gej_set_infinity(r);
#pragma unroll 1
    for (int k = 63; k >= 0; k--) {
        gej t;
        if (!r.inf) { gej_double(t, r); r = t; }
        if ((v >> k) & 1) { gej_add_ge(t, r, genp); r = t; }
    }
}

// P3: step point of ring i: -(4^i * 10^exp) * H   (rangeproof_impl.h:20-51)
HD void rp_ring_base(gej& base, const ge& genp, int exp, uint32_t ring) {
    ge ng; ge_neg(ng, genp);
    gej_set_ge(base, ng);
    if (exp < 0) exp = 0;
#pragma unroll 1
    for (int e = 0; e < exp; e++) {     // x10 = x8 + x2
        gej t2, t8, t;
        gej_double(t2, base);
        gej_double(t, t2);
        gej_double(t8, t);
        gej_add(base, t8, t2);
    }
#pragma unroll 1
    for (uint32_t k = 0; k < 2 * ring; k++) { gej t; gej_double(t, base); base = t; }
}

// borromean_hash (borromean_impl.h:23-37): SHA256(e || m || be32(ring) || be32(idx)) -> scalar; returns overflow flag.
// e is e0 (32 bytes) or a serialized point (33 bytes): one full block plus 8 or 9 bytes.  Both layouts are assembled as big-endian words
// straight from the inputs -- no byte buffer, no generic write loop: this hash runs 128 times per 64-bit proof in every lane (r2m: the
// byte-oriented form was 10 % of k_rangeproof_verify's instructions and 46 % of its code size).
HD uint32_t rp_be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }
HD int rp_borromean_challenge(sc& ens, const uint8_t* e, uint32_t elen, const uint8_t* m32, uint32_t ring, uint32_t idx) {
    uint8_t out[32];
    if (elen == 32 || elen == 33) {
        uint32_t st[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        uint32_t w[16];
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = rp_be32(e + 4 * k);
        if (elen == 32) {
#pragma unroll
            for (int k = 0; k < 8; k++) w[8 + k] = rp_be32(m32 + 4 * k);
        } else {                                                    // the stream is m shifted by one byte
            w[8] = ((uint32_t)e[32] << 24) | ((uint32_t)m32[0] << 16) | ((uint32_t)m32[1] << 8) | (uint32_t)m32[2];
#pragma unroll
            for (int k = 1; k < 8; k++) w[8 + k] = rp_be32(m32 + 4 * k - 1);
        }
        sha256_compress(st, w);
#pragma unroll
        for (int k = 0; k < 16; k++) w[k] = 0;
        if (elen == 32) {
            w[0] = ring; w[1] = idx; w[2] = 0x80000000u;
        } else {
            w[0] = ((uint32_t)m32[31] << 24) | (ring >> 8);
            w[1] = (ring << 24) | (idx >> 8);
            w[2] = (idx << 24) | 0x00800000u;
        }
        w[15] = (elen + 40) * 8;
        sha256_compress(st, w);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            out[4 * k] = (uint8_t)(st[k] >> 24); out[4 * k + 1] = (uint8_t)(st[k] >> 16); out[4 * k + 2] = (uint8_t)(st[k] >> 8); out[4 * k + 3] = (uint8_t)st[k];
        }
    } else {
        sha256 s;
        uint8_t w[8] = {(uint8_t)(ring >> 24), (uint8_t)(ring >> 16), (uint8_t)(ring >> 8), (uint8_t)ring,
                        (uint8_t)(idx >> 24), (uint8_t)(idx >> 16), (uint8_t)(idx >> 8), (uint8_t)idx};
        sha256_init(s);
        sha256_write(s, e, elen); sha256_write(s, m32, 32); sha256_write(s, w, 8);
        sha256_finalize(s, out);
    }
    return sc_set_b32(ens, out);
}

// P4: walk ring `ring` (borromean_impl.h:69-99).  pub = first key of the ring (Jacobian), base = step point.
// Writes the 33-byte serialization of the ring's last R.  Returns 0 on any reject condition.
// ev_out (optional): 32 bytes per position, index 4*ring + j, big-endian challenge e_{ring,j}.
template <class Tab>
HD int rp_ring_walk(uint8_t* r33, const rp_header& h, const uint8_t* proof, const uint8_t* m32, uint32_t ring,
                    gej pub, const gej& base, const ecmult_gtable& gt, Tab& tab, uint8_t* ev_out = nullptr) {
    uint32_t rsize = rp_rsize(h, ring);
    const uint8_t* sbase = proof + h.off_s + 32 * (4 * ring);    // rings before the last all have 4 keys
    sc ens;
    int overflow = rp_borromean_challenge(ens, proof + h.off_e0, 32, m32, ring, 0);
    uint8_t tmp[33];
#pragma unroll 1
    for (uint32_t j = 0; j < rsize; j++) {
        if (j > 0) { gej t; gej_add(t, pub, base); pub = t; }
        sc s;
        if (sc_set_b32(s, sbase + 32 * j)) return 0;              // s >= n (rangeproof_impl.h:637-640)
        if (overflow || sc_is_zero(s) || sc_is_zero(ens) || pub.inf) return 0;
        if (ev_out) sc_get_b32(ev_out + 32 * (4 * ring + j), ens);   // the challenges, for rewinding (borromean_impl.h:76-78 "evalues")
        gej R;
        ecmult(R, pub, ens, s, gt, tab);
        if (R.inf) return 0;
        ge Ra; ge_set_gej(Ra, R);
        tmp[0] = fe_is_odd(Ra.y) ? 3 : 2;                         // eckey_impl.h:38-45
        fe_get_b32(tmp + 1, Ra.x);
        if (j + 1 != rsize) overflow = rp_borromean_challenge(ens, tmp, 33, m32, ring, j + 1);
    }
    for (int k = 0; k < 33; k++) r33[k] = tmp[k];
    return 1;
}

// P5: e0 == SHA256(R_0 || ... || R_{rings-1} || m)   (borromean_impl.h:100-103)
HD int rp_final_check(const rp_header& h, const uint8_t* proof, const uint8_t* r33_all /*[rings][33]*/, const uint8_t* m32) {
    sha256 s;
    uint8_t out[32];
    sha256_init(s);
    sha256_write(s, r33_all, 33 * (size_t)h.rings);
    sha256_write(s, m32, 32);
    sha256_finalize(s, out);
    int diff = 0;
    for (int k = 0; k < 32; k++) diff |= out[k] ^ proof[h.off_e0 + k];
    return diff == 0;
}

// Pedersen commitment: blind*G + value*H, serialized (generator/main_impl.h:309-335, :275-279).
// Synthetic-throughput path: NOT constant time (the batched commit is for public recomputation / benchmarking).
HD int pedersen_commit_item(uint8_t* commit33, const uint8_t* blind32, uint64_t value, const uint8_t* gen64, const ecmult_gtable& gt) {
    sc sec;
    ge genp;
    rp_generator_load(genp, gen64);
    if (sc_set_b32(sec, blind32)) { for (int k = 0; k < 33; k++) commit33[k] = 0; return 0; }
    gej rj;
    rp_small_mult(rj, value, genp);
    ecmult_gen_accumulate(rj, sec, gt);
    if (rj.inf) { for (int k = 0; k < 33; k++) commit33[k] = 0; return 0; }
    ge r; ge_set_gej(r, rj);
    commit33[0] = (uint8_t)(9 ^ fe_is_square(r.y));
    fe_get_b32(commit33 + 1, r.x);
    return 1;
}

}  // namespace B200_NS
