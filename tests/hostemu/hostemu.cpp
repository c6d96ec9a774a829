// tests/hostemu/hostemu.cpp -- TEST-ONLY host build of the device headers (HD functions compile
// as plain C++ when __CUDACC__ is undefined, using the portable big-number path).
// Lets the CPU-only test tier check the group law / ecmult / verify LOGIC against the oracle
// without a GPU.  Never linked into, nor loaded by, the product library.
#include "../../secp256k1-zkp_b200/csrc/modinv.cuh"
using namespace b200;
extern "C" {
void emu_fe_mul(uint8_t* r, const uint8_t* a, const uint8_t* b) { fe x, y, z; fe_set_b32(x, a); fe_set_b32(y, b); fe_mul(z, x, y); fe_get_b32(r, z); }
void emu_fe_sqr(uint8_t* r, const uint8_t* a) { fe x, z; fe_set_b32(x, a); fe_sqr(z, x); fe_get_b32(r, z); }
void emu_fe_add(uint8_t* r, const uint8_t* a, const uint8_t* b) { fe x, y, z; fe_set_b32(x, a); fe_set_b32(y, b); fe_add(z, x, y); fe_get_b32(r, z); }
void emu_fe_sub(uint8_t* r, const uint8_t* a, const uint8_t* b) { fe x, y, z; fe_set_b32(x, a); fe_set_b32(y, b); fe_sub(z, x, y); fe_get_b32(r, z); }
void emu_fe_neg(uint8_t* r, const uint8_t* a) { fe x, z; fe_set_b32(x, a); fe_neg(z, x); fe_get_b32(r, z); }
void emu_fe_half(uint8_t* r, const uint8_t* a) { fe x, z; fe_set_b32(x, a); fe_half(z, x); fe_get_b32(r, z); }
void emu_fe_mul_small(uint8_t* r, const uint8_t* a, uint32_t k) { fe x, z; fe_set_b32(x, a); fe_mul_small(z, x, k); fe_get_b32(r, z); }
void emu_fe_inv(uint8_t* r, const uint8_t* a) { fe x, z; fe_set_b32(x, a); fe_inv(z, x); fe_get_b32(r, z); }
int emu_fe_sqrt(uint8_t* r, const uint8_t* a) { fe x, z; fe_set_b32(x, a); int ok = fe_sqrt(z, x); fe_get_b32(r, z); return ok; }
int emu_fe_is_zero(const uint8_t* a) { fe x; fe_set_b32(x, a); return fe_is_zero(x); }
}

// ---- safegcd inverse / Jacobi symbol (modinv.cuh) ---------------------------------------------------
#include "../../secp256k1-zkp_b200/csrc/modinv.cuh"
extern "C" {
void emu_fe_inv_safegcd(uint8_t* r, const uint8_t* a, size_t n) { for (size_t i = 0; i < n; i++) { fe x, z; fe_set_b32(x, a + 32 * i); fe_inv_safegcd(z, x); fe_get_b32(r + 32 * i, z); } }
void emu_fe_is_square_jacobi(uint8_t* r, const uint8_t* a, size_t n) { for (size_t i = 0; i < n; i++) { fe x; fe_set_b32(x, a + 32 * i); r[i] = (uint8_t)fe_is_square_jacobi(x); } }
void emu_sc_inverse_safegcd(uint8_t* r, const uint8_t* a, size_t n) { for (size_t i = 0; i < n; i++) { sc x, z; sc_set_b32(x, a + 32 * i); sc_inverse_safegcd(z, x); sc_get_b32(r + 32 * i, z); } }
}

// ---- scalar / group / ecmult / verify logic -------------------------------
#include "../../secp256k1-zkp_b200/csrc/gtable.cuh"
#include "../../secp256k1-zkp_b200/csrc/verify.cuh"
#include "../../secp256k1-zkp_b200/csrc/testdata.cuh"
#include "../../secp256k1-zkp_b200/csrc/rewind.cuh"
#include "../../secp256k1-zkp_b200/csrc/msm.cuh"
#include "../../secp256k1-zkp_b200/csrc/bppp.cuh"
#include "../../secp256k1-zkp_b200/csrc/surjection.cuh"
#include "../../secp256k1-zkp_b200/csrc/halfagg.cuh"
#include <vector>
#include <string.h>
static std::vector<uint32_t> g_tab;
static ecmult_gtable g_gt = {nullptr, 0, 0};
static bip340_midstate g_mid;
static void ge_from_xy64(ge& p, const uint8_t* xy) {
    int z = 1; for (int i = 0; i < 64; i++) z &= (xy[i] == 0);
    if (z) { ge_set_infinity(p); return; }
    fe_set_b32(p.x, xy); fe_set_b32(p.y, xy + 32); p.inf = 0;
}
static int gej_to_xy64(uint8_t* xy, const gej& a) {
    if (a.inf) { memset(xy, 0, 64); return 0; }
    ge p; ge_set_gej(p, a);
    fe_get_b32(xy, p.x); fe_get_b32(xy + 32, p.y);
    return 1;
}
extern "C" {
void emu_init(int gw) {
    int windows = (256 + gw - 1) / gw;
    g_tab.assign(ecmult_gtable_words(gw), 0);
    std::vector<uint32_t> bases(16 * windows);
    gtable_compute_bases(bases.data(), gw, windows);
    for (int j = 0; j < windows; j++)
        for (uint32_t d = 1; d < (1u << gw); d++)
            gtable_compute_entry(g_tab.data() + (((size_t)j << gw) + d) * 16, bases.data() + 16 * j, d, gw);
    g_gt.data = g_tab.data(); g_gt.gw = gw; g_gt.windows = windows;
    sha256 h; sha256_init_tagged(h, (const uint8_t*)"BIP0340/challenge", 17);
    for (int i = 0; i < 8; i++) g_mid.s[i] = h.s[i];
}
void emu_sc_mul(uint8_t* r, const uint8_t* a, const uint8_t* b) { sc x, y, z; sc_set_b32(x, a); sc_set_b32(y, b); sc_mul(z, x, y); sc_get_b32(r, z); }
void emu_sc_add(uint8_t* r, const uint8_t* a, const uint8_t* b) { sc x, y, z; sc_set_b32(x, a); sc_set_b32(y, b); sc_add(z, x, y); sc_get_b32(r, z); }
void emu_sc_negate(uint8_t* r, const uint8_t* a) { sc x, z; sc_set_b32(x, a); sc_negate(z, x); sc_get_b32(r, z); }
void emu_sc_inverse(uint8_t* r, const uint8_t* a) { sc x, z; sc_set_b32(x, a); sc_inverse(z, x); sc_get_b32(r, z); }
int emu_sc_set_b32(uint8_t* r, const uint8_t* a) { sc x; int ov = sc_set_b32(x, a); sc_get_b32(r, x); return ov; }
int emu_sc_is_high(const uint8_t* a) { sc x; sc_set_b32(x, a); return sc_is_high(x); }
void emu_sc_split_lambda(uint8_t* r1, uint8_t* r2, const uint8_t* k) { sc x, a, b; sc_set_b32(x, k); sc_split_lambda(a, b, x); sc_get_b32(r1, a); sc_get_b32(r2, b); }
void emu_sha256(uint8_t* out, const uint8_t* data, size_t len) { sha256 h; sha256_init(h); sha256_write(h, data, len); sha256_finalize(h, out); }
int emu_ge_double(uint8_t* r, const uint8_t* a) { ge p; ge_from_xy64(p, a); gej j, t; gej_set_ge(j, p); gej_double(t, j); return gej_to_xy64(r, t); }
int emu_ge_add(uint8_t* r, const uint8_t* a, const uint8_t* b) { ge p, q; ge_from_xy64(p, a); ge_from_xy64(q, b); gej j, k, t; gej_set_ge(j, p); gej_set_ge(k, q); gej_add(t, j, k); return gej_to_xy64(r, t); }
int emu_ge_add_mixed(uint8_t* r, const uint8_t* a, const uint8_t* b) { ge p, q; ge_from_xy64(p, a); ge_from_xy64(q, b); gej j, t; gej_set_ge(j, p);
    gej d; gej_double(d, j); gej nd; /* make z != 1: (2a) + (-a) = a, then + b */ ge np; ge_neg(np, p); gej_add_ge(nd, d, np); gej_add_ge(t, nd, q); return gej_to_xy64(r, t); }
int emu_ecmult(uint8_t* r, const uint8_t* a, const uint8_t* na, const uint8_t* ng) {
    ge p; ge_from_xy64(p, a); gej pj, R; gej_set_ge(pj, p);
    sc x, y; sc_set_b32(x, na); if (ng) sc_set_b32(y, ng); else sc_set_zero(y);
    tab_local tab;
    ecmult(R, pj, x, y, g_gt, tab);
    return gej_to_xy64(r, R);
}
int emu_schnorr_verify(const uint8_t* sig, const uint8_t* msg, uint32_t msglen, const uint8_t* pk, int pkfmt) {
    tab_local tab; return schnorr_verify_item(sig, msg, msglen, pk, pkfmt, g_mid, g_gt, tab);
}
int emu_schnorr_sign(uint8_t* sig64, uint8_t* pk32, const uint8_t* sk32, const uint8_t* msg, uint32_t msglen) {
    bip340_sign_consts c; sha256 h;
    sha256_init_tagged(h, (const uint8_t*)"BIP0340/challenge", 17); for (int i = 0; i < 8; i++) c.mid_challenge[i] = h.s[i];
    sha256_init_tagged(h, (const uint8_t*)"BIP0340/nonce", 13); for (int i = 0; i < 8; i++) c.mid_nonce[i] = h.s[i];
    const uint8_t zeros[32] = {0};
    sha256_init_tagged(h, (const uint8_t*)"BIP0340/aux", 11); sha256_write(h, zeros, 32); sha256_finalize(h, c.zero_mask);
    return schnorr_sign_item(sig64, pk32, sk32, msg, msglen, c, g_gt);
}
// same phases as k_rangeproof_verify, lanes executed one after the other
static int rp_verify_core(uint64_t* minv, uint64_t* maxv, const uint8_t* commit33, const uint8_t* proof, size_t plen,
                          const uint8_t* gen64, const uint8_t* extra, size_t extra_len, uint8_t* ev) {
    rp_header h; uint64_t mn = 0, mx = 0;
    int hdr_ok = rp_parse_header(h, mn, mx, proof, plen);
    *minv = mn; *maxv = mx;
    if (!hdr_ok) return 0;
    ge genp; rp_generator_load(genp, gen64);
    ge mine[32]; gej pts[32]; int good = 1;
    for (uint32_t lane = 0; lane < 32; lane++) {
        mine[lane].inf = 1; int g = 1;
        if (lane + 1 < h.rings) g = rp_digit_commitment_load(mine[lane], h, proof, lane);
        else if (lane + 1 == h.rings) g = rp_commitment_load(mine[lane], commit33);
        if (lane + 1 < h.rings && g) gej_set_ge(pts[lane], mine[lane]); else gej_set_infinity(pts[lane]);
        if (lane + 1 == h.rings && !g) { *minv = 0; *maxv = 0; }
        good &= g;
    }
    uint8_t m32[32], r33[32 * 33];
    rp_message_hash(m32, h, proof, commit33, gen64, fe_is_square(genp.y), extra, extra_len);
    for (int d = 16; d >= 1; d >>= 1) for (int lane = 0; lane < d; lane++) { gej t; gej_add(t, pts[lane], pts[lane + d]); pts[lane] = t; }
    gej pub[32];
    for (uint32_t lane = 0; lane < 32; lane++) {
        gej_set_infinity(pub[lane]);
        if (lane + 1 == h.rings && good) {
            gej acc = pts[0];
            if (h.min_value) { gej mh, t; rp_small_mult(mh, h.min_value, genp); gej_add(t, acc, mh); acc = t; }
            gej nacc; gej_neg(nacc, acc);
            gej_add_ge(pub[lane], nacc, mine[lane]);
            if (pub[lane].inf) good = 0;
        } else if (lane + 1 < h.rings) gej_set_ge(pub[lane], mine[lane]);
    }
    if (!good) return 0;
    int ring_ok = 1;
    for (uint32_t lane = 0; lane < h.rings; lane++) {
        gej base; rp_ring_base(base, genp, h.exp, lane);
        tab_local tab;
        ring_ok &= rp_ring_walk(r33 + 33 * lane, h, proof, m32, lane, pub[lane], base, g_gt, tab, ev);
    }
    return ring_ok && rp_final_check(h, proof, r33, m32);
}
int emu_rangeproof_verify(uint64_t* minv, uint64_t* maxv, const uint8_t* commit33, const uint8_t* proof, size_t plen,
                          const uint8_t* gen64, const uint8_t* extra, size_t extra_len) {
    return rp_verify_core(minv, maxv, commit33, proof, plen, gen64, extra, extra_len, nullptr);
}
// k_rangeproof_verify (challenges kept) followed by k_rangeproof_rewind, for one proof
int emu_rangeproof_rewind(uint8_t* blind32, uint64_t* value, uint8_t* msg, size_t msg_cap, size_t* msg_len, uint64_t* minv, uint64_t* maxv,
                          const uint8_t* nonce32, const uint8_t* commit33, const uint8_t* proof, size_t plen, const uint8_t* gen64,
                          const uint8_t* extra, size_t extra_len) {
    std::vector<uint8_t> ev(4096);
    memset(blind32, 0, 32); *value = 0; *msg_len = 0;
    if (!rp_verify_core(minv, maxv, commit33, proof, plen, gen64, extra, extra_len, ev.data())) return 0;
    rp_header h; uint64_t mn, mx;
    if (!rp_parse_header(h, mn, mx, proof, plen)) return 0;
    ge genp; rp_generator_load(genp, gen64);
    uint8_t bl[32]; uint64_t v = 0; size_t ml = 0;
    std::vector<sc> sec(RP_MAX_RINGS); std::vector<uint8_t> prep(4096);
    if (!rp_rewind_item(bl, v, msg_cap ? msg : nullptr, msg_cap, ml, h, proof, commit33, gen64, fe_is_square(genp.y), nonce32, ev.data(), g_gt, sec.data(), prep.data())) return 0;
    memcpy(blind32, bl, 32); *value = v; *msg_len = ml;
    return 1;
}
// the ring challenge hash with its two word-assembled layouts (e of 32 / 33 bytes) and the generic byte path (any other length)
int emu_borromean_challenge(uint8_t* out32, const uint8_t* e, uint32_t elen, const uint8_t* m32, uint32_t ring, uint32_t idx) {
    sc s;
    int ov = rp_borromean_challenge(s, e, elen, m32, ring, idx);
    sc_get_b32(out32, s);
    return ov;
}
int emu_pedersen_commit(uint8_t* out33, const uint8_t* blind32, uint64_t value, const uint8_t* gen64) {
    return pedersen_commit_item(out33, blind32, value, gen64, g_gt);
}
// the k_msm_* kernel pipeline (k_msm.cu) executed sequentially with the same index logic: digits -> histogram -> scan (entries / pieces /
// parts) -> scatter -> units -> unit sums -> multi-piece combine -> chunk running sums -> plain / bit sums -> window sums -> Horner -> join.
// c = 0: window chosen from n; tp != 0 forces the piece size (to exercise multi-piece buckets with few terms).
struct emu_emit_hist { std::vector<uint32_t>* h; void operator()(uint32_t b, uint32_t) { (*h)[b]++; } };
struct emu_emit_scatter { std::vector<uint32_t>* off; std::vector<uint32_t>* h; std::vector<uint32_t>* ent; uint32_t idx;
    void operator()(uint32_t b, uint32_t sgn) { uint32_t pos = (*off)[b] + ((*h)[b]--) - 1u; (*ent)[pos] = idx | (sgn << 31); } };
static void emu_msm_core(gej& r, const uint8_t* g_sc32, const uint8_t* scalars32, const uint8_t* points64, size_t n, int c, uint32_t tp, uint32_t shard, uint32_t shards) {
    const size_t N = n + 1;
    msm_plan p; msm_make_plan_c(p, N, c ? (uint32_t)c : msm_choose_c(N), shard, shards);
    if (tp) p.Tp = tp;
    std::vector<uint32_t> k12(N * 8), hist(p.TB, 0), off(p.TB + 1), pbase(p.TB + 1), mbase(p.TB + 1), multi, entries((size_t)2 * p.nw * N);
    std::vector<uint8_t> flags(N);
    uint8_t gpoint[64]; { ge g; ge_set_generator(g); fe_get_b32(gpoint, g.x); fe_get_b32(gpoint + 32, g.y); }
    for (size_t i = 0; i < N; i++) {
        sc k; int inf = 0;
        if (i < n) { sc_set_b32(k, scalars32 + 32 * i); inf = 1; for (int t = 0; t < 64; t++) if (points64[64 * i + t]) inf = 0; }
        else { if (g_sc32) sc_set_b32(k, g_sc32); else sc_set_zero(k); }
        uint32_t f; msm_prepare_term(&k12[8 * i], f, k, inf); flags[i] = (uint8_t)f;
        if (f & 4u) continue;
        emu_emit_hist em = {&hist};
        msm_emit_digits(p, &k12[8 * i], f & 1u, 0, em);
        msm_emit_digits(p, &k12[8 * i + 4], (f >> 1) & 1u, 1, em);
    }
    uint32_t ro = 0, rp = 0, rm = 0;
    for (uint32_t b = 0; b < p.TB; b++) {
        uint32_t pc = (hist[b] + p.Tp - 1) / p.Tp;
        off[b] = ro; pbase[b] = rp; mbase[b] = rm;
        if (pc > 1) multi.push_back(b);
        ro += hist[b]; rp += pc; rm += (pc > 1 ? pc : 0);
    }
    off[p.TB] = ro; pbase[p.TB] = rp; mbase[p.TB] = rm;
    for (size_t i = 0; i < N; i++) {
        uint32_t f = flags[i];
        if (f & 4u) continue;
        emu_emit_scatter em = {&off, &hist, &entries, (uint32_t)i};
        msm_emit_digits(p, &k12[8 * i], f & 1u, 0, em);
        msm_emit_digits(p, &k12[8 * i + 4], (f >> 1) & 1u, 1, em);
    }
    std::vector<gej_store> buckets(p.TB), parts(rm + 1), Lq(p.TB / MSM_CHUNK), Aq(p.TB / MSM_CHUNK), T((size_t)2 * p.nw * MSM_MAX_SUMS * MSM_R2_PARTS);
    memset(T.data(), 0, T.size() * sizeof(gej_store));
    memset(buckets.data(), 0, buckets.size() * sizeof(gej_store));
    for (uint32_t u = 0; u < rp; u++) {                                   // k_msm_units + k_msm_bucket_sum
        uint32_t lo = 0, hi = p.TB;
        while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (pbase[mid] <= u) lo = mid; else hi = mid; }
        uint32_t b = lo, j = u - pbase[b], npieces = pbase[b + 1] - pbase[b];
        msm_unit d; d.begin = off[b] + j * p.Tp; d.count = off[b + 1] - d.begin < p.Tp ? off[b + 1] - d.begin : p.Tp; d.first = entries[d.begin];
        d.dst = npieces == 1 ? b : (MSM_DST_PART | (mbase[b] + j));
        gej acc; msm_unit_sum(acc, d, entries.data(), points64, n, gpoint);
        msm_store_gej((d.dst & MSM_DST_PART) ? parts[d.dst & ~MSM_DST_PART] : buckets[d.dst], acc);
    }
    for (uint32_t b : multi) {                                            // k_msm_combine
        gej acc; gej_set_infinity(acc);
        for (uint32_t k = 0; k < pbase[b + 1] - pbase[b]; k++) { gej o, t; msm_load_gej(o, parts[mbase[b] + k]); gej_add(t, acc, o); acc = t; }
        msm_store_gej(buckets[b], acc);
    }
    for (uint32_t q = 0; q < p.TB / MSM_CHUNK; q++) { gej L, A; msm_chunk_r1(L, A, buckets.data() + (size_t)q * MSM_CHUNK); msm_store_gej(Lq[q], L); msm_store_gej(Aq[q], A); }
    std::vector<gej> W(2 * p.nw);
    for (uint32_t wi = 0; wi < 2 * p.nw; wi++) {                          // k_msm_r2 + first phase of k_msm_tail
        msm_win w; msm_window_by_index(w, p, wi);
        uint32_t Q = w.Bw / MSM_CHUNK, nbits = msm_ilog2(Q), cb = w.base / MSM_CHUNK;
        for (uint32_t s = 0; s <= nbits; s++) {
            gej tot; gej_set_infinity(tot);
            for (uint32_t part = 0; part < MSM_R2_PARTS; part++) {                 // the blocks of one sum, then the last block's combine
                gej acc; gej_set_infinity(acc);
                uint32_t count = s == 0 ? Q : Q / 2, lo = part * MSM_R2_ELEMS, hi = lo + MSM_R2_ELEMS < count ? lo + MSM_R2_ELEMS : count;
                for (uint32_t k = lo; k < hi; k++) { gej o, t; msm_load_gej(o, s == 0 ? Lq[cb + k] : Aq[cb + msm_bit_index(k, s - 1)]); gej_add(t, acc, o); acc = t; }
                gej t; gej_add(t, tot, acc); tot = t;
            }
            msm_store_gej(T[((size_t)wi * MSM_MAX_SUMS + s) * MSM_R2_PARTS], tot);
        }
        uint32_t maxbits = msm_ilog2((p.B > p.Btop ? p.B : p.Btop) / MSM_CHUNK);
        msm_window_from_sums(W[wi], T.data() + (size_t)wi * MSM_MAX_SUMS * MSM_R2_PARTS, MSM_R2_PARTS, nbits);
    }
    gej H[2];
    msm_horner_half(H[0], W.data(), p);
    msm_horner_half(H[1], W.data() + p.nw, p);
    msm_join_halves(r, H[0], H[1]);
}
int emu_msm2(uint8_t* out64, const uint8_t* g_sc32, const uint8_t* scalars32, const uint8_t* points64, size_t n, int c, uint32_t tp) {
    gej r;
    emu_msm_core(r, g_sc32, scalars32, points64, n, c, tp, 0, 1);
    return gej_to_xy64(out64, r);
}
// window sharding (secp256k1_b200_ecmult_multi_window_shard_dev): every shard over all terms, the partial results added
int emu_msm_window_shards(uint8_t* out64, const uint8_t* g_sc32, const uint8_t* scalars32, const uint8_t* points64, size_t n, int c, uint32_t shards) {
    gej tot; gej_set_infinity(tot);
    for (uint32_t s = 0; s < shards; s++) {
        gej r, t;
        emu_msm_core(r, g_sc32, scalars32, points64, n, c, 0, s, shards);
        gej_add(t, tot, r); tot = t;
    }
    return gej_to_xy64(out64, tot);
}
int emu_msm(uint8_t* out64, const uint8_t* g_sc32, const uint8_t* scalars32, const uint8_t* points64, size_t n) { return emu_msm2(out64, g_sc32, scalars32, points64, n, 0, 0); }
int emu_msm_choose_c(size_t N) { return (int)msm_choose_c(N); }
// k_bppp_norm_verify lane by lane; mode 0 = plain SHA-256 transcript, 1 = tests' initial-data transcript
int emu_bppp_norm_verify(const uint8_t* proof, size_t plen, const uint8_t* commit64, const uint8_t* rho32, const uint8_t* gens64, uint32_t g_len,
                         const uint8_t* c_vec32, uint32_t h_len, int mode) {
    sha256 tr; memset(tr.buf, 0, 64);
    if (mode == 0) sha256_init(tr);
    else {
        sha256_init_tagged(tr, (const uint8_t*)"Bulletproofs_pp/v0/commitment", 29);
        uint8_t ser[33], le[8];
        ser[0] = (commit64[63] & 1) ? 3 : 2; memcpy(ser + 1, commit64, 32); sha256_write(tr, ser, 33);
        sha256_write(tr, rho32, 32);
        for (int k = 0; k < 8; k++) le[k] = (uint8_t)((uint64_t)g_len >> (8 * k)); sha256_write(tr, le, 8);
        for (int k = 0; k < 8; k++) le[k] = (uint8_t)((uint64_t)(g_len + h_len) >> (8 * k)); sha256_write(tr, le, 8);
        for (uint32_t i = 0; i < g_len + h_len; i++) { ser[0] = (gens64[64 * i + 63] & 1) ? 3 : 2; memcpy(ser + 1, gens64 + 64 * i, 32); sha256_write(tr, ser, 33); }
        for (int k = 0; k < 8; k++) le[k] = (uint8_t)((uint64_t)h_len >> (8 * k)); sha256_write(tr, le, 8);
        sha256_write(tr, c_vec32, 32 * (size_t)h_len);
    }
    sc rho; if (sc_set_b32(rho, rho32)) return 0;
    bppp_ctx c;
    if (!bppp_prepare(c, proof, plen, tr, rho, g_len, h_len)) return 0;
    gej total; gej_set_infinity(total);
    uint32_t T = bppp_num_terms(c);
    sc zero; sc_set_zero(zero);
    for (uint32_t t = 0; t < T; t++) {
        ge pt; sc s;
        if (!bppp_term(pt, s, c, t, proof, gens64)) return 0;
        if (pt.inf) continue;
        gej pj, r, sum; gej_set_ge(pj, pt);
        tab_local tab; ecmult(r, pj, s, zero, g_gt, tab);
        gej_add(sum, total, r); total = sum;
    }
    sc v; bppp_v(v, c, c_vec32); sc_negate(v, v);
    gej inf, r, sum; gej_set_infinity(inf);
    tab_local tab; ecmult(r, inf, zero, v, g_gt, tab);
    gej_add(sum, total, r); total = sum;
    ge cm; ge_from_xy64(cm, commit64);
    if (!cm.inf) { gej_add_ge(sum, total, cm); total = sum; }
    return total.inf;
}
int emu_surjection_verify(const uint8_t* proof, size_t plen, const uint8_t* in_tags64, uint32_t n_tags, const uint8_t* out_tag64) {
    tab_local tab; return surjection_verify_item(proof, plen, in_tags64, n_tags, out_tag64, g_gt, tab);
}
int emu_pedersen_tally(const uint8_t* pos33, uint32_t pcnt, const uint8_t* neg33, uint32_t ncnt) { return pedersen_tally_item(pos33, pcnt, neg33, ncnt); }
void emu_halfagg_z_chain(uint8_t* z32, const uint8_t* pk32, const uint8_t* msgs32, size_t n, const uint8_t* aggsig) {
    halfagg_midstate hm; sha256 h;
    sha256_init_tagged(h, (const uint8_t*)"HalfAgg/randomizer", 18);
    for (int i = 0; i < 8; i++) hm.s[i] = h.s[i];
    halfagg_z_chain(z32, hm, aggsig, pk32, msgs32, n);
}
int emu_halfagg_verify(const uint8_t* pk32, const uint8_t* msgs32, size_t n, const uint8_t* aggsig, size_t aggsig_len) {
    halfagg_midstate hm; bip340_midstate bm; sha256 h;
    sha256_init_tagged(h, (const uint8_t*)"HalfAgg/randomizer", 18);
    for (int i = 0; i < 8; i++) hm.s[i] = h.s[i];
    sha256_init_tagged(h, (const uint8_t*)"BIP0340/challenge", 17);
    for (int i = 0; i < 8; i++) bm.s[i] = h.s[i];
    std::vector<uint8_t> z(32 * n + 32);
    tab_local tab; return halfagg_verify_serial(hm, bm, pk32, msgs32, n, aggsig, aggsig_len, z.data(), g_gt, tab);
}
int emu_ecdsa_recover(uint8_t* out64, const uint8_t* sig64, int recid, const uint8_t* msg32) {
    if (recid < 0 || recid > 3) { memset(out64, 0, 64); return 0; }
    tab_local tab; return ecdsa_recover_item(out64, sig64, recid, msg32, g_gt, tab);
}
int emu_ecdsa_verify(const uint8_t* sig, int sigfmt, const uint8_t* msg32, const uint8_t* pk, int pkfmt) {
    tab_local tab; return ecdsa_verify_item(sig, sigfmt, msg32, pk, pkfmt, g_gt, tab);
}
}
