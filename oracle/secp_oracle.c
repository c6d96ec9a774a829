/* oracle/secp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's algorithms for the accelerated hot path (see secp_oracle.h).
 * Written independently of the CUDA code: 4x64-bit limbs with unsigned __int128, always fully reduced
 * values, simple windowed double-and-add -- the group elements and accept/reject decisions it produces are
 * the unique mathematical results the reference's own (5x52-limb, wNAF / Pippenger) code produces, so
 * serialized outputs are comparable byte for byte.  Every function cites the reference file:line whose
 * behaviour it restates (paths relative to the reference tree).
 */
#include "secp_oracle.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } u256;   /* little-endian limbs */

/* ------------------------------------------------------------------------------------------------ */
/* 256-bit helpers                                                                                     */
static const u256 FE_P = {{0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL}};
static const u256 SC_N = {{0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL}};
/* 2^256 - p and 2^256 - n as 3-limb numbers */
static const uint64_t FE_C[3] = {0x1000003D1ULL, 0, 0};
static const uint64_t SC_C[3] = {0x402DA1732FC9BEBFULL, 0x4551231950B75FC4ULL, 1};

static void u256_from_be(u256 *r, const unsigned char *b) {
    int i, j;
    for (i = 0; i < 4; i++) {
        uint64_t w = 0;
        for (j = 0; j < 8; j++) w = (w << 8) | b[8 * (3 - i) + j];
        r->v[i] = w;
    }
}
static void u256_to_be(unsigned char *b, const u256 *a) {
    int i, j;
    for (i = 0; i < 4; i++) for (j = 0; j < 8; j++) b[8 * (3 - i) + j] = (unsigned char)(a->v[i] >> (56 - 8 * j));
}
static int u256_cmp(const u256 *a, const u256 *b) {
    int i;
    for (i = 3; i >= 0; i--) { if (a->v[i] < b->v[i]) return -1; if (a->v[i] > b->v[i]) return 1; }
    return 0;
}
static int u256_is_zero(const u256 *a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static uint64_t u256_add(u256 *r, const u256 *a, const u256 *b) {
    u128 c = 0; int i;
    for (i = 0; i < 4; i++) { c += (u128)a->v[i] + b->v[i]; r->v[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static uint64_t u256_sub(u256 *r, const u256 *a, const u256 *b) {
    uint64_t borrow = 0; int i;
    for (i = 0; i < 4; i++) {
        uint64_t t = a->v[i] - b->v[i];
        uint64_t b2 = (a->v[i] < b->v[i]) | (t < borrow);
        r->v[i] = t - borrow; borrow = b2;
    }
    return borrow;
}
/* r = a*b mod m, where m = 2^256 - c (c given as 3 limbs, < 2^130) */
static void mod_reduce512(u256 *r, const uint64_t t[8], const uint64_t c[3], const u256 *m) {
    uint64_t w[8]; int round, i, j;
    memcpy(w, t, sizeof(w));
    for (round = 0; round < 3; round++) {
        /* w = lo + hi * c */
        uint64_t acc[8] = {0};
        u128 carry;
        for (i = 0; i < 4; i++) acc[i] = w[i];
        for (i = 0; i < 4; i++) {
            carry = 0;
            for (j = 0; j < 3; j++) {
                u128 p = (u128)w[4 + i] * c[j] + acc[i + j] + carry;
                acc[i + j] = (uint64_t)p; carry = p >> 64;
            }
            for (j = i + 3; carry && j < 8; j++) { u128 p = (u128)acc[j] + carry; acc[j] = (uint64_t)p; carry = p >> 64; }
        }
        memcpy(w, acc, sizeof(w));
    }
    /* now w < 2^256 + small: hi part is 0 after three folds (values shrink 512 -> 386 -> 260 -> 256+) */
    {
        u256 x = {{w[0], w[1], w[2], w[3]}};
        /* w[4] can still be 1 in rare cases: fold once more by adding c */
        while (w[4]) {
            u256 cc = {{c[0], c[1], c[2], 0}};
            w[4] -= 1;
            w[4] += u256_add(&x, &x, &cc);
        }
        while (u256_cmp(&x, m) >= 0) u256_sub(&x, &x, m);
        *r = x;
    }
}
static void mulmod(u256 *r, const u256 *a, const u256 *b, const uint64_t c[3], const u256 *m) {
    uint64_t t[8] = {0}; int i, j;
    for (i = 0; i < 4; i++) {
        u128 carry = 0;
        for (j = 0; j < 4; j++) { u128 p = (u128)a->v[i] * b->v[j] + t[i + j] + carry; t[i + j] = (uint64_t)p; carry = p >> 64; }
        t[i + 4] = (uint64_t)carry;
    }
    mod_reduce512(r, t, c, m);
}
static void addmod(u256 *r, const u256 *a, const u256 *b, const u256 *m) {
    uint64_t c = u256_add(r, a, b);
    if (c || u256_cmp(r, m) >= 0) u256_sub(r, r, m);
}
static void submod(u256 *r, const u256 *a, const u256 *b, const u256 *m) {
    if (u256_sub(r, a, b)) u256_add(r, r, m);
}
static void powmod(u256 *r, const u256 *a, const u256 *e, const uint64_t c[3], const u256 *m) {
    u256 acc = {{1, 0, 0, 0}}; int i;
    for (i = 255; i >= 0; i--) {
        mulmod(&acc, &acc, &acc, c, m);
        if ((e->v[i >> 6] >> (i & 63)) & 1) mulmod(&acc, &acc, a, c, m);
    }
    *r = acc;
}

/* ---- field: field_5x52_int128_impl.h:18 (mul), :154 (sqr), field_impl.h:37 (sqrt), field_5x52_impl.h:491 (inv) */
typedef u256 fe;
static void fe_mul(fe *r, const fe *a, const fe *b) { mulmod(r, a, b, FE_C, &FE_P); }
static void fe_sqr(fe *r, const fe *a) { mulmod(r, a, a, FE_C, &FE_P); }
static void fe_add(fe *r, const fe *a, const fe *b) { addmod(r, a, b, &FE_P); }
static void fe_sub(fe *r, const fe *a, const fe *b) { submod(r, a, b, &FE_P); }
static void fe_neg(fe *r, const fe *a) { fe z = {{0, 0, 0, 0}}; submod(r, &z, a, &FE_P); }
static void fe_set_int(fe *r, uint64_t a) { r->v[0] = a; r->v[1] = r->v[2] = r->v[3] = 0; }
static int fe_is_odd(const fe *a) { return (int)(a->v[0] & 1); }
/* field_5x52_impl.h:265 set_b32_limit: fails if >= p ; :228 set_b32_mod reduces */
static int fe_set_b32_limit(fe *r, const unsigned char *b) { u256_from_be(r, b); return u256_cmp(r, &FE_P) < 0; }
static void fe_set_b32_mod(fe *r, const unsigned char *b) { u256_from_be(r, b); if (u256_cmp(r, &FE_P) >= 0) u256_sub(r, r, &FE_P); }
static void fe_inv(fe *r, const fe *a) {
    u256 e = FE_P; e.v[0] -= 2;
    powmod(r, a, &e, FE_C, &FE_P);
}
/* field_impl.h:37-146: r = a^((p+1)/4), returns whether r^2 == a */
static int fe_sqrt(fe *r, const fe *a) {
    u256 e = {{0xFFFFFFFFBFFFFF0CULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0x3FFFFFFFFFFFFFFFULL}};
    fe t, c;
    powmod(&t, a, &e, FE_C, &FE_P);
    fe_sqr(&c, &t);
    *r = t;
    return u256_cmp(&c, a) == 0;
}
/* field_5x52_impl.h:501 is_square_var (Jacobi symbol >= 0) */
static int fe_is_square(const fe *a) { fe r; return fe_sqrt(&r, a); }

/* ---- scalar: scalar_4x64_impl.h:155 set_b32, :1031 mul, :1162 inverse_var, :252 is_high */
typedef u256 sc;
static int sc_set_b32(sc *r, const unsigned char *b) {
    u256_from_be(r, b);
    if (u256_cmp(r, &SC_N) >= 0) { u256_sub(r, r, &SC_N); return 1; }
    return 0;
}
static void sc_mul(sc *r, const sc *a, const sc *b) { mulmod(r, a, b, SC_C, &SC_N); }
static void sc_add(sc *r, const sc *a, const sc *b) { addmod(r, a, b, &SC_N); }
static void sc_neg(sc *r, const sc *a) { sc z = {{0, 0, 0, 0}}; submod(r, &z, a, &SC_N); }
static void sc_inv(sc *r, const sc *a) { u256 e = SC_N; e.v[0] -= 2; powmod(r, a, &e, SC_C, &SC_N); }
static int sc_is_high(const sc *a) {
    const u256 h = {{0xDFE92F46681B20A0ULL, 0x5D576E7357A4501DULL, 0xFFFFFFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFFFFULL}};
    return u256_cmp(a, &h) > 0;
}
static void sc_set_u64(sc *r, uint64_t a) { r->v[0] = a; r->v[1] = r->v[2] = r->v[3] = 0; }

/* ---- SHA-256: hash_impl.h:51-194 ------------------------------------------------------------------ */
typedef struct { uint32_t s[8]; unsigned char buf[64]; uint64_t bytes; } sha256_t;
static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3,
    0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
static void sha_block(sha256_t *h) {
    uint32_t w[64], a, b, c, d, e, f, g, hh; int i;
    for (i = 0; i < 16; i++) w[i] = ((uint32_t)h->buf[4 * i] << 24) | ((uint32_t)h->buf[4 * i + 1] << 16) | ((uint32_t)h->buf[4 * i + 2] << 8) | h->buf[4 * i + 3];
    for (i = 16; i < 64; i++) {
        uint32_t s0 = ROR(w[i - 15], 7) ^ ROR(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = ROR(w[i - 2], 17) ^ ROR(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    a = h->s[0]; b = h->s[1]; c = h->s[2]; d = h->s[3]; e = h->s[4]; f = h->s[5]; g = h->s[6]; hh = h->s[7];
    for (i = 0; i < 64; i++) {
        uint32_t t1 = hh + (ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25)) + ((e & f) ^ (~e & g)) + SHA_K[i] + w[i];
        uint32_t t2 = (ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h->s[0] += a; h->s[1] += b; h->s[2] += c; h->s[3] += d; h->s[4] += e; h->s[5] += f; h->s[6] += g; h->s[7] += hh;
}
static void sha_init(sha256_t *h) {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(h->s, iv, sizeof(iv)); h->bytes = 0;
}
static void sha_write(sha256_t *h, const unsigned char *d, size_t len) {
    size_t fill = h->bytes & 63, i;
    h->bytes += len;
    for (i = 0; i < len; i++) { h->buf[fill++] = d[i]; if (fill == 64) { sha_block(h); fill = 0; } }
}
static void sha_final(sha256_t *h, unsigned char *out) {
    uint64_t bits = h->bytes * 8; size_t fill = h->bytes & 63; int i;
    h->buf[fill++] = 0x80;
    if (fill > 56) { while (fill < 64) h->buf[fill++] = 0; sha_block(h); fill = 0; }
    while (fill < 56) h->buf[fill++] = 0;
    for (i = 0; i < 8; i++) h->buf[56 + i] = (unsigned char)(bits >> (56 - 8 * i));
    sha_block(h);
    for (i = 0; i < 8; i++) { out[4 * i] = h->s[i] >> 24; out[4 * i + 1] = h->s[i] >> 16; out[4 * i + 2] = h->s[i] >> 8; out[4 * i + 3] = h->s[i]; }
}
/* hash_impl.h:196-205 */
static void sha_init_tagged(sha256_t *h, const char *tag) {
    unsigned char th[32];
    sha_init(h); sha_write(h, (const unsigned char *)tag, strlen(tag)); sha_final(h, th);
    sha_init(h); sha_write(h, th, 32); sha_write(h, th, 32);
}

/* ---- group: group_impl.h:468 (double), :534 (add_var), :177 (to affine), :347 (set_xquad) ------------- */
typedef struct { fe x, y; int inf; } ge;
typedef struct { fe x, y, z; int inf; } gej;
static const ge GEN_G = {{{0x59F2815B16F81798ULL, 0x029BFCDB2DCE28D9ULL, 0x55A06295CE870B07ULL, 0x79BE667EF9DCBBACULL}},
                         {{0x9C47D08FFB10D4B8ULL, 0xFD17B448A6855419ULL, 0x5DA4FBFC0E1108A8ULL, 0x483ADA7726A3C465ULL}}, 0};
static void gej_set_ge(gej *r, const ge *a) { r->x = a->x; r->y = a->y; fe_set_int(&r->z, 1); r->inf = a->inf; }
static void gej_set_inf(gej *r) { memset(r, 0, sizeof(*r)); r->inf = 1; }
static void gej_neg(gej *r, const gej *a) { *r = *a; fe_neg(&r->y, &a->y); }
static void ge_neg(ge *r, const ge *a) { *r = *a; fe_neg(&r->y, &a->y); }
static void gej_double(gej *r, const gej *a) {
    /* dbl-2009-l for a = 0 */
    fe A, B, C, D, E, F, t;
    if (a->inf) { gej_set_inf(r); return; }
    fe_sqr(&A, &a->x); fe_sqr(&B, &a->y); fe_sqr(&C, &B);
    fe_add(&t, &a->x, &B); fe_sqr(&t, &t); fe_sub(&t, &t, &A); fe_sub(&t, &t, &C); fe_add(&D, &t, &t);
    fe_add(&E, &A, &A); fe_add(&E, &E, &A);
    fe_sqr(&F, &E);
    fe_mul(&t, &a->y, &a->z);
    fe_add(&r->z, &t, &t);
    fe_sub(&r->x, &F, &D); fe_sub(&r->x, &r->x, &D);
    fe_sub(&t, &D, &r->x); fe_mul(&t, &t, &E);
    fe_add(&C, &C, &C); fe_add(&C, &C, &C); fe_add(&C, &C, &C);
    fe_sub(&r->y, &t, &C);
    r->inf = 0;
}
static void gej_add(gej *r, const gej *a, const gej *b) {
    fe z1z1, z2z2, u1, u2, s1, s2, h, rr, hh, hhh, v, t;
    if (a->inf) { *r = *b; return; }
    if (b->inf) { *r = *a; return; }
    fe_sqr(&z1z1, &a->z); fe_sqr(&z2z2, &b->z);
    fe_mul(&u1, &a->x, &z2z2); fe_mul(&u2, &b->x, &z1z1);
    fe_mul(&s1, &a->y, &b->z); fe_mul(&s1, &s1, &z2z2);
    fe_mul(&s2, &b->y, &a->z); fe_mul(&s2, &s2, &z1z1);
    fe_sub(&h, &u2, &u1); fe_sub(&rr, &s2, &s1);
    if (u256_is_zero(&h)) {
        if (u256_is_zero(&rr)) gej_double(r, a); else gej_set_inf(r);
        return;
    }
    fe_sqr(&hh, &h); fe_mul(&hhh, &hh, &h); fe_mul(&v, &u1, &hh);
    fe_sqr(&t, &rr); fe_sub(&t, &t, &hhh); fe_sub(&t, &t, &v); fe_sub(&t, &t, &v);
    fe_sub(&v, &v, &t); fe_mul(&v, &v, &rr); fe_mul(&hhh, &hhh, &s1);
    fe_mul(&u1, &a->z, &b->z); fe_mul(&r->z, &u1, &h);
    fe_sub(&r->y, &v, &hhh);
    r->x = t; r->inf = 0;
}
static void gej_add_ge(gej *r, const gej *a, const ge *b) { gej bj; gej_set_ge(&bj, b); gej_add(r, a, &bj); }
static void ge_set_gej(ge *r, const gej *a) {
    fe zi, zi2, zi3;
    if (a->inf) { memset(r, 0, sizeof(*r)); r->inf = 1; return; }
    fe_inv(&zi, &a->z); fe_sqr(&zi2, &zi); fe_mul(&zi3, &zi2, &zi);
    fe_mul(&r->x, &a->x, &zi2); fe_mul(&r->y, &a->y, &zi3); r->inf = 0;
}
static int ge_set_xquad(ge *r, const fe *x) {
    fe c, seven;
    fe_sqr(&c, x); fe_mul(&c, &c, x); fe_set_int(&seven, 7); fe_add(&c, &c, &seven);
    r->x = *x; r->inf = 0;
    return fe_sqrt(&r->y, &c);
}
static int ge_set_xo(ge *r, const fe *x, int odd) {
    if (!ge_set_xquad(r, x)) return 0;
    if (fe_is_odd(&r->y) != odd) fe_neg(&r->y, &r->y);
    return 1;
}
static int ge_is_valid(const ge *a) {
    fe y2, x3, seven;
    if (a->inf) return 0;
    fe_sqr(&y2, &a->y); fe_sqr(&x3, &a->x); fe_mul(&x3, &x3, &a->x); fe_set_int(&seven, 7); fe_add(&x3, &x3, &seven);
    return u256_cmp(&y2, &x3) == 0;
}
static void ge_from_xy64(ge *r, const unsigned char *xy) {
    int i, z = 1;
    for (i = 0; i < 64; i++) z &= (xy[i] == 0);
    if (z) { memset(r, 0, sizeof(*r)); r->inf = 1; return; }
    fe_set_b32_mod(&r->x, xy); fe_set_b32_mod(&r->y, xy + 32); r->inf = 0;
}
static int gej_to_xy64(unsigned char *xy, const gej *a) {
    ge p;
    if (a->inf) { memset(xy, 0, 64); return 0; }
    ge_set_gej(&p, a);
    u256_to_be(xy, &p.x); u256_to_be(xy + 32, &p.y);
    return 1;
}
/* eckey_impl.h:38-45 */
static void ge_serialize33(unsigned char *out, const ge *a) { out[0] = fe_is_odd(&a->y) ? 3 : 2; u256_to_be(out + 1, &a->x); }

/* ---- ecmult: the VALUE of secp256k1_ecmult (ecmult_impl.h:365): r = na*a + ng*G ------------------------ */
static void gej_mul(gej *r, const gej *a, const sc *k) {
    /* 4-bit fixed window double-and-add */
    gej tab[16], acc; int i, j;
    gej_set_inf(&tab[0]); tab[1] = *a;
    for (i = 2; i < 16; i++) gej_add(&tab[i], &tab[i - 1], a);
    gej_set_inf(&acc);
    for (i = 63; i >= 0; i--) {
        for (j = 0; j < 4; j++) gej_double(&acc, &acc);
        gej_add(&acc, &acc, &tab[(k->v[i >> 4] >> ((i & 15) * 4)) & 15]);
    }
    *r = acc;
}
static void ecmult(gej *r, const gej *a, const sc *na, const sc *ng) {
    gej t1, t2, g;
    gej_mul(&t1, a, na);
    if (ng) { gej_set_ge(&g, &GEN_G); gej_mul(&t2, &g, ng); gej_add(r, &t1, &t2); } else *r = t1;
}

/* ------------------------------------------------------------------------------------------------ */
/* exported byte-level wrappers                                                                        */
void orc_fe_mul(unsigned char *r32, const unsigned char *a32, const unsigned char *b32) { fe a, b, r; fe_set_b32_mod(&a, a32); fe_set_b32_mod(&b, b32); fe_mul(&r, &a, &b); u256_to_be(r32, &r); }
void orc_fe_inv(unsigned char *r32, const unsigned char *a32) { fe a, r; fe_set_b32_mod(&a, a32); fe_inv(&r, &a); u256_to_be(r32, &r); }
int orc_fe_sqrt(unsigned char *r32, const unsigned char *a32) { fe a, r; int ok; fe_set_b32_mod(&a, a32); ok = fe_sqrt(&r, &a); u256_to_be(r32, &r); return ok; }
void orc_scalar_mul(unsigned char *r32, const unsigned char *a32, const unsigned char *b32) { sc a, b, r; sc_set_b32(&a, a32); sc_set_b32(&b, b32); sc_mul(&r, &a, &b); u256_to_be(r32, &r); }
void orc_scalar_inverse(unsigned char *r32, const unsigned char *a32) { sc a, r; sc_set_b32(&a, a32); sc_inv(&r, &a); u256_to_be(r32, &r); }
void orc_sha256(unsigned char *out32, const unsigned char *data, size_t len) { sha256_t h; sha_init(&h); sha_write(&h, data, len); sha_final(&h, out32); }

int orc_ecmult(unsigned char *r64, const unsigned char *a64, const unsigned char *na32, const unsigned char *ng32) {
    ge a; gej aj, r; sc na, ng;
    ge_from_xy64(&a, a64); gej_set_ge(&aj, &a);
    sc_set_b32(&na, na32); if (ng32) sc_set_b32(&ng, ng32);
    ecmult(&r, &aj, &na, ng32 ? &ng : NULL);
    return gej_to_xy64(r64, &r);
}
void orc_ecmult_batch(unsigned char *r64, const unsigned char *a64, const unsigned char *na32, const unsigned char *ng32, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) orc_ecmult(r64 + 64 * i, a64 + 64 * i, na32 + 32 * i, ng32 ? ng32 + 32 * i : NULL);
}
/* run_ecmult_chain (tests.c:4617-4674): X <- xn*X + gn*G; xn *= xf; gn *= gf */
int orc_ecmult_chain(unsigned char *r64, const unsigned char *a64, const unsigned char *xn32, const unsigned char *gn32,
                     const unsigned char *xf32, const unsigned char *gf32, size_t iterations) {
    ge a; gej x; sc xn, gn, xf, gf; size_t i;
    ge_from_xy64(&a, a64); gej_set_ge(&x, &a);
    sc_set_b32(&xn, xn32); sc_set_b32(&gn, gn32); sc_set_b32(&xf, xf32); sc_set_b32(&gf, gf32);
    for (i = 0; i < iterations; i++) {
        gej t; ecmult(&t, &x, &xn, &gn); x = t;
        sc_mul(&xn, &xn, &xf); sc_mul(&gn, &gn, &gf);
    }
    return gej_to_xy64(r64, &x);
}
/* secp256k1_ecmult_multi_var (ecmult_impl.h:823): r = g_sc*G + sum sc_i*pt_i -- stated as the simple
 * algorithm the reference itself falls back to (ecmult_multi_simple_var, :779-800) */
int orc_ecmult_multi(unsigned char *r64, const unsigned char *g_sc32, const unsigned char *sc32, const unsigned char *pt64, size_t n) {
    gej acc, t; size_t i;
    gej_set_inf(&acc);
    if (g_sc32) { sc g; gej gj; sc_set_b32(&g, g_sc32); gej_set_ge(&gj, &GEN_G); gej_mul(&acc, &gj, &g); }
    for (i = 0; i < n; i++) {
        ge p; gej pj; sc s;
        ge_from_xy64(&p, pt64 + 64 * i); gej_set_ge(&pj, &p); sc_set_b32(&s, sc32 + 32 * i);
        gej_mul(&t, &pj, &s); gej_add(&acc, &acc, &t);
    }
    gej_to_xy64(r64, &acc);
    return 1;
}

/* ---- keys: extrakeys/main_impl.h:22-42 (xonly parse), eckey_impl.h:18-36 (pubkey parse) ----------------- */
static int xonly_parse(ge *p, const unsigned char *pk32) {
    fe x;
    if (!fe_set_b32_limit(&x, pk32)) return 0;
    return ge_set_xo(p, &x, 0);
}
static int pubkey_parse(ge *p, const unsigned char *pk, size_t len) {
    fe x, y;
    if (len == 33 && (pk[0] == 2 || pk[0] == 3)) {
        return fe_set_b32_limit(&x, pk + 1) && ge_set_xo(p, &x, pk[0] == 3);
    } else if (len == 65 && (pk[0] == 4 || pk[0] == 6 || pk[0] == 7)) {
        if (!fe_set_b32_limit(&x, pk + 1) || !fe_set_b32_limit(&y, pk + 33)) return 0;
        p->x = x; p->y = y; p->inf = 0;
        if ((pk[0] == 6 || pk[0] == 7) && fe_is_odd(&y) != (pk[0] == 7)) return 0;
        return ge_is_valid(p);
    }
    return 0;
}
int orc_xonly_pubkey_parse(unsigned char *xy64, const unsigned char *pk32) {
    ge p; if (!xonly_parse(&p, pk32)) { memset(xy64, 0, 64); return 0; }
    u256_to_be(xy64, &p.x); u256_to_be(xy64 + 32, &p.y); return 1;
}
int orc_ec_pubkey_parse(unsigned char *xy64, const unsigned char *pk, size_t len) {
    ge p; if (!pubkey_parse(&p, pk, len)) { memset(xy64, 0, 64); return 0; }
    u256_to_be(xy64, &p.x); u256_to_be(xy64 + 32, &p.y); return 1;
}

/* ---- BIP-340 verify: modules/schnorrsig/main_impl.h:215-261 (+ challenge :106-120) ---------------------- */
int orc_schnorrsig_verify(const unsigned char *sig64, const unsigned char *msg, size_t msglen, const unsigned char *pk32) {
    fe rx; sc s, e; ge P, R; gej Pj, Rj; sha256_t h; unsigned char buf[32];
    if (!fe_set_b32_limit(&rx, sig64)) return 0;
    if (sc_set_b32(&s, sig64 + 32)) return 0;
    if (!xonly_parse(&P, pk32)) return 0;
    sha_init_tagged(&h, "BIP0340/challenge");
    sha_write(&h, sig64, 32); u256_to_be(buf, &P.x); sha_write(&h, buf, 32); sha_write(&h, msg, msglen);
    sha_final(&h, buf);
    sc_set_b32(&e, buf);
    sc_neg(&e, &e);
    gej_set_ge(&Pj, &P);
    ecmult(&Rj, &Pj, &e, &s);
    if (Rj.inf) return 0;
    ge_set_gej(&R, &Rj);
    if (fe_is_odd(&R.y)) return 0;
    return u256_cmp(&rx, &R.x) == 0;
}
void orc_schnorrsig_verify_batch(unsigned char *ok, const unsigned char *sig64, const unsigned char *msg32, const unsigned char *pk32, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) ok[i] = (unsigned char)orc_schnorrsig_verify(sig64 + 64 * i, msg32 + 32 * i, 32, pk32 + 32 * i);
}

/* ---- BIP-340 half-aggregation verify: modules/schnorrsig_halfagg/main_impl.h:108-196 (pubkeys as x-only serializations) --- */
int orc_schnorrsig_aggverify(const unsigned char *pk32, const unsigned char *msgs32, size_t n, const unsigned char *aggsig, size_t aggsig_len) {
    sha256_t hz; gej rhs, lhs, gj; sc s; size_t i;
    if (aggsig_len / 32 == 0 || aggsig_len / 32 - 1 != n || aggsig_len % 32 != 0) return 0;
    rhs.inf = 1;
    sha_init_tagged(&hz, "HalfAgg/randomizer");
    for (i = 0; i < n; i++) {
        fe rx; ge rp, pp; sc ei, zi; gej ppj, ti, t2; sha256_t hc, he; unsigned char buf[32], pk_ser[32];
        if (!xonly_parse(&pp, pk32 + 32 * i)) return 0;
        u256_to_be(pk_ser, &pp.x);
        sha_write(&hz, aggsig + 32 * i, 32); sha_write(&hz, pk_ser, 32); sha_write(&hz, msgs32 + 32 * i, 32);
        hc = hz; sha_final(&hc, buf); sc_set_b32(&zi, buf);
        if (!fe_set_b32_limit(&rx, aggsig + 32 * i)) return 0;
        if (!ge_set_xo(&rp, &rx, 0)) return 0;
        sha_init_tagged(&he, "BIP0340/challenge");
        sha_write(&he, aggsig + 32 * i, 32); sha_write(&he, pk_ser, 32); sha_write(&he, msgs32 + 32 * i, 32);
        sha_final(&he, buf); sc_set_b32(&ei, buf);
        gej_set_ge(&ppj, &pp);
        ecmult(&ti, &ppj, &ei, NULL);
        gej_add_ge(&t2, &ti, &rp); ti = t2;
        if (i != 0) { ecmult(&t2, &ti, &zi, NULL); ti = t2; }
        gej_add(&t2, &rhs, &ti); rhs = t2;
    }
    if (sc_set_b32(&s, aggsig + 32 * n)) return 0;
    gej_set_ge(&gj, &GEN_G);
    gej_mul(&lhs, &gj, &s);
    gej_neg(&lhs, &lhs);
    gej_add(&gj, &lhs, &rhs);
    return gj.inf;
}

/* ---- ECDSA verify: secp256k1.c:498-512 + ecdsa_impl.h:195-272 (compact sig: secp256k1.c:372-391) -------- */
int orc_ecdsa_verify(const unsigned char *sig64, const unsigned char *msghash32, const unsigned char *pk, size_t pklen) {
    sc r, s, m, sn, u1, u2; ge P, R; gej Pj, Rj; sc xr;
    if (sc_set_b32(&r, sig64) || sc_set_b32(&s, sig64 + 32)) return 0;
    sc_set_b32(&m, msghash32);
    if (sc_is_high(&s)) return 0;
    if (!pubkey_parse(&P, pk, pklen)) return 0;
    if (u256_is_zero(&r) || u256_is_zero(&s)) return 0;
    sc_inv(&sn, &s); sc_mul(&u1, &sn, &m); sc_mul(&u2, &sn, &r);
    gej_set_ge(&Pj, &P);
    ecmult(&Rj, &Pj, &u2, &u1);
    if (Rj.inf) return 0;
    ge_set_gej(&R, &Rj);
    /* x(R) mod n == r */
    xr = R.x;
    if (u256_cmp(&xr, &SC_N) >= 0) u256_sub(&xr, &xr, &SC_N);
    return u256_cmp(&xr, &r) == 0;
}

/* ---- ECDSA public-key recovery: modules/recovery/main_impl.h:38-59 (parse_compact), :87-121 (sig_recover), :137-159 ---- */
int orc_ecdsa_recover(unsigned char *pubkey64, const unsigned char *sig64, int recid, const unsigned char *msghash32) {
    sc r, s, m, rn, u1, u2; fe fx; ge X, Q; gej Xj, Qj;
    memset(pubkey64, 0, 64);
    if (recid < 0 || recid > 3) return 0;
    if (sc_set_b32(&r, sig64) || sc_set_b32(&s, sig64 + 32)) return 0;
    sc_set_b32(&m, msghash32);
    if (u256_is_zero(&r) || u256_is_zero(&s)) return 0;
    fx = r;
    if (recid & 2) {
        u256 pmn; u256_sub(&pmn, &FE_P, &SC_N);
        if (u256_cmp(&fx, &pmn) >= 0) return 0;
        u256_add(&fx, &fx, &SC_N);
    }
    if (!ge_set_xo(&X, &fx, recid & 1)) return 0;
    sc_inv(&rn, &r); sc_mul(&u1, &rn, &m); sc_neg(&u1, &u1); sc_mul(&u2, &rn, &s);
    gej_set_ge(&Xj, &X);
    ecmult(&Qj, &Xj, &u2, &u1);
    if (Qj.inf) return 0;
    ge_set_gej(&Q, &Qj);
    u256_to_be(pubkey64, &Q.x); u256_to_be(pubkey64 + 32, &Q.y);
    return 1;
}

/* ---- generator / Pedersen: modules/generator/main_impl.h:40-49, :67-92, :266-335, pedersen_impl.h:20-49 --- */
static void generator_load(ge *g, const unsigned char *gen64) { fe_set_b32_mod(&g->x, gen64); fe_set_b32_mod(&g->y, gen64 + 32); g->inf = 0; }
int orc_generator_parse(unsigned char *gen64, const unsigned char *in33) {
    fe x; ge g;
    if ((in33[0] & 0xFE) != 10 || !fe_set_b32_limit(&x, in33 + 1) || !ge_set_xquad(&g, &x)) return 0;
    if (in33[0] & 1) ge_neg(&g, &g);
    u256_to_be(gen64, &g.x); u256_to_be(gen64 + 32, &g.y);
    return 1;
}
/* commitment wire form: tag 8/9 = squareness of y (main_impl.h:275-297) */
static int commitment_parse(ge *c, const unsigned char *in33) {
    fe x;
    if ((in33[0] & 0xFE) != 8 || !fe_set_b32_limit(&x, in33 + 1) || !ge_set_xquad(c, &x)) return 0;
    if (in33[0] & 1) ge_neg(c, c);
    return 1;
}
int orc_pedersen_commitment_parse(unsigned char *xy64, const unsigned char *commit33) {
    ge c; if (!commitment_parse(&c, commit33)) { memset(xy64, 0, 64); return 0; }
    u256_to_be(xy64, &c.x); u256_to_be(xy64 + 32, &c.y); return 1;
}
int orc_pedersen_commit(unsigned char *commit33, const unsigned char *blind32, uint64_t value, const unsigned char *gen64) {
    ge genp, r; gej rj, vj, gj; sc sec, v;
    generator_load(&genp, gen64);
    if (sc_set_b32(&sec, blind32)) { memset(commit33, 0, 33); return 0; }
    gej_set_ge(&gj, &GEN_G); gej_mul(&rj, &gj, &sec);
    sc_set_u64(&v, value); gej_set_ge(&gj, &genp); gej_mul(&vj, &gj, &v);
    gej_add(&rj, &rj, &vj);
    if (rj.inf) { memset(commit33, 0, 33); return 0; }
    ge_set_gej(&r, &rj);
    commit33[0] = 9 ^ fe_is_square(&r.y);
    u256_to_be(commit33 + 1, &r.x);
    return 1;
}

/* ---- range proof verify: modules/rangeproof/rangeproof_impl.h:487-538 (header), :541-683 (verify),
 *      :20-51 (pub_expand), :53-59 (serialize_point); borromean_impl.h:23-37, :53-104 ------------------------ */
static void rp_serialize_point(unsigned char *out33, const ge *p) { out33[0] = !fe_is_square(&p->y); u256_to_be(out33 + 1, &p->x); }
static void borromean_hash(unsigned char *out, const unsigned char *m, size_t mlen, const unsigned char *e, size_t elen, size_t ridx, size_t eidx) {
    sha256_t h; unsigned char ring[4] = {ridx >> 24, ridx >> 16, ridx >> 8, ridx}, epos[4] = {eidx >> 24, eidx >> 16, eidx >> 8, eidx};
    sha_init(&h); sha_write(&h, e, elen); sha_write(&h, m, mlen); sha_write(&h, ring, 4); sha_write(&h, epos, 4); sha_final(&h, out);
}
/* ---- rewinding: hash_impl.h:211-314 (HMAC-SHA256, RFC 6979 generator), rangeproof_impl.h:61-112 (genrand), :334-350, :354-471 --- */
typedef struct { sha256_t inner, outer; } hmac_t;
static void hmac_init(hmac_t *h, const unsigned char *key, size_t keylen) {
    unsigned char rkey[64]; size_t n;
    memset(rkey, 0, 64); memcpy(rkey, key, keylen);            /* keylen <= 64 on this path */
    for (n = 0; n < 64; n++) rkey[n] ^= 0x5c;
    sha_init(&h->outer); sha_write(&h->outer, rkey, 64);
    for (n = 0; n < 64; n++) rkey[n] ^= 0x5c ^ 0x36;
    sha_init(&h->inner); sha_write(&h->inner, rkey, 64);
}
static void hmac_final(hmac_t *h, unsigned char *out32) {
    unsigned char t[32];
    sha_final(&h->inner, t); sha_write(&h->outer, t, 32); sha_final(&h->outer, out32);
}
typedef struct { unsigned char v[32], k[32]; int retry; } rfc6979_t;
static void rfc6979_init(rfc6979_t *r, const unsigned char *key, size_t keylen) {
    hmac_t h; unsigned char sep;
    memset(r->v, 1, 32); memset(r->k, 0, 32);
    for (sep = 0; sep < 2; sep++) {
        hmac_init(&h, r->k, 32); sha_write(&h.inner, r->v, 32); sha_write(&h.inner, &sep, 1); sha_write(&h.inner, key, keylen); hmac_final(&h, r->k);
        hmac_init(&h, r->k, 32); sha_write(&h.inner, r->v, 32); hmac_final(&h, r->v);
    }
    r->retry = 0;
}
static void rfc6979_gen32(rfc6979_t *r, unsigned char *out32) {
    hmac_t h; unsigned char zero = 0;
    if (r->retry) {
        hmac_init(&h, r->k, 32); sha_write(&h.inner, r->v, 32); sha_write(&h.inner, &zero, 1); hmac_final(&h, r->k);
        hmac_init(&h, r->k, 32); sha_write(&h.inner, r->v, 32); hmac_final(&h, r->v);
    }
    hmac_init(&h, r->k, 32); sha_write(&h.inner, r->v, 32); hmac_final(&h, r->v);
    memcpy(out32, r->v, 32);
    r->retry = 1;
}
static void rp_genrand(sc *sec, sc *s, unsigned char *message, const size_t *rsizes, size_t rings, const unsigned char *nonce,
                       const ge *commit, const unsigned char *proof, size_t len, const ge *genp) {
    unsigned char tmp[32], seed[32 + 33 + 33 + 10]; rfc6979_t rng; sc acc = {{0, 0, 0, 0}}; size_t i, j, npub = 0; int b, overflow;
    memcpy(seed, nonce, 32); rp_serialize_point(seed + 32, commit); rp_serialize_point(seed + 65, genp); memcpy(seed + 98, proof, len);
    rfc6979_init(&rng, seed, 98 + len);
    for (i = 0; i < rings; i++) {
        if (i < rings - 1) {
            rfc6979_gen32(&rng, tmp);
            do { rfc6979_gen32(&rng, tmp); overflow = sc_set_b32(&sec[i], tmp); } while (overflow || u256_is_zero(&sec[i]));
            sc_add(&acc, &acc, &sec[i]);
        } else sc_neg(&sec[i], &acc);
        for (j = 0; j < rsizes[i]; j++) {
            rfc6979_gen32(&rng, tmp);
            for (b = 0; b < 32; b++) { tmp[b] ^= message[(i * 4 + j) * 32 + b]; message[(i * 4 + j) * 32 + b] = tmp[b]; }
            sc_set_b32(&s[npub], tmp);
            npub++;
        }
    }
}
static void rp_recover_x(sc *x, const sc *k, const sc *e, const sc *s) { sc t; sc_neg(x, s); sc_add(x, x, k); sc_inv(&t, e); sc_mul(x, x, &t); }
static int rp_rewind_inner(sc *blind, uint64_t *v, unsigned char *m, size_t *mlen, const sc *ev, const sc *s, const size_t *rsizes, size_t rings,
                           const unsigned char *nonce, const ge *commit, const unsigned char *proof, size_t len, const ge *genp) {
    sc s_orig[128], sec[32], stmp; unsigned char prep[4096], tmp[32]; uint64_t value = 0; size_t offset, i, j, npub, skip1, skip2; int b;
    memset(prep, 0, 4096);
    rp_genrand(sec, s_orig, prep, rsizes, rings, nonce, commit, proof, len, genp);
    *v = UINT64_MAX;
    if (rings == 1 && rsizes[0] == 1) { rp_recover_x(blind, &s_orig[0], &ev[0], &s[0]); *v = 0; if (mlen) *mlen = 0; return 1; }
    npub = (rings - 1) << 2;
    for (j = 0; j < 2; j++) {
        size_t idx = npub + rsizes[rings - 1] - 1 - j;
        u256_to_be(tmp, &s[idx]);
        for (b = 0; b < 32; b++) tmp[b] ^= prep[idx * 32 + b];
        if ((tmp[0] & 128) && memcmp(tmp + 16, tmp + 24, 8) == 0 && memcmp(tmp + 8, tmp + 16, 8) == 0) {
            for (i = 0; i < 8; i++) value = (value << 8) + tmp[24 + i];
            *v = value; memcpy(prep + idx * 32, tmp, 32);
            break;
        }
    }
    if (j > 1) { if (mlen) *mlen = 0; return 0; }
    skip1 = rsizes[rings - 1] - 1 - j; skip2 = (value >> ((rings - 1) << 1)) & 3;
    if (skip1 == skip2) { if (mlen) *mlen = 0; return 0; }
    skip1 += (rings - 1) << 2; skip2 += (rings - 1) << 2;
    rp_recover_x(&stmp, &s_orig[skip2], &ev[skip2], &s[skip2]);
    sc_neg(&sec[rings - 1], &sec[rings - 1]);
    sc_add(blind, &stmp, &sec[rings - 1]);
    if (!m || !mlen || *mlen == 0) { if (mlen) *mlen = 0; return 1; }
    offset = 0; npub = 0;
    for (i = 0; i < rings; i++) {
        size_t idx = (value >> (i << 1)) & 3;
        for (j = 0; j < rsizes[i]; j++) {
            if (npub == skip1 || npub == skip2) { npub++; continue; }
            if (idx == j) { sc t; sc_mul(&t, &sec[i], &ev[npub]); sc_add(&stmp, &s[npub], &t); } else stmp = s[npub];
            u256_to_be(tmp, &stmp);
            for (b = 0; b < 32; b++) tmp[b] ^= prep[npub * 32 + b];
            for (b = 0; b < 32 && offset < *mlen; b++) m[offset++] = tmp[b];
            npub++;
        }
    }
    *mlen = offset;
    return 1;
}

static int rangeproof_verify_impl(unsigned char *blind_out, uint64_t *value_out, unsigned char *message_out, size_t *outlen, const unsigned char *nonce,
                          uint64_t *min_value, uint64_t *max_value, const unsigned char *commit33, const unsigned char *proof, size_t plen,
                          const unsigned char *extra, size_t extra_len, const unsigned char *gen64) {
    ge commit, genp, c; gej accj, pubs[128], base; sc s[128], evalues[128]; sha256_t sha_m, sha_e0; size_t offset_post_header;
    size_t rsizes[32], rings, npub, offset = 0, i, j, count; int exp, mantissa, has_nz, has_min; uint64_t scale;
    unsigned char signs[31], m[33], tmp[33]; const unsigned char *e0;
    if (!commitment_parse(&commit, commit33)) return 0;
    generator_load(&genp, gen64);
    /* header (:487-538) */
    if (plen < 65 || (proof[0] & 128)) return 0;
    has_nz = proof[0] & 64; has_min = proof[0] & 32; exp = -1; mantissa = 0;
    if (has_nz) {
        exp = proof[0] & 31; offset += 1;
        if (exp > 18) return 0;
        mantissa = proof[offset] + 1;
        if (mantissa > 64) return 0;
        *max_value = UINT64_MAX >> (64 - mantissa);
    } else *max_value = 0;
    offset += 1; scale = 1;
    for (i = 0; (int)i < exp; i++) { if (*max_value > UINT64_MAX / 10) return 0; *max_value *= 10; scale *= 10; }
    *min_value = 0;
    if (has_min) {
        if (plen - offset < 8) return 0;
        for (i = 0; i < 8; i++) *min_value = (*min_value << 8) | proof[offset + i];
        offset += 8;
    }
    if (*max_value > UINT64_MAX - *min_value) return 0;
    *max_value += *min_value;
    offset_post_header = offset;
    /* ring layout (:566-586) */
    rings = 1; rsizes[0] = 1; npub = 1;
    if (mantissa != 0) {
        rings = mantissa >> 1;
        for (i = 0; i < rings; i++) rsizes[i] = 4;
        npub = (mantissa >> 1) << 2;
        if (mantissa & 1) { rsizes[rings] = 2; npub += 2; rings++; }
    }
    if (plen - offset < 32 * (npub + rings - 1) + 32 + ((rings + 6) >> 3)) return 0;
    sha_init(&sha_m);
    rp_serialize_point(m, &commit); sha_write(&sha_m, m, 33);
    rp_serialize_point(m, &genp); sha_write(&sha_m, m, 33);
    sha_write(&sha_m, proof, offset);
    for (i = 0; i < rings - 1; i++) signs[i] = (proof[offset + (i >> 3)] & (1 << (i & 7))) != 0;
    offset += (rings + 6) >> 3;
    if ((rings - 1) & 7) { if ((proof[offset - 1] >> ((rings - 1) & 7)) != 0) return 0; }
    npub = 0;
    gej_set_inf(&accj);
    if (*min_value) { sc mv; gej gj; sc_set_u64(&mv, *min_value); gej_set_ge(&gj, &genp); gej_mul(&accj, &gj, &mv); }
    for (i = 0; i < rings - 1; i++) {
        fe x;
        if (!fe_set_b32_limit(&x, proof + offset) || !ge_set_xquad(&c, &x)) return 0;
        if (signs[i]) ge_neg(&c, &c);
        sha_write(&sha_m, &signs[i], 1); sha_write(&sha_m, proof + offset, 32);
        gej_set_ge(&pubs[npub], &c);
        gej_add_ge(&accj, &accj, &c);
        offset += 32; npub += rsizes[i];
    }
    gej_neg(&accj, &accj);
    gej_add_ge(&pubs[npub], &accj, &commit);
    if (pubs[npub].inf) return 0;
    /* pub_expand (:20-51) */
    { ge ng; ge_neg(&ng, &genp); gej_set_ge(&base, &ng); }
    { int e = exp < 0 ? 0 : exp; while (e--) { gej t2, t8; gej_double(&t2, &base); gej_double(&t8, &t2); gej_double(&t8, &t8); gej_add(&base, &t8, &t2); } }
    { size_t np = 0;
      for (i = 0; i < rings; i++) {
          for (j = 1; j < rsizes[i]; j++) gej_add(&pubs[np + j], &pubs[np + j - 1], &base);
          if (i < rings - 1) { gej_double(&base, &base); gej_double(&base, &base); }
          np += rsizes[i];
      } }
    npub += rsizes[rings - 1];
    e0 = proof + offset; offset += 32;
    for (i = 0; i < npub; i++) { if (sc_set_b32(&s[i], proof + offset)) return 0; offset += 32; }
    if (offset != plen) return 0;
    if (extra != NULL) sha_write(&sha_m, extra, extra_len);
    sha_final(&sha_m, m);
    /* borromean verify (borromean_impl.h:53-104) */
    count = 0; sha_init(&sha_e0);
    for (i = 0; i < rings; i++) {
        sc ens; int overflow;
        borromean_hash(tmp, m, 32, e0, 32, i, 0);
        overflow = sc_set_b32(&ens, tmp);
        for (j = 0; j < rsizes[i]; j++) {
            gej rgej; ge rge;
            if (overflow || u256_is_zero(&s[count]) || u256_is_zero(&ens) || pubs[count].inf) return 0;
            evalues[count] = ens;
            ecmult(&rgej, &pubs[count], &ens, &s[count]);
            if (rgej.inf) return 0;
            ge_set_gej(&rge, &rgej);
            ge_serialize33(tmp, &rge);
            if (j != rsizes[i] - 1) { borromean_hash(tmp, m, 32, tmp, 33, i, j + 1); overflow = sc_set_b32(&ens, tmp); }
            else sha_write(&sha_e0, tmp, 33);
            count++;
        }
    }
    sha_write(&sha_e0, m, 32); sha_final(&sha_e0, tmp);
    if (memcmp(e0, tmp, 32) != 0) return 0;
    if (nonce) {                                               /* rewind (:653-681) */
        sc blind, vs; uint64_t vv; gej bj, vj, gj;
        memset(&blind, 0, sizeof(blind));
        if (!rp_rewind_inner(&blind, &vv, message_out, outlen, evalues, s, rsizes, rings, nonce, &commit, proof, offset_post_header, &genp)) return 0;
        vv = vv * scale + *min_value;
        gej_set_ge(&gj, &GEN_G); gej_mul(&bj, &gj, &blind);
        sc_set_u64(&vs, vv); gej_set_ge(&gj, &genp); gej_mul(&vj, &gj, &vs);
        gej_add(&accj, &bj, &vj);
        if (accj.inf) return 0;
        gej_neg(&accj, &accj); gej_add_ge(&accj, &accj, &commit);
        if (!accj.inf) return 0;
        if (blind_out) u256_to_be(blind_out, &blind);
        if (value_out) *value_out = vv;
    }
    return 1;
}
int orc_rangeproof_verify(uint64_t *min_value, uint64_t *max_value, const unsigned char *commit33, const unsigned char *proof, size_t plen,
                          const unsigned char *extra, size_t extra_len, const unsigned char *gen64) {
    return rangeproof_verify_impl(NULL, NULL, NULL, NULL, NULL, min_value, max_value, commit33, proof, plen, extra, extra_len, gen64);
}
/* secp256k1_rangeproof_rewind (main_impl.h:31-52): *outlen = capacity in, bytes recovered out */
int orc_rangeproof_rewind(unsigned char *blind_out32, uint64_t *value_out, unsigned char *message_out, size_t *outlen, const unsigned char *nonce32,
                          uint64_t *min_value, uint64_t *max_value, const unsigned char *commit33, const unsigned char *proof, size_t plen,
                          const unsigned char *extra, size_t extra_len, const unsigned char *gen64) {
    return rangeproof_verify_impl(blind_out32, value_out, message_out, outlen, nonce32, min_value, max_value, commit33, proof, plen, extra, extra_len, gen64);
}

/* ---- surjection proof verify: modules/surjection/main_impl.h:46-83 (parse), :360-402 (verify), surjection_impl.h:18-36, :66-92 */
int orc_surjectionproof_verify(const unsigned char *proof, size_t plen, const unsigned char *in_tags64, size_t n_tags, const unsigned char *out_tag64) {
    size_t n_inputs, bm, n_used = 0, i, j, k; const unsigned char *e0, *sd; unsigned char m[32], tmp[33], ser[33]; sha256_t h; ge outp; sc ens; int overflow;
    if (plen < 2) return 0;
    n_inputs = (size_t)proof[0] | ((size_t)proof[1] << 8);
    if (n_inputs > 256) return 0;
    bm = (n_inputs + 7) / 8;
    if (plen < 2 + bm) return 0;
    if (n_inputs % 8 != 0 && (proof[2 + bm - 1] & (unsigned char)(0xFFu << (n_inputs % 8)))) return 0;
    for (i = 0; i < bm; i++) for (k = 0; k < 8; k++) n_used += (proof[2 + i] >> k) & 1;
    if (plen != 2 + bm + 32 * (1 + n_used)) return 0;
    if (n_used == 0 || n_used > n_inputs || n_inputs != n_tags) return 0;
    e0 = proof + 2 + bm; sd = e0 + 32;
    for (j = 0; j < n_used; j++) { sc t; if (sc_set_b32(&t, sd + 32 * j)) return 0; }
    sha_init(&h);
    for (i = 0; i <= n_inputs; i++) {
        const unsigned char *t = i < n_inputs ? in_tags64 + 64 * i : out_tag64;
        ser[0] = 2 + (t[63] & 1); memcpy(ser + 1, t, 32); sha_write(&h, ser, 33);
    }
    sha_final(&h, m);
    generator_load(&outp, out_tag64);
    borromean_hash(tmp, m, 32, e0, 32, 0, 0);
    overflow = sc_set_b32(&ens, tmp);
    j = 0;
    for (i = 0; i < n_inputs; i++) {
        ge inp, rge; gej pub, rgej; sc s;
        if (!((proof[2 + (i >> 3)] >> (i & 7)) & 1)) continue;
        generator_load(&inp, in_tags64 + 64 * i); ge_neg(&inp, &inp);
        gej_set_ge(&pub, &inp); gej_add_ge(&pub, &pub, &outp);
        sc_set_b32(&s, sd + 32 * j);
        if (overflow || u256_is_zero(&s) || u256_is_zero(&ens) || pub.inf) return 0;
        ecmult(&rgej, &pub, &ens, &s);
        if (rgej.inf) return 0;
        ge_set_gej(&rge, &rgej); ge_serialize33(tmp, &rge);
        j++;
        if (j != n_used) { borromean_hash(tmp, m, 32, tmp, 33, 0, j); overflow = sc_set_b32(&ens, tmp); }
    }
    sha_init(&h); sha_write(&h, tmp, 33); sha_write(&h, m, 32); sha_final(&h, ser);
    return memcmp(ser, e0, 32) == 0;
}
/* secp256k1_pedersen_verify_tally (modules/generator/main_impl.h:371-396) over serialized commitments */
int orc_pedersen_verify_tally(const unsigned char *pos33, size_t pcnt, const unsigned char *neg33, size_t ncnt) {
    gej acc; size_t i; ge c;
    gej_set_inf(&acc);
    for (i = 0; i < ncnt; i++) { if (!commitment_parse(&c, neg33 + 33 * i)) return 0; gej_add_ge(&acc, &acc, &c); }
    gej_neg(&acc, &acc);
    for (i = 0; i < pcnt; i++) { if (!commitment_parse(&c, pos33 + 33 * i)) return 0; gej_add_ge(&acc, &acc, &c); }
    return acc.inf;
}

/* ---- BP++ norm-product argument verify: modules/bppp/bppp_norm_product_impl.h:425-552; transcript set-up as
 *      modules/bppp/tests_impl.h:273-311; point pair codec bppp_util.h:28-45; challenge bppp_transcript_impl.h:25-33 */
static size_t ilog2(size_t n) { size_t k = 0; while (n >>= 1) k++; return k; }
static void le64(unsigned char *o, uint64_t n) { int i; for (i = 0; i < 8; i++) o[i] = (unsigned char)(n >> (8 * i)); }
static int bppp_parse_one_of_points(ge *pt, const unsigned char *in65, int idx) {
    unsigned char tmp[33] = {0}; static const unsigned char zeros[33] = {0};
    if (in65[0] > 3) return 0;
    if (memcmp(zeros, in65 + 1 + 32 * idx, 32) != 0) {
        tmp[0] = 2 | ((in65[0] & (2 - idx)) >> (1 - idx));
        memcpy(tmp + 1, in65 + 1 + 32 * idx, 32);
    } else if ((in65[0] & (2 - idx)) != 0) return 0;
    if (memcmp(tmp, zeros, 33) == 0) { memset(pt, 0, sizeof(*pt)); pt->inf = 1; return 1; }
    return pubkey_parse(pt, tmp, 33);
}
int orc_bppp_norm_verify(const unsigned char *proof, size_t plen, const unsigned char *commit_xy64, const unsigned char *rho32,
                         const unsigned char *gens_xy64, size_t g_len, const unsigned char *c_vec32, size_t h_len) {
    sha256_t tr; unsigned char buf[33], b8[8]; ge commit; size_t i, log_g, log_h, n_rounds, ngens = g_len + h_len;
    sc rho, rho_inv, rho_f, mu_f, n, l, v, h_c, *gammas, *s_g, *s_h, *rip, *cv; gej res1, res2, t, pj; int ret = 0;
    if (g_len == 0 || h_len == 0) return 0;
    log_g = ilog2(g_len); log_h = ilog2(h_len); n_rounds = log_g > log_h ? log_g : log_h;
    if (plen != 65 * n_rounds + 64) return 0;
    if ((g_len & (g_len - 1)) || (h_len & (h_len - 1))) return 0;
    sc_set_b32(&rho, rho32);
    ge_from_xy64(&commit, commit_xy64);
    /* transcript */
    sha_init_tagged(&tr, "Bulletproofs_pp/v0/commitment");
    ge_serialize33(buf, &commit); sha_write(&tr, buf, 33);
    u256_to_be(buf, &rho); sha_write(&tr, buf, 32);
    le64(b8, g_len); sha_write(&tr, b8, 8); le64(b8, ngens); sha_write(&tr, b8, 8);
    for (i = 0; i < ngens; i++) { ge g; ge_from_xy64(&g, gens_xy64 + 64 * i); ge_serialize33(buf, &g); sha_write(&tr, buf, 33); }
    le64(b8, h_len); sha_write(&tr, b8, 8);
    cv = malloc(sizeof(sc) * h_len);
    for (i = 0; i < h_len; i++) { sc_set_b32(&cv[i], c_vec32 + 32 * i); u256_to_be(buf, &cv[i]); sha_write(&tr, buf, 32); }
    gammas = malloc(sizeof(sc) * (n_rounds + 1)); s_g = malloc(sizeof(sc) * g_len); s_h = malloc(sizeof(sc) * h_len); rip = malloc(sizeof(sc) * (log_g + 1));
    if (sc_set_b32(&n, proof + 65 * n_rounds)) goto done;
    if (sc_set_b32(&l, proof + 65 * n_rounds + 32)) goto done;
    if (u256_is_zero(&rho)) goto done;
    sc_inv(&rho_inv, &rho);
    if (log_g) { rip[0] = rho_inv; for (i = 1; i < log_g; i++) sc_mul(&rip[i], &rip[i - 1], &rip[i - 1]); }
    rho_f = rho; for (i = 0; i < log_g; i++) sc_mul(&rho_f, &rho_f, &rho_f);
    for (i = 0; i < n_rounds; i++) {
        sha256_t c2; unsigned char hb[32];
        sha_write(&tr, proof + 65 * i, 65);
        c2 = tr; le64(b8, 0); sha_write(&c2, b8, 8); sha_final(&c2, hb);
        sc_set_b32(&gammas[i], hb);
    }
    sc_mul(&s_g[0], &n, &rho_f); sc_mul(&s_g[0], &s_g[0], &rho_inv);
    for (i = 1; i < g_len; i++) { size_t li = ilog2(i), p2 = (size_t)1 << li; sc_mul(&s_g[i], &s_g[i - p2], &gammas[li]); sc_mul(&s_g[i], &s_g[i], &rip[li]); }
    s_h[0] = l;
    for (i = 1; i < h_len; i++) { size_t li = ilog2(i), p2 = (size_t)1 << li; sc_mul(&s_h[i], &s_h[i - p2], &gammas[li]); }
    memset(&h_c, 0, sizeof(h_c));
    for (i = 0; i < h_len; i++) { sc term; sc_mul(&term, &cv[i], &s_h[i]); sc_add(&h_c, &h_c, &term); }
    sc_mul(&mu_f, &rho_f, &rho_f); sc_mul(&v, &n, &n); sc_mul(&v, &v, &mu_f); sc_add(&v, &v, &h_c);
    /* res1 = C + sum gamma_i X_i + (gamma_i^2 - 1) R_i */
    gej_set_ge(&res1, &commit);
    for (i = 0; i < n_rounds; i++) {
        ge X, R; sc g2, one;
        if (!bppp_parse_one_of_points(&X, proof + 65 * i, 0)) goto done;
        if (!bppp_parse_one_of_points(&R, proof + 65 * i, 1)) goto done;
        gej_set_ge(&pj, &X); gej_mul(&t, &pj, &gammas[i]); gej_add(&res1, &res1, &t);
        sc_mul(&g2, &gammas[i], &gammas[i]); sc_set_u64(&one, 1); sc_neg(&one, &one); sc_add(&g2, &g2, &one);
        gej_set_ge(&pj, &R); gej_mul(&t, &pj, &g2); gej_add(&res1, &res1, &t);
    }
    /* res2 = v*G + sum s_g G_i + sum s_h H_i */
    gej_set_ge(&pj, &GEN_G); gej_mul(&res2, &pj, &v);
    for (i = 0; i < ngens; i++) {
        ge g; ge_from_xy64(&g, gens_xy64 + 64 * i); gej_set_ge(&pj, &g);
        gej_mul(&t, &pj, i < g_len ? &s_g[i] : &s_h[i - g_len]); gej_add(&res2, &res2, &t);
    }
    /* gej_eq_var */
    { gej nr; gej_neg(&nr, &res2); gej_add(&t, &res1, &nr); ret = t.inf; }
done:
    free(gammas); free(s_g); free(s_h); free(rip); free(cv);
    return ret;
}
