/* oracle/ref_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Builds the UNMODIFIED reference (BlockstreamResearch/secp256k1-zkp, mounted
 * read-only at /root/reference) as one translation unit, the same way the
 * reference's own bench_ecmult.c:9 and tests.c:20 do (`#include "secp256k1.c"`),
 * and exports thin byte-oriented wrappers around the `static` internals that
 * the hot path is made of (ecmult.h:47, ecmult.h:62, field.h, scalar.h,
 * group.h).  The public API (secp256k1_schnorrsig_verify, ...) is exported by
 * the included TU itself.
 *
 * No reference source is copied: this file only #includes it from where it
 * lies.  Output goes to oracle/_ref/libsecp256k1_ref.so (git-ignored).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * load that library.
 */
#include "secp256k1.c"
#include "../include/secp256k1.h"

#include <stdlib.h>
#include <string.h>

#define REF_EXPORT __attribute__((visibility("default")))

/* ---- helpers ---------------------------------------------------------- */
static void ref_fe_from_b32(secp256k1_fe *r, const unsigned char *b) {
    /* mod-p load: callers feed arbitrary 32-byte strings */
    secp256k1_fe_set_b32_mod(r, b);
}
static void ref_fe_to_b32(unsigned char *b, secp256k1_fe *a) {
    secp256k1_fe_normalize_var(a);
    secp256k1_fe_get_b32(b, a);
}
/* point <- 64 bytes: x||y big-endian; all-zero = infinity */
static void ref_ge_from_xy(secp256k1_ge *r, const unsigned char *xy) {
    int i, z = 1;
    for (i = 0; i < 64; i++) z &= (xy[i] == 0);
    if (z) { secp256k1_ge_set_infinity(r); return; }
    {
        secp256k1_fe x, y;
        secp256k1_fe_set_b32_mod(&x, xy);
        secp256k1_fe_set_b32_mod(&y, xy + 32);
        secp256k1_ge_set_xy(r, &x, &y);
    }
}
/* gej -> 64 bytes x||y big-endian affine, all-zero for infinity. returns !infinity */
static int ref_gej_to_xy(unsigned char *xy, const secp256k1_gej *pj) {
    secp256k1_ge p;
    if (secp256k1_gej_is_infinity(pj)) { memset(xy, 0, 64); return 0; }
    secp256k1_ge_set_gej_var(&p, (secp256k1_gej *)pj);
    secp256k1_fe_normalize_var(&p.x);
    secp256k1_fe_normalize_var(&p.y);
    secp256k1_fe_get_b32(xy, &p.x);
    secp256k1_fe_get_b32(xy + 32, &p.y);
    return 1;
}

/* ---- field (field_5x52_int128_impl.h:18,:154; field_impl.h:37; field_5x52_impl.h:491) */
REF_EXPORT void ref_fe_mul(unsigned char *r32, const unsigned char *a32, const unsigned char *b32) {
    secp256k1_fe a, b, r;
    ref_fe_from_b32(&a, a32); ref_fe_from_b32(&b, b32);
    secp256k1_fe_mul(&r, &a, &b);
    ref_fe_to_b32(r32, &r);
}
REF_EXPORT void ref_fe_sqr(unsigned char *r32, const unsigned char *a32) {
    secp256k1_fe a, r;
    ref_fe_from_b32(&a, a32);
    secp256k1_fe_sqr(&r, &a);
    ref_fe_to_b32(r32, &r);
}
REF_EXPORT void ref_fe_add(unsigned char *r32, const unsigned char *a32, const unsigned char *b32) {
    secp256k1_fe a, b;
    ref_fe_from_b32(&a, a32); ref_fe_from_b32(&b, b32);
    secp256k1_fe_add(&a, &b);
    ref_fe_to_b32(r32, &a);
}
REF_EXPORT void ref_fe_negate(unsigned char *r32, const unsigned char *a32) {
    secp256k1_fe a, r;
    ref_fe_from_b32(&a, a32);
    secp256k1_fe_normalize_var(&a);
    secp256k1_fe_negate(&r, &a, 1);
    ref_fe_to_b32(r32, &r);
}
REF_EXPORT void ref_fe_half(unsigned char *r32, const unsigned char *a32) {
    secp256k1_fe a;
    ref_fe_from_b32(&a, a32);
    secp256k1_fe_half(&a);
    ref_fe_to_b32(r32, &a);
}
REF_EXPORT void ref_fe_inv(unsigned char *r32, const unsigned char *a32) {
    secp256k1_fe a, r;
    ref_fe_from_b32(&a, a32);
    secp256k1_fe_inv_var(&r, &a);
    ref_fe_to_b32(r32, &r);
}
/* returns 1 iff a is a square; r = sqrt candidate either way (field_impl.h:37) */
REF_EXPORT int ref_fe_sqrt(unsigned char *r32, const unsigned char *a32) {
    secp256k1_fe a, r; int ok;
    ref_fe_from_b32(&a, a32);
    ok = secp256k1_fe_sqrt(&r, &a);
    ref_fe_to_b32(r32, &r);
    return ok;
}
REF_EXPORT int ref_fe_is_square(const unsigned char *a32) {
    secp256k1_fe a;
    ref_fe_from_b32(&a, a32);
    secp256k1_fe_normalize_var(&a);
    return secp256k1_fe_is_square_var(&a);
}

/* ---- scalar (scalar_4x64_impl.h, scalar_impl.h:142) --------------------- */
REF_EXPORT int ref_scalar_set_b32(unsigned char *r32, const unsigned char *a32) {
    secp256k1_scalar a; int ov;
    secp256k1_scalar_set_b32(&a, a32, &ov);
    secp256k1_scalar_get_b32(r32, &a);
    return ov;
}
REF_EXPORT void ref_scalar_mul(unsigned char *r32, const unsigned char *a32, const unsigned char *b32) {
    secp256k1_scalar a, b, r;
    secp256k1_scalar_set_b32(&a, a32, NULL); secp256k1_scalar_set_b32(&b, b32, NULL);
    secp256k1_scalar_mul(&r, &a, &b);
    secp256k1_scalar_get_b32(r32, &r);
}
REF_EXPORT void ref_scalar_add(unsigned char *r32, const unsigned char *a32, const unsigned char *b32) {
    secp256k1_scalar a, b, r;
    secp256k1_scalar_set_b32(&a, a32, NULL); secp256k1_scalar_set_b32(&b, b32, NULL);
    secp256k1_scalar_add(&r, &a, &b);
    secp256k1_scalar_get_b32(r32, &r);
}
REF_EXPORT void ref_scalar_negate(unsigned char *r32, const unsigned char *a32) {
    secp256k1_scalar a, r;
    secp256k1_scalar_set_b32(&a, a32, NULL);
    secp256k1_scalar_negate(&r, &a);
    secp256k1_scalar_get_b32(r32, &r);
}
REF_EXPORT void ref_scalar_inverse(unsigned char *r32, const unsigned char *a32) {
    secp256k1_scalar a, r;
    secp256k1_scalar_set_b32(&a, a32, NULL);
    secp256k1_scalar_inverse_var(&r, &a);
    secp256k1_scalar_get_b32(r32, &r);
}
/* k = r1 + r2*lambda (mod n), both "small" in absolute value mod n (scalar_impl.h:142) */
REF_EXPORT void ref_scalar_split_lambda(unsigned char *r1_32, unsigned char *r2_32, const unsigned char *k32) {
    secp256k1_scalar k, r1, r2;
    secp256k1_scalar_set_b32(&k, k32, NULL);
    secp256k1_scalar_split_lambda(&r1, &r2, &k);
    secp256k1_scalar_get_b32(r1_32, &r1);
    secp256k1_scalar_get_b32(r2_32, &r2);
}

/* ---- group (group_impl.h:503,:534,:598) --------------------------------- */
/* all three take/return affine x||y (64 bytes, zero = infinity) */
REF_EXPORT int ref_ge_double(unsigned char *r64, const unsigned char *a64) {
    secp256k1_ge a; secp256k1_gej aj, rj;
    ref_ge_from_xy(&a, a64);
    secp256k1_gej_set_ge(&aj, &a);
    secp256k1_gej_double_var(&rj, &aj, NULL);
    return ref_gej_to_xy(r64, &rj);
}
REF_EXPORT int ref_ge_add(unsigned char *r64, const unsigned char *a64, const unsigned char *b64) {
    secp256k1_ge a, b; secp256k1_gej aj, bj, rj;
    ref_ge_from_xy(&a, a64); ref_ge_from_xy(&b, b64);
    secp256k1_gej_set_ge(&aj, &a); secp256k1_gej_set_ge(&bj, &b);
    secp256k1_gej_add_var(&rj, &aj, &bj, NULL);
    return ref_gej_to_xy(r64, &rj);
}
REF_EXPORT int ref_ge_is_valid(const unsigned char *a64) {
    secp256k1_ge a;
    ref_ge_from_xy(&a, a64);
    if (secp256k1_ge_is_infinity(&a)) return 0;
    return secp256k1_ge_is_valid_var(&a);
}

/* batched unit operations with the operand layout of secp256k1_b200_unit_op_batch (same op codes): scalar ops on n x 32 bytes, group
 * ops on n x 96 bytes x||y||s lifted to the Jacobian point (x s^2, y s^3, s) with secp256k1_gej_rescale (group_impl.h:869) */
REF_EXPORT void ref_unit_op_batch(int op, unsigned char *out, const unsigned char *a, const unsigned char *b, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) {
        if (op < 8) {
            secp256k1_scalar x, y, r;
            secp256k1_scalar_set_b32(&x, a + 32 * i, NULL);
            if (b) secp256k1_scalar_set_b32(&y, b + 32 * i, NULL); else secp256k1_scalar_set_int(&y, 0);
            if (op == 0) secp256k1_scalar_mul(&r, &x, &y);
            else if (op == 1) secp256k1_scalar_add(&r, &x, &y);
            else if (op == 2) secp256k1_scalar_negate(&r, &x);
            else if (op == 3) secp256k1_scalar_inverse_var(&r, &x);
            else if (op == 4) {
                secp256k1_scalar r1, r2;
                secp256k1_scalar_split_lambda(&r1, &r2, &x);
                secp256k1_scalar_get_b32(out + 64 * i, &r1); secp256k1_scalar_get_b32(out + 64 * i + 32, &r2);
                continue;
            } else r = x;
            secp256k1_scalar_get_b32(out + 32 * i, &r);
        } else {
            secp256k1_ge ga, gb; secp256k1_gej aj, bj, rj; secp256k1_fe s;
            ref_ge_from_xy(&ga, a + 96 * i);
            secp256k1_gej_set_ge(&aj, &ga);
            if (!ga.infinity) {
                secp256k1_fe_set_b32_mod(&s, a + 96 * i + 64);
                if (secp256k1_fe_normalizes_to_zero_var(&s)) secp256k1_fe_set_int(&s, 1);
                secp256k1_gej_rescale(&aj, &s);
            }
            if (b) ref_ge_from_xy(&gb, b + 96 * i); else secp256k1_ge_set_infinity(&gb);
            secp256k1_gej_set_ge(&bj, &gb);
            if (b && !gb.infinity) {
                secp256k1_fe_set_b32_mod(&s, b + 96 * i + 64);
                if (secp256k1_fe_normalizes_to_zero_var(&s)) secp256k1_fe_set_int(&s, 1);
                secp256k1_gej_rescale(&bj, &s);
            }
            if ((op & 15) == 8) secp256k1_gej_double_var(&rj, &aj, NULL);
            else if ((op & 15) == 9) secp256k1_gej_add_var(&rj, &aj, &bj, NULL);
            else secp256k1_gej_add_ge_var(&rj, &aj, &gb, NULL);
            ref_gej_to_xy(out + 64 * i, &rj);
        }
    }
}

/* ---- ecmult (ecmult_impl.h:365): R = na*A + ng*G ------------------------ */
/* a64: affine x||y BE (zero = infinity); na32/ng32 BE scalars reduced mod n; ng32 may be NULL.
 * out: affine x||y or zero.  returns !infinity */
REF_EXPORT int ref_ecmult(unsigned char *r64, const unsigned char *a64, const unsigned char *na32, const unsigned char *ng32) {
    secp256k1_ge a; secp256k1_gej aj, rj; secp256k1_scalar na, ng;
    ref_ge_from_xy(&a, a64);
    secp256k1_gej_set_ge(&aj, &a);
    secp256k1_scalar_set_b32(&na, na32, NULL);
    if (ng32) secp256k1_scalar_set_b32(&ng, ng32, NULL);
    secp256k1_ecmult(&rj, &aj, &na, ng32 ? &ng : NULL);
    return ref_gej_to_xy(r64, &rj);
}
REF_EXPORT void ref_ecmult_batch(unsigned char *r64, const unsigned char *a64, const unsigned char *na32, const unsigned char *ng32, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) {
        ref_ecmult(r64 + 64 * i, a64 + 64 * i, na32 + 32 * i, ng32 ? ng32 + 32 * i : NULL);
    }
}

/* ---- ecmult_multi_var (ecmult_impl.h:823) ------------------------------- */
typedef struct { const unsigned char *sc; const unsigned char *pt; } ref_msm_cbdata;
static int ref_msm_cb(secp256k1_scalar *sc, secp256k1_ge *pt, size_t idx, void *data) {
    ref_msm_cbdata *d = (ref_msm_cbdata *)data;
    secp256k1_scalar_set_b32(sc, d->sc + 32 * idx, NULL);
    ref_ge_from_xy(pt, d->pt + 64 * idx);
    return 1;
}
/* r = g_sc*G + sum sc[i]*pt[i]; scratch_bytes = 0 selects the NULL-scratch simple path.
 * returns 1 on success (result may be infinity => r64 zero), 0 on scratch failure */
REF_EXPORT int ref_ecmult_multi(unsigned char *r64, const unsigned char *g_sc32, const unsigned char *sc32, const unsigned char *pt64, size_t n, size_t scratch_bytes) {
    secp256k1_callback cb; secp256k1_scratch *scratch = NULL; secp256k1_gej rj; secp256k1_scalar g; int ok;
    ref_msm_cbdata d;
    cb.fn = secp256k1_default_error_callback_fn; cb.data = NULL;
    d.sc = sc32; d.pt = pt64;
    if (scratch_bytes) scratch = secp256k1_scratch_create(&cb, scratch_bytes);
    if (g_sc32) secp256k1_scalar_set_b32(&g, g_sc32, NULL);
    ok = secp256k1_ecmult_multi_var(&cb, scratch, &rj, g_sc32 ? &g : NULL, ref_msm_cb, &d, n);
    if (scratch) secp256k1_scratch_destroy(&cb, scratch);
    if (!ok) { memset(r64, 0, 64); return 0; }
    ref_gej_to_xy(r64, &rj);
    return 1;
}

/* ---- opaque-type bridges ------------------------------------------------ */
/* secp256k1_pubkey (64-byte opaque) -> affine x||y BE (secp256k1.c:280) */
REF_EXPORT int ref_pubkey_to_xy(unsigned char *xy64, const secp256k1_pubkey *pk) {
    secp256k1_ge p;
    secp256k1_ge_from_bytes(&p, pk->data);
    secp256k1_fe_normalize_var(&p.x); secp256k1_fe_normalize_var(&p.y);
    secp256k1_fe_get_b32(xy64, &p.x); secp256k1_fe_get_b32(xy64 + 32, &p.y);
    return 1;
}
REF_EXPORT void ref_xy_to_pubkey(secp256k1_pubkey *pk, const unsigned char *xy64) {
    secp256k1_ge p;
    ref_ge_from_xy(&p, xy64);
    secp256k1_ge_to_bytes(pk->data, &p);
}

/* ---- batch loops for the CPU baseline (one thread each; the caller fans
 *      out over threads).  Each mirrors the reference's own bench loop. ---- */
/* bench schnorrsig_verify (modules/schnorrsig/bench_impl.h:39-48): xonly parse + verify */
REF_EXPORT void ref_schnorrsig_verify_batch(const secp256k1_context *ctx, unsigned char *ok,
        const unsigned char *sig64, const unsigned char *msg32, const unsigned char *pk32, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) {
        secp256k1_xonly_pubkey pk;
        ok[i] = 0;
        if (!secp256k1_xonly_pubkey_parse(ctx, &pk, pk32 + 32 * i)) continue;
        ok[i] = (unsigned char)secp256k1_schnorrsig_verify(ctx, sig64 + 64 * i, msg32 + 32 * i, 32, &pk);
    }
}
/* sig64 = compact r||s; pk33 = compressed SEC1 */
REF_EXPORT void ref_ecdsa_verify_batch(const secp256k1_context *ctx, unsigned char *ok,
        const unsigned char *sig64, const unsigned char *msg32, const unsigned char *pk33, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) {
        secp256k1_pubkey pk; secp256k1_ecdsa_signature sig;
        ok[i] = 0;
        if (!secp256k1_ec_pubkey_parse(ctx, &pk, pk33 + 33 * i, 33)) continue;
        if (!secp256k1_ecdsa_signature_parse_compact(ctx, &sig, sig64 + 64 * i)) continue;
        ok[i] = (unsigned char)secp256k1_ecdsa_verify(ctx, &sig, msg32 + 32 * i, &pk);
    }
}
/* synthetic-input generator: BIP-340 signatures like modules/schnorrsig/bench_impl.h:60-76 */
REF_EXPORT int ref_schnorrsig_make_batch(const secp256k1_context *ctx, unsigned char *sig64, unsigned char *pk32,
        const unsigned char *sk32, const unsigned char *msg32, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) {
        secp256k1_keypair kp; secp256k1_xonly_pubkey pk;
        if (!secp256k1_keypair_create(ctx, &kp, sk32 + 32 * i)) return 0;
        if (!secp256k1_schnorrsig_sign_custom(ctx, sig64 + 64 * i, msg32 + 32 * i, 32, &kp, NULL)) return 0;
        if (!secp256k1_keypair_xonly_pub(ctx, &pk, NULL, &kp)) return 0;
        if (!secp256k1_xonly_pubkey_serialize(ctx, pk32 + 32 * i, &pk)) return 0;
    }
    return 1;
}
REF_EXPORT int ref_ecdsa_make_batch(const secp256k1_context *ctx, unsigned char *sig64, unsigned char *pk33,
        const unsigned char *sk32, const unsigned char *msg32, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) {
        secp256k1_pubkey pk; secp256k1_ecdsa_signature sig; size_t l = 33;
        if (!secp256k1_ec_pubkey_create(ctx, &pk, sk32 + 32 * i)) return 0;
        if (!secp256k1_ecdsa_sign(ctx, &sig, msg32 + 32 * i, sk32 + 32 * i, NULL, NULL)) return 0;
        secp256k1_ecdsa_signature_serialize_compact(ctx, sig64 + 64 * i, &sig);
        secp256k1_ec_pubkey_serialize(ctx, pk33 + 33 * i, &l, &pk, SECP256K1_EC_COMPRESSED);
    }
    return 1;
}
/* recoverable signatures (modules/recovery/main_impl.h:123-159): sign -> compact sig + recid; recover -> uncompressed key x||y */
REF_EXPORT int ref_ecdsa_sign_recoverable_batch(const secp256k1_context *ctx, unsigned char *sig64, unsigned char *recid,
        const unsigned char *sk32, const unsigned char *msg32, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) {
        secp256k1_ecdsa_recoverable_signature sig; int rid;
        if (!secp256k1_ecdsa_sign_recoverable(ctx, &sig, msg32 + 32 * i, sk32 + 32 * i, NULL, NULL)) return 0;
        secp256k1_ecdsa_recoverable_signature_serialize_compact(ctx, sig64 + 64 * i, &rid, &sig);
        recid[i] = (unsigned char)rid;
    }
    return 1;
}
REF_EXPORT void ref_ecdsa_recover_batch(const secp256k1_context *ctx, unsigned char *ok, unsigned char *pubkey64,
        const unsigned char *sig64, const unsigned char *recid, const unsigned char *msg32, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) {
        secp256k1_ecdsa_recoverable_signature sig; secp256k1_pubkey pk; unsigned char ser[65]; size_t l = 65;
        ok[i] = 0; memset(pubkey64 + 64 * i, 0, 64);
        if (recid[i] > 3) continue;   /* the API treats this as an illegal argument (ARG_CHECK) */
        if (!secp256k1_ecdsa_recoverable_signature_parse_compact(ctx, &sig, sig64 + 64 * i, recid[i])) continue;
        if (!secp256k1_ecdsa_recover(ctx, &pk, &sig, msg32 + 32 * i)) continue;
        secp256k1_ec_pubkey_serialize(ctx, ser, &l, &pk, SECP256K1_EC_UNCOMPRESSED);
        memcpy(pubkey64 + 64 * i, ser + 1, 64);
        ok[i] = 1;
    }
}
/* half-aggregation (modules/schnorrsig_halfagg/main_impl.h:99,:103): pubkeys as 32-byte x-only serializations */
REF_EXPORT int ref_schnorrsig_aggregate(const secp256k1_context *ctx, unsigned char *aggsig, size_t *aggsig_len,
        const unsigned char *pk32, const unsigned char *msgs32, const unsigned char *sigs64, size_t n) {
    secp256k1_xonly_pubkey *pks = (secp256k1_xonly_pubkey*)malloc((n ? n : 1) * sizeof(*pks));
    size_t i; int r = 1;
    for (i = 0; i < n && r; i++) r = secp256k1_xonly_pubkey_parse(ctx, &pks[i], pk32 + 32 * i);
    if (r) r = secp256k1_schnorrsig_aggregate(ctx, aggsig, aggsig_len, pks, msgs32, sigs64, n);
    free(pks);
    return r;
}
REF_EXPORT int ref_schnorrsig_aggverify(const secp256k1_context *ctx, const unsigned char *pk32, const unsigned char *msgs32, size_t n,
        const unsigned char *aggsig, size_t aggsig_len) {
    secp256k1_xonly_pubkey *pks = (secp256k1_xonly_pubkey*)malloc((n ? n : 1) * sizeof(*pks));
    size_t i; int r = 1;
    for (i = 0; i < n && r; i++) r = secp256k1_xonly_pubkey_parse(ctx, &pks[i], pk32 + 32 * i);
    if (r) r = secp256k1_schnorrsig_aggverify(ctx, pks, msgs32, n, aggsig, aggsig_len);
    else r = -1;   /* a key that does not parse never reaches aggverify in the reference API */
    free(pks);
    return r;
}
/* pedersen commit loop: out33 = serialized commitment (generator/main_impl.h:309,:281) */
REF_EXPORT void ref_pedersen_commit_batch(const secp256k1_context *ctx, unsigned char *ok, unsigned char *out33,
        const unsigned char *blind32, const uint64_t *value, const secp256k1_generator *gen, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) {
        secp256k1_pedersen_commitment c;
        ok[i] = (unsigned char)secp256k1_pedersen_commit(ctx, &c, blind32 + 32 * i, value[i], gen);
        if (ok[i]) secp256k1_pedersen_commitment_serialize(ctx, out33 + 33 * i, &c);
        else memset(out33 + 33 * i, 0, 33);
    }
}
/* rangeproof: make (sign) and verify loops over fixed-stride proof slots */
REF_EXPORT int ref_rangeproof_make_batch(const secp256k1_context *ctx, unsigned char *proofs, size_t stride, size_t *plen,
        unsigned char *commit33, const unsigned char *blind32, const uint64_t *value, uint64_t min_value,
        int exp, int min_bits, const secp256k1_generator *gen, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) {
        secp256k1_pedersen_commitment c;
        size_t l = stride;
        if (!secp256k1_pedersen_commit(ctx, &c, blind32 + 32 * i, value[i], gen)) return 0;
        /* nonce = commitment bytes, as bench_rangeproof.c:32-34 */
        if (!secp256k1_rangeproof_sign(ctx, proofs + stride * i, &l, min_value, &c, blind32 + 32 * i,
                                       (const unsigned char *)&c, exp, min_bits, value[i], NULL, 0, NULL, 0, gen)) return 0;
        plen[i] = l;
        secp256k1_pedersen_commitment_serialize(ctx, commit33 + 33 * i, &c);
    }
    return 1;
}
REF_EXPORT void ref_rangeproof_verify_batch(const secp256k1_context *ctx, unsigned char *ok, uint64_t *min_value, uint64_t *max_value,
        const unsigned char *commit33, const unsigned char *proofs, size_t stride, const size_t *plen,
        const unsigned char *extra, size_t extra_len, const secp256k1_generator *gen, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) {
        secp256k1_pedersen_commitment c;
        ok[i] = 0; min_value[i] = 0; max_value[i] = 0;
        if (!secp256k1_pedersen_commitment_parse(ctx, &c, commit33 + 33 * i)) continue;
        ok[i] = (unsigned char)secp256k1_rangeproof_verify(ctx, &min_value[i], &max_value[i], &c,
                    proofs + stride * i, plen[i], extra_len ? extra : NULL, extra_len, gen);
    }
}

/* rangeproof with an embedded message and an explicit nonce (sign: main_impl.h:72), and the rewinding call (main_impl.h:31) */
REF_EXPORT int ref_rangeproof_make_msg_batch(const secp256k1_context *ctx, unsigned char *proofs, size_t stride, size_t *plen,
        unsigned char *commit33, const unsigned char *blind32, const uint64_t *value, const uint64_t *min_value,
        const int *exp, const int *min_bits, const unsigned char *nonce32, const unsigned char *msg, size_t msg_stride, const size_t *msg_len,
        const unsigned char *extra, size_t extra_len, const secp256k1_generator *gen, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) {
        secp256k1_pedersen_commitment c;
        size_t l = stride;
        if (!secp256k1_pedersen_commit(ctx, &c, blind32 + 32 * i, value[i], gen)) return 0;
        if (!secp256k1_rangeproof_sign(ctx, proofs + stride * i, &l, min_value[i], &c, blind32 + 32 * i, nonce32 + 32 * i, exp[i], min_bits[i],
                                       value[i], msg_len[i] ? msg + msg_stride * i : NULL, msg_len[i], extra_len ? extra : NULL, extra_len, gen)) return 0;
        plen[i] = l;
        secp256k1_pedersen_commitment_serialize(ctx, commit33 + 33 * i, &c);
    }
    return 1;
}
REF_EXPORT void ref_rangeproof_rewind_batch(const secp256k1_context *ctx, unsigned char *ok, unsigned char *blind_out32, uint64_t *value_out,
        unsigned char *msg_out, size_t msg_cap, size_t *msg_len_out, uint64_t *min_value, uint64_t *max_value, const unsigned char *nonce32,
        const unsigned char *commit33, const unsigned char *proofs, size_t stride, const size_t *plen,
        const unsigned char *extra, size_t extra_len, const secp256k1_generator *gen, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) {
        secp256k1_pedersen_commitment c;
        size_t ml = msg_cap;
        ok[i] = 0; min_value[i] = 0; max_value[i] = 0; value_out[i] = 0; msg_len_out[i] = 0;
        memset(blind_out32 + 32 * i, 0, 32);
        if (!secp256k1_pedersen_commitment_parse(ctx, &c, commit33 + 33 * i)) continue;
        ok[i] = (unsigned char)secp256k1_rangeproof_rewind(ctx, blind_out32 + 32 * i, &value_out[i], msg_cap ? msg_out + msg_cap * i : NULL, msg_cap ? &ml : NULL,
                    nonce32 + 32 * i, &min_value[i], &max_value[i], &c, proofs + stride * i, plen[i], extra_len ? extra : NULL, extra_len, gen);
        if (ok[i]) msg_len_out[i] = msg_cap ? ml : 0;
        else { value_out[i] = 0; memset(blind_out32 + 32 * i, 0, 32); }
    }
}

/* ---- BP++ norm-product argument (bppp_norm_product_impl.h:105,:223,:425) -- */
/* The transcript is initialised exactly as the reference tests do
 * (modules/bppp/tests_impl.h:273-311): tagged "Bulletproofs_pp/v0/commitment"
 * then commit||rho||g_len||gens->n||gens||c_len||c_vec.                     */
static void ref_bppp_transcript_init(const secp256k1_context *ctx, secp256k1_sha256 *transcript, const secp256k1_scalar *rho,
        const secp256k1_bppp_generators *gens, size_t g_len, const secp256k1_scalar *c_vec, size_t c_len, const secp256k1_ge *commit) {
    const secp256k1_hash_ctx *hash_ctx = secp256k1_get_hash_context(ctx);
    unsigned char ser_commit[33], ser_scalar[32], ser_le64[8];
    size_t i;
    secp256k1_ge comm = *commit;
    secp256k1_bppp_sha256_tagged_commitment_init(transcript);
    secp256k1_fe_normalize(&comm.x); secp256k1_fe_normalize(&comm.y);
    secp256k1_bppp_serialize_pt(&ser_commit[0], &comm);
    secp256k1_sha256_write(hash_ctx, transcript, ser_commit, 33);
    secp256k1_scalar_get_b32(ser_scalar, rho);
    secp256k1_sha256_write(hash_ctx, transcript, ser_scalar, 32);
    secp256k1_bppp_le64(ser_le64, g_len);
    secp256k1_sha256_write(hash_ctx, transcript, ser_le64, 8);
    secp256k1_bppp_le64(ser_le64, gens->n);
    secp256k1_sha256_write(hash_ctx, transcript, ser_le64, 8);
    for (i = 0; i < gens->n; i++) {
        secp256k1_fe_normalize(&gens->gens[i].x); secp256k1_fe_normalize(&gens->gens[i].y);
        secp256k1_bppp_serialize_pt(&ser_commit[0], &gens->gens[i]);
        secp256k1_sha256_write(hash_ctx, transcript, ser_commit, 33);
    }
    secp256k1_bppp_le64(ser_le64, c_len);
    secp256k1_sha256_write(hash_ctx, transcript, ser_le64, 8);
    for (i = 0; i < c_len; i++) {
        secp256k1_scalar_get_b32(ser_scalar, &c_vec[i]);
        secp256k1_sha256_write(hash_ctx, transcript, ser_scalar, 32);
    }
}
/* export generators as n * 64 bytes affine x||y */
REF_EXPORT void ref_bppp_generators_xy(unsigned char *xy, const secp256k1_bppp_generators *gens) {
    size_t i;
    for (i = 0; i < gens->n; i++) {
        secp256k1_ge g = gens->gens[i];
        secp256k1_fe_normalize_var(&g.x); secp256k1_fe_normalize_var(&g.y);
        secp256k1_fe_get_b32(xy + 64 * i, &g.x); secp256k1_fe_get_b32(xy + 64 * i + 32, &g.y);
    }
}
/* commit + prove: n_vec (g_len scalars), l_vec and c_vec (h_len scalars), rho; all 32-byte BE.
 * writes commit xy64 and the proof. returns 1 on success. */
REF_EXPORT int ref_bppp_norm_prove(const secp256k1_context *ctx, unsigned char *proof, size_t *plen, unsigned char *commit_xy64,
        const unsigned char *rho32, const secp256k1_bppp_generators *gens,
        const unsigned char *n_vec32, size_t g_len, const unsigned char *l_vec32, const unsigned char *c_vec32, size_t h_len) {
    secp256k1_scalar *nv, *lv, *cv, rho, mu; secp256k1_ge *gs; secp256k1_ge commit; secp256k1_sha256 transcript;
    secp256k1_scratch *scratch = secp256k1_scratch_space_create(ctx, 1000 * 1000);
    size_t i; int ok;
    nv = malloc(sizeof(*nv) * g_len); lv = malloc(sizeof(*lv) * h_len); cv = malloc(sizeof(*cv) * h_len);
    gs = malloc(sizeof(*gs) * (g_len + h_len));
    for (i = 0; i < g_len; i++) secp256k1_scalar_set_b32(&nv[i], n_vec32 + 32 * i, NULL);
    for (i = 0; i < h_len; i++) { secp256k1_scalar_set_b32(&lv[i], l_vec32 + 32 * i, NULL); secp256k1_scalar_set_b32(&cv[i], c_vec32 + 32 * i, NULL); }
    secp256k1_scalar_set_b32(&rho, rho32, NULL);
    secp256k1_scalar_sqr(&mu, &rho);
    ok = secp256k1_bppp_commit(ctx, scratch, &commit, gens, nv, g_len, lv, h_len, cv, h_len, &mu);
    if (ok) {
        secp256k1_gej cj; secp256k1_gej_set_ge(&cj, &commit); ref_gej_to_xy(commit_xy64, &cj);
        ref_bppp_transcript_init(ctx, &transcript, &rho, gens, g_len, cv, h_len, &commit);
        memcpy(gs, gens->gens, sizeof(*gs) * (g_len + h_len));
        ok = secp256k1_bppp_rangeproof_norm_product_prove(ctx, scratch, proof, plen, &transcript, &rho, gs, g_len + h_len, nv, g_len, lv, h_len, cv, h_len);
    }
    free(nv); free(lv); free(cv); free(gs);
    secp256k1_scratch_space_destroy(ctx, scratch);
    return ok;
}
REF_EXPORT int ref_bppp_norm_verify(const secp256k1_context *ctx, const unsigned char *proof, size_t plen, const unsigned char *commit_xy64,
        const unsigned char *rho32, const secp256k1_bppp_generators *gens, size_t g_len, const unsigned char *c_vec32, size_t h_len) {
    secp256k1_scalar *cv, rho; secp256k1_ge commit; secp256k1_sha256 transcript;
    secp256k1_scratch *scratch = secp256k1_scratch_space_create(ctx, 1000 * 1000);
    size_t i; int ok;
    cv = malloc(sizeof(*cv) * (h_len ? h_len : 1));
    for (i = 0; i < h_len; i++) secp256k1_scalar_set_b32(&cv[i], c_vec32 + 32 * i, NULL);
    secp256k1_scalar_set_b32(&rho, rho32, NULL);
    ref_ge_from_xy(&commit, commit_xy64);
    ref_bppp_transcript_init(ctx, &transcript, &rho, gens, g_len, cv, h_len, &commit);
    ok = secp256k1_bppp_rangeproof_norm_product_verify(ctx, scratch, proof, plen, &transcript, &rho, gens, g_len, cv, h_len, &commit);
    free(cv);
    secp256k1_scratch_space_destroy(ctx, scratch);
    return ok;
}

REF_EXPORT size_t ref_sizeof(int what) {
    switch (what) {
    case 0: return sizeof(secp256k1_fe);
    case 1: return sizeof(secp256k1_ge);
    case 2: return sizeof(secp256k1_gej);
    case 3: return sizeof(secp256k1_scalar);
    case 4: return sizeof(secp256k1_ge_storage);
    default: return 0;
    }
}

/* ---- surjection proofs (modules/surjection/main_impl.h:46, :238-357 generate, :360-402 verify) --------------------- */
/* Makes one proof: n_inputs fixed asset tags A_i = SHA256-derived from tag_seed32 (tag i = seed with byte 0 += i),
 * output asset = input `real_index`, ephemeral tags = generator_generate_blinded(A_i, blind_i) with blinds derived the
 * same way.  Writes the n ephemeral input tags (64 bytes each), the ephemeral output tag and the SERIALIZED proof. */
REF_EXPORT int ref_surjection_make(const secp256k1_context *ctx, unsigned char *proof_ser, size_t *plen, unsigned char *eph_in64,
        unsigned char *eph_out64, size_t n_inputs, size_t n_to_use, size_t real_index, const unsigned char *seed32) {
    secp256k1_fixed_asset_tag *fixed = malloc(sizeof(*fixed) * n_inputs), fixed_out;
    secp256k1_generator *eph = malloc(sizeof(*eph) * n_inputs), eph_out;
    unsigned char (*blind)[32] = malloc(32 * n_inputs), out_blind[32], tmp[33];
    secp256k1_surjectionproof proof;
    secp256k1_sha256 sha;
    const secp256k1_hash_ctx *hash_ctx = secp256k1_get_hash_context(ctx);
    size_t i, input_index = 0;
    int ok = 1;
    for (i = 0; i < n_inputs; i++) {
        memcpy(tmp, seed32, 32); tmp[32] = (unsigned char)i;
        secp256k1_sha256_initialize(&sha); secp256k1_sha256_write(hash_ctx, &sha, (const unsigned char *)"tag", 3); secp256k1_sha256_write(hash_ctx, &sha, tmp, 33);
        secp256k1_sha256_finalize(hash_ctx, &sha, fixed[i].data);
        secp256k1_sha256_initialize(&sha); secp256k1_sha256_write(hash_ctx, &sha, (const unsigned char *)"bld", 3); secp256k1_sha256_write(hash_ctx, &sha, tmp, 33);
        secp256k1_sha256_finalize(hash_ctx, &sha, blind[i]);
        blind[i][0] &= 0x7F;
        ok &= secp256k1_generator_generate_blinded(ctx, &eph[i], fixed[i].data, blind[i]);
        memcpy(eph_in64 + 64 * i, eph[i].data, 64);
    }
    fixed_out = fixed[real_index];
    secp256k1_sha256_initialize(&sha); secp256k1_sha256_write(hash_ctx, &sha, (const unsigned char *)"out", 3); secp256k1_sha256_write(hash_ctx, &sha, seed32, 32);
    secp256k1_sha256_finalize(hash_ctx, &sha, out_blind);
    out_blind[0] &= 0x7F;
    ok &= secp256k1_generator_generate_blinded(ctx, &eph_out, fixed_out.data, out_blind);
    memcpy(eph_out64, eph_out.data, 64);
    if (ok) ok = secp256k1_surjectionproof_initialize(ctx, &proof, &input_index, fixed, n_inputs, n_to_use, &fixed_out, 100, seed32) > 0;
    if (ok) ok = secp256k1_surjectionproof_generate(ctx, &proof, eph, n_inputs, &eph_out, input_index, blind[input_index], out_blind);
    if (ok) ok = secp256k1_surjectionproof_serialize(ctx, proof_ser, plen, &proof);
    free(fixed); free(eph); free(blind);
    return ok;
}
/* parse + verify, as a caller of the public API would */
REF_EXPORT int ref_surjection_verify(const secp256k1_context *ctx, const unsigned char *proof_ser, size_t plen, const unsigned char *eph_in64,
        size_t n_inputs, const unsigned char *eph_out64) {
    secp256k1_surjectionproof proof;
    secp256k1_generator *eph = malloc(sizeof(*eph) * (n_inputs ? n_inputs : 1)), eph_out;
    size_t i; int ok;
    for (i = 0; i < n_inputs; i++) memcpy(eph[i].data, eph_in64 + 64 * i, 64);
    memcpy(eph_out.data, eph_out64, 64);
    ok = secp256k1_surjectionproof_parse(ctx, &proof, proof_ser, plen);
    if (ok) ok = secp256k1_surjectionproof_verify(ctx, &proof, eph, n_inputs, &eph_out);
    free(eph);
    return ok;
}
/* secp256k1_pedersen_verify_tally (modules/generator/main_impl.h:371-396) over serialized commitments; unparsable -> 0 */
REF_EXPORT int ref_pedersen_verify_tally(const secp256k1_context *ctx, const unsigned char *pos33, size_t pcnt, const unsigned char *neg33, size_t ncnt) {
    secp256k1_pedersen_commitment *c = malloc(sizeof(*c) * (pcnt + ncnt + 1));
    const secp256k1_pedersen_commitment **pp = malloc(sizeof(*pp) * (pcnt + ncnt + 1));
    size_t i; int ok = 1;
    for (i = 0; i < pcnt && ok; i++) { ok = secp256k1_pedersen_commitment_parse(ctx, &c[i], pos33 + 33 * i); pp[i] = &c[i]; }
    for (i = 0; i < ncnt && ok; i++) { ok = secp256k1_pedersen_commitment_parse(ctx, &c[pcnt + i], neg33 + 33 * i); pp[pcnt + i] = &c[pcnt + i]; }
    if (ok) ok = secp256k1_pedersen_verify_tally(ctx, pp, pcnt, pp + pcnt, ncnt);
    free(c); free(pp);
    return ok;
}
