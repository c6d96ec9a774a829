/* tests/c_dropin/dropin_test.c -- a caller written against the REFERENCE's public headers only (no header of this repository),
 * linked against libsecp256k1_b200.so instead of libsecp256k1: the "drop-in for the single-call API" of INTEGRATION.md section 1.
 * Built by the CPU tier where /root/reference/include exists (tests/test_c_dropin.py), run by the GPU tier.  Exit code 0 = every
 * check passed; otherwise the number of the first failing check. */
#include <secp256k1.h>
#include <secp256k1_extrakeys.h>
#include <secp256k1_schnorrsig.h>
#include <secp256k1_generator.h>
#include <secp256k1_rangeproof.h>
#include <secp256k1_recovery.h>
#include <stdio.h>
#include <string.h>

static int hex(unsigned char *out, const char *s) { size_t i, n = strlen(s) / 2; for (i = 0; i < n; i++) { unsigned v; sscanf(s + 2 * i, "%2x", &v); out[i] = (unsigned char)v; } return (int)n; }
#define CHECK(k, cond) do { if (!(cond)) { fprintf(stderr, "check %d failed: %s\n", k, #cond); return k; } } while (0)

int main(void) {
    unsigned char pk32[32], sig64[64], msg[32], ser[65];
    size_t len;
    secp256k1_xonly_pubkey xpk;
    secp256k1_pubkey pub, rec;
    secp256k1_ecdsa_recoverable_signature rsig;
    secp256k1_ecdsa_signature esig;
    secp256k1_pedersen_commitment commit;
    unsigned char blind[32], c33[33], c33b[33];
    secp256k1_context *ctx = secp256k1_context_create(SECP256K1_CONTEXT_NONE);
    CHECK(1, ctx != NULL);
    /* BIP-340 test vector 0 (src/modules/schnorrsig/tests_impl.h) */
    hex(pk32, "f9308a019258c31049344f85f89d5229b531c845836f99b08601f113bce036f9");
    hex(sig64, "e907831f80848d1069a5371b402410364bdf1c5f8307b0084c55f1ce2dca821525f66a4a85ea8b71e482a74f382d2ce5ebeee8fdb2172f477df4900d310536c0");
    memset(msg, 0, 32);
    CHECK(2, secp256k1_xonly_pubkey_parse(ctx, &xpk, pk32) == 1);
    CHECK(3, secp256k1_schnorrsig_verify(ctx, sig64, msg, 32, &xpk) == 1);
    sig64[40] ^= 1;
    CHECK(4, secp256k1_schnorrsig_verify(ctx, sig64, msg, 32, &xpk) == 0);
    /* the generator as a compressed key: parse / serialize round trip */
    hex(ser, "0279be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798");
    CHECK(5, secp256k1_ec_pubkey_parse(ctx, &pub, ser, 33) == 1);
    len = 65;
    CHECK(6, secp256k1_ec_pubkey_serialize(ctx, ser, &len, &pub, SECP256K1_EC_UNCOMPRESSED) == 1 && len == 65 && ser[0] == 4 && ser[64] == 0xb8);
    /* recovery edge case (src/modules/recovery/tests_impl.h:156-199): recid 1 recovers, recid 0 does not; the recovered key verifies */
    memcpy(msg, "This is a very secret message...", 32);
    hex(sig64, "67cb285f9cd194e840d629397af5569662fde44649995963179a7dd17bd235324b1b7df34ce1f68e694ff6f11ac751dd7dd73e387ee4fc866e1be8ecc7dd9557");
    CHECK(7, secp256k1_ecdsa_recoverable_signature_parse_compact(ctx, &rsig, sig64, 0) == 1 && secp256k1_ecdsa_recover(ctx, &rec, &rsig, msg) == 0);
    CHECK(8, secp256k1_ecdsa_recoverable_signature_parse_compact(ctx, &rsig, sig64, 1) == 1 && secp256k1_ecdsa_recover(ctx, &rec, &rsig, msg) == 1);
    CHECK(9, secp256k1_ecdsa_recoverable_signature_convert(ctx, &esig, &rsig) == 1);
    secp256k1_ecdsa_signature_normalize(ctx, &esig, &esig);
    CHECK(10, secp256k1_ecdsa_verify(ctx, &esig, msg, &rec) == 1);
    /* a DER signature (Wycheproof ecdsa_secp256k1_sha256_bitcoin tcId 2): parse, normalize, verify against the uncompressed key */
    {
        unsigned char der[80], pk65[65], h[32];
        int dl = hex(der, "3045022100813ef79ccefa9a56f7ba805f0e478584fe5f0dd5f567bc09b5123ccbc983236502206ff18a52dcc0336f7af62400a6dd9b810732baf1ff758000d6f613a556eb31ba");
        hex(pk65, "04b838ff44e5bc177bf21189d0766082fc9d843226887fc9760371100b7ee20a6ff0c9d75bfba7b31a6bca1974496eeb56de357071955d83c4b1badaa0b21832e9");
        hex(h, "bb5a52f42f9c9261ed4361f59422a1e30036e7c32b270c8807a419feca605023");
        CHECK(14, secp256k1_ecdsa_signature_parse_der(ctx, &esig, der, (size_t)dl) == 1);
        CHECK(15, secp256k1_ec_pubkey_parse(ctx, &pub, pk65, 65) == 1);
        CHECK(16, secp256k1_ecdsa_verify(ctx, &esig, h, &pub) == 1);
        der[10] ^= 1;
        CHECK(17, secp256k1_ecdsa_signature_parse_der(ctx, &esig, der, (size_t)dl) == 1 && secp256k1_ecdsa_verify(ctx, &esig, h, &pub) == 0);
    }
    /* Pedersen commitment with the default value generator: commit / serialize / parse round trip */
    memset(blind, 0x11, 32);
    CHECK(11, secp256k1_pedersen_commit(ctx, &commit, blind, 12345, secp256k1_generator_h) == 1);
    CHECK(12, secp256k1_pedersen_commitment_serialize(ctx, c33, &commit) == 1 && (c33[0] & 0xFE) == 8);
    CHECK(13, secp256k1_pedersen_commitment_parse(ctx, &commit, c33) == 1 && secp256k1_pedersen_commitment_serialize(ctx, c33b, &commit) == 1 && memcmp(c33, c33b, 33) == 0);
    secp256k1_context_destroy(ctx);
    printf("dropin ok\n");
    return 0;
}
