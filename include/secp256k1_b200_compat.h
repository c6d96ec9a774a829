/* secp256k1_b200_compat.h -- the reference-named C API for the accelerated path, exported by
 * libsecp256k1_b200.so with the reference's own names, argument order and return conventions, so that code
 * compiled against the reference's headers (include/secp256k1.h, secp256k1_extrakeys.h, secp256k1_schnorrsig.h,
 * secp256k1_generator.h, secp256k1_rangeproof.h) links against this library unchanged for these symbols.
 * Each declaration cites the reference declaration it mirrors (paths relative to the reference tree).
 *
 * Semantics: 1 = success/valid, 0 = failure/invalid.  API misuse (NULL where a pointer is required, a static context
 * where a created one is required) calls the context's illegal callback and returns 0 (src/secp256k1.c:66-71); the
 * default callbacks print and abort() (src/util.h:97-105).  CUDA failures go to the error callback.
 * Every call that needs field or group arithmetic runs on the GPU as a batch of one through a process-wide engine
 * (created lazily on the current CUDA device); device state never lives inside secp256k1_context, which stays a small
 * clonable host object (src/secp256k1.c:82-88, :176-185).  For throughput use the *_batch entry points of
 * secp256k1_b200.h -- a single call costs a kernel launch.
 * Deviation: the SHA-256 compression override hook (include/secp256k1.h:443-446) is not honoured on the device.
 */
#ifndef SECP256K1_B200_COMPAT_H
#define SECP256K1_B200_COMPAT_H

#include <stddef.h>
#include <stdint.h>
#include "secp256k1_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* opaque types, same sizes as the reference's (include/secp256k1.h:62-64, :75-77; secp256k1_extrakeys.h:22-24;
 * secp256k1_generator.h:24-26, :60-62) */
typedef struct secp256k1_context_struct secp256k1_context;
typedef struct { unsigned char data[64]; } secp256k1_pubkey;
typedef struct { unsigned char data[64]; } secp256k1_xonly_pubkey;
typedef struct { unsigned char data[64]; } secp256k1_ecdsa_signature;
typedef struct { unsigned char data[64]; } secp256k1_generator;
typedef struct { unsigned char data[64]; } secp256k1_pedersen_commitment;

#ifndef SECP256K1_CONTEXT_NONE
#define SECP256K1_CONTEXT_NONE 1u               /* include/secp256k1.h:215 */
#define SECP256K1_CONTEXT_VERIFY (1u | (1u << 8))
#define SECP256K1_CONTEXT_SIGN (1u | (1u << 9))
#define SECP256K1_EC_COMPRESSED ((1u << 1) | (1u << 8))   /* :225 */
#define SECP256K1_EC_UNCOMPRESSED (1u << 1)               /* :226 */
#endif

/* include/secp256k1.h:246-362 */
SECP256K1_B200_API extern const secp256k1_context *secp256k1_context_static;
SECP256K1_B200_API void secp256k1_selftest(void);
SECP256K1_B200_API secp256k1_context *secp256k1_context_create(unsigned int flags);
SECP256K1_B200_API secp256k1_context *secp256k1_context_clone(const secp256k1_context *ctx);
SECP256K1_B200_API void secp256k1_context_destroy(secp256k1_context *ctx);
SECP256K1_B200_API void secp256k1_context_set_illegal_callback(secp256k1_context *ctx, void (*fun)(const char *message, void *data), const void *data);
SECP256K1_B200_API void secp256k1_context_set_error_callback(secp256k1_context *ctx, void (*fun)(const char *message, void *data), const void *data);
SECP256K1_B200_API int secp256k1_context_randomize(secp256k1_context *ctx, const unsigned char *seed32);   /* :879 */

/* include/secp256k1.h:462-495, :536, :622-627 */
SECP256K1_B200_API int secp256k1_ec_pubkey_parse(const secp256k1_context *ctx, secp256k1_pubkey *pubkey, const unsigned char *input, size_t inputlen);
SECP256K1_B200_API int secp256k1_ec_pubkey_serialize(const secp256k1_context *ctx, unsigned char *output, size_t *outputlen, const secp256k1_pubkey *pubkey, unsigned int flags);
SECP256K1_B200_API int secp256k1_ecdsa_signature_parse_compact(const secp256k1_context *ctx, secp256k1_ecdsa_signature *sig, const unsigned char *input64);
/* host-only helpers of the same API: include/secp256k1.h:462 (pubkey_cmp), :689 (seckey_verify), :867 (tagged_sha256); include/secp256k1_extrakeys.h:77 (xonly cmp) */
SECP256K1_B200_API int secp256k1_ec_pubkey_cmp(const secp256k1_context *ctx, const secp256k1_pubkey *pubkey1, const secp256k1_pubkey *pubkey2);
SECP256K1_B200_API int secp256k1_xonly_pubkey_cmp(const secp256k1_context *ctx, const secp256k1_xonly_pubkey *pk1, const secp256k1_xonly_pubkey *pk2);
SECP256K1_B200_API int secp256k1_ec_seckey_verify(const secp256k1_context *ctx, const unsigned char *seckey);
SECP256K1_B200_API int secp256k1_tagged_sha256(const secp256k1_context *ctx, unsigned char *hash32, const unsigned char *tag, size_t taglen, const unsigned char *msg, size_t msglen);
SECP256K1_B200_API int secp256k1_ecdsa_signature_parse_der(const secp256k1_context *ctx, secp256k1_ecdsa_signature *sig, const unsigned char *input, size_t inputlen);   /* include/secp256k1.h:544, src/secp256k1.c:353, src/ecdsa_impl.h:36-171 */
SECP256K1_B200_API int secp256k1_ecdsa_signature_serialize_der(const secp256k1_context *ctx, unsigned char *output, size_t *outputlen, const secp256k1_ecdsa_signature *sig);   /* include/secp256k1.h:569, src/ecdsa_impl.h:173-197 */
SECP256K1_B200_API int secp256k1_ecdsa_signature_serialize_compact(const secp256k1_context *ctx, unsigned char *output64, const secp256k1_ecdsa_signature *sig);   /* include/secp256k1.h:581 */
SECP256K1_B200_API int secp256k1_ecdsa_signature_normalize(const secp256k1_context *ctx, secp256k1_ecdsa_signature *sigout, const secp256k1_ecdsa_signature *sigin);    /* include/secp256k1.h:640, src/secp256k1.c:479 */
SECP256K1_B200_API int secp256k1_ecdsa_verify(const secp256k1_context *ctx, const secp256k1_ecdsa_signature *sig, const unsigned char *msghash32, const secp256k1_pubkey *pubkey);

/* include/secp256k1_extrakeys.h:47 ; include/secp256k1_schnorrsig.h:178-184 */
SECP256K1_B200_API int secp256k1_xonly_pubkey_parse(const secp256k1_context *ctx, secp256k1_xonly_pubkey *pubkey, const unsigned char *input32);
SECP256K1_B200_API int secp256k1_xonly_pubkey_serialize(const secp256k1_context *ctx, unsigned char *output32, const secp256k1_xonly_pubkey *pubkey);
SECP256K1_B200_API int secp256k1_schnorrsig_verify(const secp256k1_context *ctx, const unsigned char *sig64, const unsigned char *msg, size_t msglen, const secp256k1_xonly_pubkey *pubkey);

/* include/secp256k1_generator.h:27-53, :64-153 */
SECP256K1_B200_API extern const secp256k1_generator *secp256k1_generator_h;
SECP256K1_B200_API int secp256k1_generator_parse(const secp256k1_context *ctx, secp256k1_generator *gen, const unsigned char *input);
SECP256K1_B200_API int secp256k1_generator_serialize(const secp256k1_context *ctx, unsigned char *output, const secp256k1_generator *gen);
SECP256K1_B200_API int secp256k1_pedersen_commitment_parse(const secp256k1_context *ctx, secp256k1_pedersen_commitment *commit, const unsigned char *input);
SECP256K1_B200_API int secp256k1_pedersen_commitment_serialize(const secp256k1_context *ctx, unsigned char *output, const secp256k1_pedersen_commitment *commit);
SECP256K1_B200_API int secp256k1_pedersen_commit(const secp256k1_context *ctx, secp256k1_pedersen_commitment *commit, const unsigned char *blind, uint64_t value, const secp256k1_generator *gen);

/* include/secp256k1_rangeproof.h:70-80, :176-184 */
SECP256K1_B200_API int secp256k1_rangeproof_verify(const secp256k1_context *ctx, uint64_t *min_value, uint64_t *max_value,
        const secp256k1_pedersen_commitment *commit, const unsigned char *proof, size_t plen,
        const unsigned char *extra_commit, size_t extra_commit_len, const secp256k1_generator *gen);
SECP256K1_B200_API int secp256k1_rangeproof_info(const secp256k1_context *ctx, int *exp, int *mantissa, uint64_t *min_value, uint64_t *max_value,
        const unsigned char *proof, size_t plen);

/* include/secp256k1_rangeproof.h:95 */
SECP256K1_B200_API int secp256k1_rangeproof_rewind(const secp256k1_context *ctx, unsigned char *blind_out, uint64_t *value_out, unsigned char *message_out,
        size_t *outlen, const unsigned char *nonce, uint64_t *min_value, uint64_t *max_value, const secp256k1_pedersen_commitment *commit,
        const unsigned char *proof, size_t plen, const unsigned char *extra_commit, size_t extra_commit_len, const secp256k1_generator *gen);

/* include/secp256k1_schnorrsig_halfagg.h:103 */
SECP256K1_B200_API int secp256k1_schnorrsig_aggverify(const secp256k1_context *ctx, const secp256k1_xonly_pubkey *pubkeys, const unsigned char *msgs32,
        size_t n, const unsigned char *aggsig, size_t aggsig_len);

/* include/secp256k1_recovery.h:27-29 (opaque 65 bytes), :40 parse_compact, :52 convert, :65 serialize_compact, :105 recover */
typedef struct secp256k1_ecdsa_recoverable_signature { unsigned char data[65]; } secp256k1_ecdsa_recoverable_signature;
SECP256K1_B200_API int secp256k1_ecdsa_recoverable_signature_parse_compact(const secp256k1_context *ctx, secp256k1_ecdsa_recoverable_signature *sig,
        const unsigned char *input64, int recid);
SECP256K1_B200_API int secp256k1_ecdsa_recoverable_signature_convert(const secp256k1_context *ctx, secp256k1_ecdsa_signature *sig,
        const secp256k1_ecdsa_recoverable_signature *sigin);
SECP256K1_B200_API int secp256k1_ecdsa_recoverable_signature_serialize_compact(const secp256k1_context *ctx, unsigned char *output64, int *recid,
        const secp256k1_ecdsa_recoverable_signature *sig);
SECP256K1_B200_API int secp256k1_ecdsa_recover(const secp256k1_context *ctx, secp256k1_pubkey *pubkey, const secp256k1_ecdsa_recoverable_signature *sig,
        const unsigned char *msghash32);

/* include/secp256k1_surjectionproof.h:42-55, :71-95, :113-138, verify (src/modules/surjection/main_impl.h:360);
 * include/secp256k1_generator.h secp256k1_pedersen_verify_tally (src/modules/generator/main_impl.h:371) */
typedef struct secp256k1_surjectionproof {
    size_t n_inputs;
    unsigned char used_inputs[256 / 8];
    unsigned char data[32 * (1 + 256)];
} secp256k1_surjectionproof;
SECP256K1_B200_API int secp256k1_surjectionproof_parse(const secp256k1_context *ctx, secp256k1_surjectionproof *proof, const unsigned char *input, size_t inputlen);
SECP256K1_B200_API int secp256k1_surjectionproof_serialize(const secp256k1_context *ctx, unsigned char *output, size_t *outputlen, const secp256k1_surjectionproof *proof);
SECP256K1_B200_API size_t secp256k1_surjectionproof_n_total_inputs(const secp256k1_context *ctx, const secp256k1_surjectionproof *proof);
SECP256K1_B200_API size_t secp256k1_surjectionproof_n_used_inputs(const secp256k1_context *ctx, const secp256k1_surjectionproof *proof);
SECP256K1_B200_API size_t secp256k1_surjectionproof_serialized_size(const secp256k1_context *ctx, const secp256k1_surjectionproof *proof);
SECP256K1_B200_API int secp256k1_surjectionproof_verify(const secp256k1_context *ctx, const secp256k1_surjectionproof *proof,
        const secp256k1_generator *ephemeral_input_tags, size_t n_ephemeral_input_tags, const secp256k1_generator *ephemeral_output_tag);
SECP256K1_B200_API int secp256k1_pedersen_verify_tally(const secp256k1_context *ctx, const secp256k1_pedersen_commitment * const *commits, size_t pcnt,
        const secp256k1_pedersen_commitment * const *ncommits, size_t ncnt);

/* the process-wide engine behind the calls above (created on first use); exposed so that a host application can
 * pick the device / share the engine with the batch API */
SECP256K1_B200_API secp256k1_b200_engine *secp256k1_b200_global_engine(void);
SECP256K1_B200_API void secp256k1_b200_global_engine_shutdown(void);

/* Opt-in coalescing of concurrent single calls (no reference counterpart).  A reference-named verification call is a batch of one on the
 * GPU and is latency bound (~1 ms, profiles/r2_single_call_latency.md); a multi-threaded caller gets throughput instead when coalescing is
 * on: concurrent secp256k1_schnorrsig_verify (32-byte messages), secp256k1_ecdsa_verify and secp256k1_rangeproof_verify calls are queued, the
 * first caller that finds no flush in progress waits up to window_us microseconds (or until max_batch calls have gathered) and runs ONE
 * batch launch for all of them; every caller gets exactly the result and callback behaviour of the direct path.  window_us = 0 turns it off
 * (the default); max_batch = 0 keeps the current value (default 4096).  The environment variables SECP256K1_B200_COALESCE_US /
 * SECP256K1_B200_COALESCE_MAX set the same values for programs that cannot be changed. */
SECP256K1_B200_API int secp256k1_b200_coalesce_configure(unsigned window_us, size_t max_batch);
/* test hook: the next `count` GPU calls issued by the reference-named functions report an engine failure (exercises the error callback) */
SECP256K1_B200_API void secp256k1_b200_test_fail_next_gpu_calls(int count);

#ifdef __cplusplus
}
#endif
#endif
