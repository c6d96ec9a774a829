/* secp256k1_b200.h -- C ABI of the B200-native batch engine (libsecp256k1_b200.so).
 *
 * Plain pointers and sizes only.  Two families of entry points:
 *   *_batch      host buffers in / host buffers out (H2D + kernels + D2H inside the call)
 *   *_batch_dev  device pointers, asynchronous on the given CUDA stream (cudaStream_t passed as void*)
 * Every function returns 1 on success and 0 on an engine/CUDA failure (secp256k1_b200_last_error()
 * gives the reason); per-item accept/reject results go to the `ok` byte array (1 = valid).
 * There is NO CPU fallback: without a usable B200 every call fails.
 *
 * What each entry point replaces in the reference (paths relative to the reference tree):
 *   secp256k1_b200_fe_op_batch          batched secp256k1_fe_mul/_sqr/... (src/field_5x52_int128_impl.h:18,:154;
 *                                       src/field_5x52_impl.h; src/field_impl.h:37) -- roofline probe + self test
 *   secp256k1_b200_ecmult_batch         static secp256k1_ecmult (src/ecmult.h:47, src/ecmult_impl.h:365), one call per item
 *   secp256k1_b200_schnorrsig_verify_batch   secp256k1_schnorrsig_verify (include/secp256k1_schnorrsig.h:178,
 *                                       src/modules/schnorrsig/main_impl.h:215) [+ secp256k1_xonly_pubkey_parse,
 *                                       include/secp256k1_extrakeys.h:47 when pkfmt = XONLY32]
 *   secp256k1_b200_ecdsa_verify_batch   secp256k1_ecdsa_verify (include/secp256k1.h:622, src/secp256k1.c:498)
 *                                       [+ secp256k1_ec_pubkey_parse / secp256k1_ecdsa_signature_parse_compact]
 *   secp256k1_b200_ecdsa_recover_batch  secp256k1_ecdsa_recover (include/secp256k1_recovery.h:105, src/modules/recovery/main_impl.h:137)
 *   secp256k1_b200_schnorrsig_aggverify secp256k1_schnorrsig_aggverify (include/secp256k1_schnorrsig_halfagg.h:103)
 *   secp256k1_b200_pedersen_commit_batch     secp256k1_pedersen_commit (include/secp256k1_generator.h:114,
 *                                       src/modules/generator/main_impl.h:309) + _commitment_serialize (:281)
 *   secp256k1_b200_rangeproof_verify_batch   secp256k1_rangeproof_verify (include/secp256k1_rangeproof.h:70,
 *                                       src/modules/rangeproof/main_impl.h:54)
 *   secp256k1_b200_rangeproof_rewind_batch   secp256k1_rangeproof_rewind (include/secp256k1_rangeproof.h:95, main_impl.h:31)
 *   secp256k1_b200_ecmult_multi         static secp256k1_ecmult_multi_var (src/ecmult.h:62, src/ecmult_impl.h:823)
 *   secp256k1_b200_bppp_norm_verify_batch    static secp256k1_bppp_rangeproof_norm_product_verify
 *                                       (src/modules/bppp/bppp_norm_product_impl.h:425)
 */
#ifndef SECP256K1_B200_H
#define SECP256K1_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SECP256K1_B200_API __attribute__((visibility("default")))
#else
#define SECP256K1_B200_API
#endif

typedef struct secp256k1_b200_engine secp256k1_b200_engine;

/* public-key encodings accepted by the batch verifiers */
#define SECP256K1_B200_PK_OPAQUE64     0 /* secp256k1_pubkey / secp256k1_xonly_pubkey 64-byte opaque (include/secp256k1.h:62) */
#define SECP256K1_B200_PK_XONLY32      1 /* BIP-340 x-only 32 bytes; parsing (lift_x) is done on the GPU */
#define SECP256K1_B200_PK_COMPRESSED33 2 /* SEC1 compressed 33 bytes; decompression on the GPU */
/* signature encodings for ECDSA */
#define SECP256K1_B200_SIG_OPAQUE64    0 /* secp256k1_ecdsa_signature 64-byte opaque (src/secp256k1.c:393-415) */
#define SECP256K1_B200_SIG_COMPACT64   1 /* r||s big-endian, as secp256k1_ecdsa_signature_parse_compact takes */

/* field-op selector for secp256k1_b200_fe_op_batch */
#define SECP256K1_B200_FE_MUL  0
#define SECP256K1_B200_FE_SQR  1
#define SECP256K1_B200_FE_ADD  2
#define SECP256K1_B200_FE_SUB  3
#define SECP256K1_B200_FE_HALF 4
#define SECP256K1_B200_FE_INV  5
#define SECP256K1_B200_FE_SQRT 6
#define SECP256K1_B200_FE_NEG  7
#define SECP256K1_B200_FE_INV_FERMAT 8      /* a^(p-2): the round-1 inversion, kept as a cross-check of FE_INV (division steps, csrc/modinv.cuh) */
#define SECP256K1_B200_FE_IS_SQUARE 9       /* out = 1 / 0: Jacobi symbol (secp256k1_fe_is_square_var, src/field_5x52_impl.h:501) */
#define SECP256K1_B200_FE_IS_SQUARE_SQRT 10 /* the same through a full square root (round-1 form, cross-check) */
#define SECP256K1_B200_FE_MUL_PORTABLE 16 /* same op through the portable (non-PTX) limb code: on-device A/B self test */
#define SECP256K1_B200_FE_SQR_PORTABLE 17

/* ---- engine ------------------------------------------------------------------------------ */
/* Creates an engine on CUDA device `device` (-1 = current).  gen_window_bits = 0 picks the default
 * comb window for the generator table (16 -> 64 MiB).  If build_tables is 0 the table memory is
 * allocated but left for the caller to fill (e.g. by an NCCL broadcast from rank 0) -- see
 * secp256k1_b200_engine_table_ptr. */
SECP256K1_B200_API int secp256k1_b200_engine_create(secp256k1_b200_engine **out, int device, int gen_window_bits, int build_tables);
SECP256K1_B200_API void secp256k1_b200_engine_destroy(secp256k1_b200_engine *e);
SECP256K1_B200_API const char *secp256k1_b200_last_error(void);
/* device pointer / size of the generator comb table (for the init-time NCCL broadcast) */
SECP256K1_B200_API void *secp256k1_b200_engine_table_ptr(secp256k1_b200_engine *e);
SECP256K1_B200_API size_t secp256k1_b200_engine_table_bytes(const secp256k1_b200_engine *e);
SECP256K1_B200_API int secp256k1_b200_engine_table_build(secp256k1_b200_engine *e);
/* copy the table out of / into the engine (device pointers, table_bytes long): rank 0 builds + exports, the
 * harness broadcasts the buffer with NCCL, the other ranks import. */
SECP256K1_B200_API int secp256k1_b200_engine_table_export_dev(secp256k1_b200_engine *e, void *d_dst, void *stream);
SECP256K1_B200_API int secp256k1_b200_engine_table_import_dev(secp256k1_b200_engine *e, const void *d_src, void *stream);
/* number of kernel launches issued by this engine so far */
SECP256K1_B200_API uint64_t secp256k1_b200_engine_launch_count(const secp256k1_b200_engine *e);
SECP256K1_B200_API int secp256k1_b200_engine_device(const secp256k1_b200_engine *e);

/* ---- field micro-kernel (roofline probe / self test) ---------------------------------------- */
/* out[i] = op(a[i], b[i]) applied `iters` times (iters >= 1: r = op(r, b) repeated, for timing);
 * elements are 32-byte big-endian, any 256-bit value accepted (reduced mod p), outputs canonical. */
SECP256K1_B200_API int secp256k1_b200_fe_op_batch(secp256k1_b200_engine *e, int op, unsigned char *out32,
        const unsigned char *a32, const unsigned char *b32, size_t n, int iters);
SECP256K1_B200_API int secp256k1_b200_fe_op_batch_dev(secp256k1_b200_engine *e, int op, unsigned char *d_out32,
        const unsigned char *d_a32, const unsigned char *d_b32, size_t n, int iters, void *stream);

/* Unit kernels of the scalar and group primitives, for parity tests against the reference's own functions (scalar_4x64_impl.h,
 * scalar_impl.h:142, group_impl.h:503,:534,:598).  Scalar ops: a, b = n x 32 bytes big-endian, out = n x 32 (split_lambda: n x 64,
 * r1 || r2).  Group ops: a, b = n x 96 bytes  x || y || s  (affine point, all-zero x||y = infinity, lifted to the Jacobian point
 * (x s^2, y s^3, s) so that z != 1; the mixed addition uses b's x||y only), out = n x 64 affine x||y, zeros = infinity.  b may be NULL for
 * one-operand ops.  The _G4 ops run the lane-parallel forms (4 lanes per item) used by the MSM tail. */
#define SECP256K1_B200_UNIT_SC_MUL          0
#define SECP256K1_B200_UNIT_SC_ADD          1
#define SECP256K1_B200_UNIT_SC_NEGATE       2
#define SECP256K1_B200_UNIT_SC_INVERSE      3
#define SECP256K1_B200_UNIT_SC_SPLIT_LAMBDA 4
#define SECP256K1_B200_UNIT_SC_REDUCE       5
#define SECP256K1_B200_UNIT_GEJ_DOUBLE      8
#define SECP256K1_B200_UNIT_GEJ_ADD         9
#define SECP256K1_B200_UNIT_GEJ_ADD_GE      10
#define SECP256K1_B200_UNIT_GEJ_DOUBLE_G4   24
#define SECP256K1_B200_UNIT_GEJ_ADD_G4      25
#define SECP256K1_B200_UNIT_GEJ_ADD_GE_G4   26
#define SECP256K1_B200_UNIT_GEJ_DOUBLE_W10  27  /* the warp-wide doubling chain of the MSM tail (ge_w10.cuh), 1 doubling */
#define SECP256K1_B200_UNIT_GEJ_DOUBLE16_W10 28 /* 16 doublings */
#define SECP256K1_B200_UNIT_GEJ_ADD_W10     29  /* the warp-wide addition (complete: infinity, a = b, a = -b) */
SECP256K1_B200_API int secp256k1_b200_unit_op_batch(secp256k1_b200_engine *e, int op, unsigned char *out,
        const unsigned char *a, const unsigned char *b, size_t n);

/* Integer-pipe peak probe: blocks*threads threads each issue 16*iters register-resident IMAD.WIDE.U32
 * (32x32+64 MAC).  d_out: blocks*threads*8 bytes.  The measured MAC/s is the roofline denominator. */
SECP256K1_B200_API int secp256k1_b200_imad_probe_dev(secp256k1_b200_engine *e, void *d_out, int blocks, int threads, int iters, void *stream);

/* ---- ecmult: out = na*A + ng*G ---------------------------------------------------------------- */
/* a64: affine x||y big-endian (all-zero = infinity); na32/ng32: big-endian scalars (reduced mod n);
 * ng32 may be NULL (= 0).  out64: affine x||y big-endian, all-zero for infinity. */
SECP256K1_B200_API int secp256k1_b200_ecmult_batch(secp256k1_b200_engine *e, unsigned char *out64,
        const unsigned char *a64, const unsigned char *na32, const unsigned char *ng32, size_t n);
SECP256K1_B200_API int secp256k1_b200_ecmult_batch_dev(secp256k1_b200_engine *e, unsigned char *d_out64,
        const unsigned char *d_a64, const unsigned char *d_na32, const unsigned char *d_ng32, size_t n, void *stream);

/* ---- BIP-340 Schnorr ---------------------------------------------------------------------------- */
/* sig64[n][64], msg[n][msglen] (same length for the whole batch), pk[n][32|33|64] per pkfmt. */
SECP256K1_B200_API int secp256k1_b200_schnorrsig_verify_batch(secp256k1_b200_engine *e, unsigned char *ok,
        const unsigned char *sig64, const unsigned char *msg, size_t msglen, const unsigned char *pk, int pkfmt, size_t n);
SECP256K1_B200_API int secp256k1_b200_schnorrsig_verify_batch_dev(secp256k1_b200_engine *e, unsigned char *d_ok,
        const unsigned char *d_sig64, const unsigned char *d_msg, size_t msglen, const unsigned char *d_pk, int pkfmt, size_t n, void *stream);

/* ---- ECDSA ------------------------------------------------------------------------------------ */
SECP256K1_B200_API int secp256k1_b200_ecdsa_verify_batch(secp256k1_b200_engine *e, unsigned char *ok,
        const unsigned char *sig64, int sigfmt, const unsigned char *msghash32, const unsigned char *pk, int pkfmt, size_t n);
SECP256K1_B200_API int secp256k1_b200_ecdsa_verify_batch_dev(secp256k1_b200_engine *e, unsigned char *d_ok,
        const unsigned char *d_sig64, int sigfmt, const unsigned char *d_msghash32, const unsigned char *d_pk, int pkfmt, size_t n, void *stream);

/* Public-key recovery: secp256k1_ecdsa_recoverable_signature_parse_compact + secp256k1_ecdsa_recover
 * (include/secp256k1_recovery.h:40,105; src/modules/recovery/main_impl.h:38,87,137).  sig64 = r||s big-endian, recid[i] in 0..3
 * (anything else: ok = 0).  pubkey64[i] = affine x||y big-endian, zeros when ok[i] == 0. */
SECP256K1_B200_API int secp256k1_b200_ecdsa_recover_batch(secp256k1_b200_engine *e, unsigned char *ok, unsigned char *pubkey64,
        const unsigned char *sig64, const unsigned char *recid, const unsigned char *msghash32, size_t n);
SECP256K1_B200_API int secp256k1_b200_ecdsa_recover_batch_dev(secp256k1_b200_engine *e, unsigned char *d_ok, unsigned char *d_pubkey64,
        const unsigned char *d_sig64, const unsigned char *d_recid, const unsigned char *d_msghash32, size_t n, void *stream);

/* BIP-340 half-aggregation: secp256k1_schnorrsig_aggverify (include/secp256k1_schnorrsig_halfagg.h:103,
 * src/modules/schnorrsig_halfagg/main_impl.h:108).  pubkeys32 = n x 32-byte x-only serializations (the reference takes parsed
 * secp256k1_xonly_pubkey objects; a key that does not parse gives *ok = 0 here), msgs32 = n x 32, aggsig = 32*(n+1) bytes
 * (r_0..r_{n-1}, s).  *ok = verdict; the return value reports engine errors only.  Evaluated as ONE multi-scalar
 * multiplication over 2n+1 terms (sum z_i*R_i + z_i*e_i*P_i - s*G == infinity).  The randomizers z_i are a running hash over all
 * signatures (serial): walked on the host for the host-buffer call, by a single GPU thread for the _dev call. */
SECP256K1_B200_API int secp256k1_b200_schnorrsig_aggverify(secp256k1_b200_engine *e, int *ok, const unsigned char *pubkeys32,
        const unsigned char *msgs32, size_t n, const unsigned char *aggsig, size_t aggsig_len);
SECP256K1_B200_API size_t secp256k1_b200_schnorrsig_aggverify_workspace_bytes(size_t n);
/* host-only helper (no GPU): the randomizers z_0..z_{n-1} of an aggregate as 32-byte hash outputs (z_0 = 1), the serial part of the
 * verification above (main_impl.h:144-155).  Returns 1 if the CPU's SHA extensions were used, 0 for the portable code;
 * force_portable != 0 selects the portable code (tests compare the two). */
SECP256K1_B200_API int secp256k1_b200_halfagg_randomizers(unsigned char *z32, const unsigned char *pubkeys32, const unsigned char *msgs32, size_t n,
        const unsigned char *aggsig, int force_portable);
SECP256K1_B200_API int secp256k1_b200_schnorrsig_aggverify_dev(secp256k1_b200_engine *e, unsigned char *d_ok, const unsigned char *d_pubkeys32,
        const unsigned char *d_msgs32, size_t n, const unsigned char *d_aggsig, void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- multi-scalar multiplication: out = g_scalar*G + sum_i scalars[i]*points[i] ----------------------------- */
/* Pippenger bucket method on the GPU (replaces static secp256k1_ecmult_multi_var, src/ecmult.h:62; the callback
 * interface becomes two flat arrays).  scalars32: n x 32 bytes big-endian (reduced mod n); points64: n x 64 bytes affine
 * x||y big-endian, all-zero = infinity; g_scalar32 may be NULL.  out64: affine x||y big-endian, all-zero = infinity.
 * The _dev variant needs a caller-provided device workspace of secp256k1_b200_ecmult_multi_workspace_bytes(n) bytes. */
SECP256K1_B200_API size_t secp256k1_b200_ecmult_multi_workspace_bytes(size_t n);
SECP256K1_B200_API int secp256k1_b200_ecmult_multi(secp256k1_b200_engine *e, unsigned char *out64, const unsigned char *g_scalar32,
        const unsigned char *scalars32, const unsigned char *points64, size_t n);
SECP256K1_B200_API int secp256k1_b200_ecmult_multi_dev(secp256k1_b200_engine *e, unsigned char *d_out64, const unsigned char *d_g_scalar32,
        const unsigned char *d_scalars32, const unsigned char *d_points64, size_t n, void *d_workspace, size_t workspace_bytes, void *stream);
/* Sharded MSM (one shard of the terms per GPU): _partial_dev leaves the shard's sum as a 96-byte Jacobian point
 * (x, y, z as 8 x u32 little-endian limbs each, z = 0 <=> infinity; no field inversion); after the partial sums of all shards have been
 * gathered (96 bytes per GPU), _combine_dev adds `count` of them and converts once: out64 = affine x||y big-endian. */
SECP256K1_B200_API int secp256k1_b200_ecmult_multi_partial_dev(secp256k1_b200_engine *e, unsigned char *d_out96, const unsigned char *d_g_scalar32,
        const unsigned char *d_scalars32, const unsigned char *d_points64, size_t n, void *d_workspace, size_t workspace_bytes, void *stream);
SECP256K1_B200_API int secp256k1_b200_ecmult_multi_combine_dev(secp256k1_b200_engine *e, unsigned char *d_out64, const unsigned char *d_partials96,
        size_t count, void *stream);
/* The other way to shard an MSM whose terms are resident on EVERY GPU: by bucket windows.  Shard `shard` of `shards` sums window j of the first GLV
 * half and window nw - 1 - j of the second for every j with j % shards == shard (nw = ceil(128 / c) windows per half) over ALL n terms and leaves a 96-byte partial result like
 * _partial_dev; the partial results of all shards go through _combine_dev.  Against sharding by terms each GPU does 1/shards of the bucket
 * additions AND of the bucket reduction (which does not shrink with the number of terms); the price is that every GPU reads all terms. */
SECP256K1_B200_API int secp256k1_b200_ecmult_multi_window_shard_dev(secp256k1_b200_engine *e, unsigned char *d_out96, const unsigned char *d_g_scalar32,
        const unsigned char *d_scalars32, const unsigned char *d_points64, size_t n, unsigned int shard, unsigned int shards,
        void *d_workspace, size_t workspace_bytes, void *stream);
/* Bucket window: 0 (default) = chosen from n like secp256k1_pippenger_bucket_window (src/ecmult_impl.h:597-621); 4..16 forces it
 * (the run-time counterpart of the reference's build-time ECMULT_WINDOW_SIZE knob; above 4096 terms at least 8).  Any value gives
 * the same result. */
SECP256K1_B200_API int secp256k1_b200_ecmult_multi_set_window(secp256k1_b200_engine *e, int window_bits);

/* ---- surjection proofs and commitment tallies (Confidential Assets) ---------------------------- */
/* secp256k1_surjectionproof_parse + _verify (include/secp256k1_surjectionproof.h:71, src/modules/surjection/main_impl.h:360) per item.
 * proofs: SERIALIZED proofs (proofs + i*proof_stride, plen[i] bytes).  Ephemeral input tags of all items are concatenated in
 * in_tags64 (64-byte secp256k1_generator objects); item i owns tags [tag_offset[i], tag_offset[i+1]) (tag_offset has n+1
 * entries).  out_tags64: one 64-byte output tag per item. */
SECP256K1_B200_API int secp256k1_b200_surjectionproof_verify_batch(secp256k1_b200_engine *e, unsigned char *ok, const unsigned char *proofs, size_t proof_stride,
        const uint32_t *plen, const unsigned char *in_tags64, const uint32_t *tag_offset, const unsigned char *out_tags64, size_t n);
SECP256K1_B200_API int secp256k1_b200_surjectionproof_verify_batch_dev(secp256k1_b200_engine *e, unsigned char *d_ok, const unsigned char *d_proofs, size_t proof_stride,
        const uint32_t *d_plen, const unsigned char *d_in_tags64, const uint32_t *d_tag_offset, const unsigned char *d_out_tags64, size_t n, void *stream);
/* secp256k1_pedersen_verify_tally (include/secp256k1_generator.h, src/modules/generator/main_impl.h:371): item i checks
 * sum(pos33[pos_offset[i] .. pos_offset[i+1])) - sum(neg33[neg_offset[i] .. neg_offset[i+1])) == infinity over 33-byte serialized
 * commitments; a commitment that does not parse rejects the item. */
SECP256K1_B200_API int secp256k1_b200_pedersen_verify_tally_batch(secp256k1_b200_engine *e, unsigned char *ok, const unsigned char *pos33, const uint32_t *pos_offset,
        const unsigned char *neg33, const uint32_t *neg_offset, size_t n);
SECP256K1_B200_API int secp256k1_b200_pedersen_verify_tally_batch_dev(secp256k1_b200_engine *e, unsigned char *d_ok, const unsigned char *d_pos33, const uint32_t *d_pos_offset,
        const unsigned char *d_neg33, const uint32_t *d_neg_offset, size_t n, void *stream);

/* ---- Bulletproofs++ norm-product argument (the only BP++ proof this reference snapshot implements) ------------- */
/* Replaces static secp256k1_bppp_rangeproof_norm_product_verify (src/modules/bppp/bppp_norm_product_impl.h:425).
 * All proofs of a batch share the geometry (g_len G-generators, h_len H-generators, both powers of two <= 64) and the generator
 * list gens64 ((g_len+h_len) x 64 bytes affine x||y big-endian, G's first).  Per proof: proof bytes (65*rounds + 64),
 * commitment (64 bytes x||y), rho (32 bytes), c_vec (h_len x 32 bytes) and the parent protocol's transcript: a 104-byte
 * SHA-256 object {8 x u32 state (little-endian words), 64-byte block buffer, u64 byte count (little-endian)} in the state
 * the reference's `secp256k1_sha256 *transcript` argument would be in.  Helpers: _sha256_state_init = a fresh SHA-256
 * (what the reference's verify vectors use, modules/bppp/tests_impl.h:548); _bppp_transcript_init = the set-up of the
 * reference's round-trip tests (modules/bppp/tests_impl.h:273-311).  Both are host-side hashing only. */
SECP256K1_B200_API void secp256k1_b200_sha256_state_init(unsigned char *state104);
SECP256K1_B200_API void secp256k1_b200_bppp_transcript_init(unsigned char *state104, const unsigned char *commit_xy64, const unsigned char *rho32,
        const unsigned char *gens64, size_t g_len, size_t n_gens, const unsigned char *c_vec32, size_t c_len);
SECP256K1_B200_API int secp256k1_b200_bppp_norm_verify_batch(secp256k1_b200_engine *e, unsigned char *ok, const unsigned char *proofs, size_t proof_stride,
        const uint32_t *plen, const unsigned char *transcripts104, const unsigned char *rho32, const unsigned char *gens64, size_t g_len,
        const unsigned char *c_vec32, size_t h_len, const unsigned char *commits64, size_t n);
SECP256K1_B200_API int secp256k1_b200_bppp_norm_verify_batch_dev(secp256k1_b200_engine *e, unsigned char *d_ok, const unsigned char *d_proofs, size_t proof_stride,
        const uint32_t *d_plen, const unsigned char *d_transcripts104, const unsigned char *d_rho32, const unsigned char *d_gens64, size_t g_len,
        const unsigned char *d_c_vec32, size_t h_len, const unsigned char *d_commits64, size_t n, void *stream);

/* ---- batched point parsing / serialization (everything that needs a field square root) ------------------ */
#define SECP256K1_B200_CODEC_XONLY32       1 /* 32-byte x-only key  -> 64-byte opaque secp256k1_xonly_pubkey (include/secp256k1_extrakeys.h:47) */
#define SECP256K1_B200_CODEC_EC_PUBKEY     2 /* 33/65-byte SEC1 key -> 64-byte opaque secp256k1_pubkey (include/secp256k1.h:462) */
#define SECP256K1_B200_CODEC_COMMIT33      3 /* 33-byte commitment  -> validity + affine x||y big-endian (include/secp256k1_generator.h:64) */
#define SECP256K1_B200_CODEC_GENERATOR33   4 /* 33-byte generator   -> 64-byte secp256k1_generator (include/secp256k1_generator.h:27) */
#define SECP256K1_B200_CODEC_GENERATOR_SER 5 /* 64-byte generator   -> 33-byte serialization in out64[0..32] (include/secp256k1_generator.h:42) */
/* item i is read from in + i*in_stride (in_len bytes) and written to out64 + 64*i; ok[i] = parse result */
SECP256K1_B200_API int secp256k1_b200_point_codec_batch(secp256k1_b200_engine *e, unsigned char *ok, unsigned char *out64,
        const unsigned char *in, size_t in_stride, uint32_t in_len, int fmt, size_t n);
SECP256K1_B200_API int secp256k1_b200_point_codec_batch_dev(secp256k1_b200_engine *e, unsigned char *d_ok, unsigned char *d_out64,
        const unsigned char *d_in, size_t in_stride, uint32_t in_len, int fmt, size_t n, void *stream);

/* ---- Pedersen commitments -------------------------------------------------------------------------- */
/* commit33[i] = serialize(blind[i]*G + value[i]*gen[i]) (tag 8/9 + x); ok[i] = 0 (and zeros) if blind >= n or the
 * result is infinity.  gen64: 64-byte secp256k1_generator objects (x||y big-endian, include/secp256k1_generator.h:24);
 * gen_stride = 0 uses one generator for the whole batch, otherwise gen64 + i*gen_stride.
 * NOT constant time in the blinding factor: for public recomputation / benchmarking. */
SECP256K1_B200_API int secp256k1_b200_pedersen_commit_batch(secp256k1_b200_engine *e, unsigned char *ok, unsigned char *commit33,
        const unsigned char *blind32, const uint64_t *value, const unsigned char *gen64, size_t gen_stride, size_t n);
SECP256K1_B200_API int secp256k1_b200_pedersen_commit_batch_dev(secp256k1_b200_engine *e, unsigned char *d_ok, unsigned char *d_commit33,
        const unsigned char *d_blind32, const uint64_t *d_value, const unsigned char *d_gen64, size_t gen_stride, size_t n, void *stream);

/* ---- range proofs ------------------------------------------------------------------------------------ */
/* One proof per item: proofs + i*proof_stride, plen[i] bytes; commit33: serialized commitments (33 bytes, as
 * secp256k1_pedersen_commitment_serialize emits / the first 33 bytes of the opaque object); generators as above;
 * extra: optional extra_commit data (extra + i*extra_stride, extra_len[i] bytes; NULL = none).
 * ok[i] = accept bit; min_value[i]/max_value[i] are written from the proof header exactly when the reference writes
 * them (also for proofs that fail later); both are 0 if the header is rejected before they are written or if the
 * commitment does not parse (the reference caller would not get past secp256k1_pedersen_commitment_parse). */
SECP256K1_B200_API int secp256k1_b200_rangeproof_verify_batch(secp256k1_b200_engine *e, unsigned char *ok, uint64_t *min_value, uint64_t *max_value,
        const unsigned char *commit33, const unsigned char *proofs, size_t proof_stride, const uint32_t *plen,
        const unsigned char *gen64, size_t gen_stride, const unsigned char *extra, size_t extra_stride, const uint32_t *extra_len, size_t n);
SECP256K1_B200_API int secp256k1_b200_rangeproof_verify_batch_dev(secp256k1_b200_engine *e, unsigned char *d_ok, uint64_t *d_min_value, uint64_t *d_max_value,
        const unsigned char *d_commit33, const unsigned char *d_proofs, size_t proof_stride, const uint32_t *d_plen,
        const unsigned char *d_gen64, size_t gen_stride, const unsigned char *d_extra, size_t extra_stride, const uint32_t *d_extra_len,
        size_t n, void *stream);

/* Rewinding: secp256k1_rangeproof_rewind (include/secp256k1_rangeproof.h:95, src/modules/rangeproof/main_impl.h:31; rangeproof_impl.h:61-112,
 * :354-471, :653-681).  Verifies like the call above and, with the 32-byte nonce the prover used, recovers per proof the blinding factor
 * (blind32), the proven value, and the embedded message (up to message_stride bytes per proof into message + i*message_stride,
 * message_len[i] = bytes recovered; message may be NULL).  ok[i] = 1 only if the proof verifies AND the recovered (value, blind)
 * rebuild the commitment; otherwise blind/value/message_len are zero.  The _dev form needs a device workspace of
 * secp256k1_b200_rangeproof_rewind_workspace_bytes(n) bytes (ring challenges and the replayed random stream, 9216 per proof). */
SECP256K1_B200_API size_t secp256k1_b200_rangeproof_rewind_workspace_bytes(size_t n);
SECP256K1_B200_API int secp256k1_b200_rangeproof_rewind_batch(secp256k1_b200_engine *e, unsigned char *ok, unsigned char *blind32, uint64_t *value,
        unsigned char *message, size_t message_stride, uint32_t *message_len, uint64_t *min_value, uint64_t *max_value,
        const unsigned char *nonce32, const unsigned char *commit33, const unsigned char *proofs, size_t proof_stride, const uint32_t *plen,
        const unsigned char *gen64, size_t gen_stride, const unsigned char *extra, size_t extra_stride, const uint32_t *extra_len, size_t n);
SECP256K1_B200_API int secp256k1_b200_rangeproof_rewind_batch_dev(secp256k1_b200_engine *e, unsigned char *d_ok, unsigned char *d_blind32, uint64_t *d_value,
        unsigned char *d_message, size_t message_stride, uint32_t *d_message_len, uint64_t *d_min_value, uint64_t *d_max_value,
        const unsigned char *d_nonce32, const unsigned char *d_commit33, const unsigned char *d_proofs, size_t proof_stride, const uint32_t *d_plen,
        const unsigned char *d_gen64, size_t gen_stride, const unsigned char *d_extra, size_t extra_stride, const uint32_t *d_extra_len, size_t n,
        void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- one process, several GPUs -------------------------------------------------------------------------------------------
 * The reference has no counterpart (it is a single-threaded library); this is the C-level form of north_star's multi-GPU plan:
 * _multi_create makes one engine per listed device (devices = NULL: devices 0 .. n_devices-1; n_devices <= 0: every device of the box),
 * device 0 builds the generator table, ncclCommInitAll + ONE ncclBroadcast replicate it (NCCL is bound at run time from libnccl.so.2 and is
 * only needed for more than one device).  Every *_batch call below has the semantics of its single-engine namesake: the host arrays are cut
 * into contiguous index ranges, one per device, processed concurrently (one host thread and one stream pair per device), results written
 * straight into the caller's arrays.  _multi_ecmult_multi shards the terms, gathers the per-device Jacobian partial sums with ONE
 * ncclAllGather (96 bytes per device) and adds them on device 0.  A multi engine may be used from several host threads (calls serialize).
 * _multi_engine exposes the per-device engines for the device-pointer (_dev) entry points; do not run a sharded call concurrently with
 * host-buffer calls on those engines. */
typedef struct secp256k1_b200_multi secp256k1_b200_multi;
SECP256K1_B200_API int secp256k1_b200_multi_create(secp256k1_b200_multi **out, const int *devices, int n_devices, int gen_window_bits);
SECP256K1_B200_API void secp256k1_b200_multi_destroy(secp256k1_b200_multi *m);
SECP256K1_B200_API int secp256k1_b200_multi_device_count(const secp256k1_b200_multi *m);
SECP256K1_B200_API int secp256k1_b200_multi_nccl_version(const secp256k1_b200_multi *m);   /* e.g. 22809; 0 with a single device */
SECP256K1_B200_API secp256k1_b200_engine *secp256k1_b200_multi_engine(secp256k1_b200_multi *m, int index);
SECP256K1_B200_API uint64_t secp256k1_b200_multi_launch_count(const secp256k1_b200_multi *m);
SECP256K1_B200_API int secp256k1_b200_multi_schnorrsig_verify_batch(secp256k1_b200_multi *m, unsigned char *ok, const unsigned char *sig64,
        const unsigned char *msg, size_t msglen, const unsigned char *pk, int pkfmt, size_t n);
SECP256K1_B200_API int secp256k1_b200_multi_ecdsa_verify_batch(secp256k1_b200_multi *m, unsigned char *ok, const unsigned char *sig64, int sigfmt,
        const unsigned char *msghash32, const unsigned char *pk, int pkfmt, size_t n);
SECP256K1_B200_API int secp256k1_b200_multi_ecmult_batch(secp256k1_b200_multi *m, unsigned char *out64, const unsigned char *a64,
        const unsigned char *na32, const unsigned char *ng32, size_t n);
SECP256K1_B200_API int secp256k1_b200_multi_pedersen_commit_batch(secp256k1_b200_multi *m, unsigned char *ok, unsigned char *commit33,
        const unsigned char *blind32, const uint64_t *value, const unsigned char *gen64, size_t gen_stride, size_t n);
SECP256K1_B200_API int secp256k1_b200_multi_rangeproof_verify_batch(secp256k1_b200_multi *m, unsigned char *ok, uint64_t *min_value, uint64_t *max_value,
        const unsigned char *commit33, const unsigned char *proofs, size_t proof_stride, const uint32_t *plen, const unsigned char *gen64, size_t gen_stride,
        const unsigned char *extra, size_t extra_stride, const uint32_t *extra_len, size_t n);
SECP256K1_B200_API int secp256k1_b200_multi_bppp_norm_verify_batch(secp256k1_b200_multi *m, unsigned char *ok, const unsigned char *proofs, size_t proof_stride,
        const uint32_t *plen, const unsigned char *transcripts104, const unsigned char *rho32, const unsigned char *gens64, size_t g_len,
        const unsigned char *c_vec32, size_t h_len, const unsigned char *commits64, size_t n);
SECP256K1_B200_API int secp256k1_b200_multi_ecmult_multi(secp256k1_b200_multi *m, unsigned char *out64, const unsigned char *g_scalar32,
        const unsigned char *scalars32, const unsigned char *points64, size_t n);

/* ---- synthetic inputs for tests and benchmarks (NOT constant time; never use with real secrets) ---- */
/* BIP-340 signatures exactly as secp256k1_schnorrsig_sign_custom(ctx, sig, msg, msglen, keypair, NULL) emits them
 * (src/modules/schnorrsig/main_impl.h:125-190), plus the x-only public keys; sk must be in [1, n-1]. */
SECP256K1_B200_API int secp256k1_b200_testdata_schnorrsig_sign_batch(secp256k1_b200_engine *e, unsigned char *sig64, unsigned char *pk32,
        const unsigned char *sk32, const unsigned char *msg, size_t msglen, size_t n);
SECP256K1_B200_API int secp256k1_b200_testdata_schnorrsig_sign_batch_dev(secp256k1_b200_engine *e, unsigned char *d_sig64, unsigned char *d_pk32,
        const unsigned char *d_sk32, const unsigned char *d_msg, size_t msglen, size_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SECP256K1_B200_H */
