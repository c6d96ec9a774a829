/* oracle/secp_oracle.h -- TEST INFRASTRUCTURE: plain-C CPU restatement ("port") of the reference's
 * algorithms for the accelerated hot path.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs
 * may load the library built from this; the product never does.
 *
 * Parity of this restatement is PINNED: tests/test_oracle_port.py checks it against the reference's own
 * golden vectors (the JSON files under tests/golden: BIP-340, Wycheproof ECDSA, run_ecmult_constants digests,
 * run_ecmult_chain endpoint, rangeproof vectors, BP++ verify vectors) and, where oracle/_ref exists,
 * against the unmodified reference on random inputs.
 *
 * Byte conventions: field elements / scalars are 32-byte big-endian; points are 64-byte x||y big-endian
 * with all-zero meaning infinity.
 */
#ifndef SECP_ORACLE_H
#define SECP_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#define ORC_API __attribute__((visibility("default")))

ORC_API void orc_fe_mul(unsigned char *r32, const unsigned char *a32, const unsigned char *b32);
ORC_API void orc_fe_inv(unsigned char *r32, const unsigned char *a32);
ORC_API int  orc_fe_sqrt(unsigned char *r32, const unsigned char *a32);
ORC_API void orc_scalar_mul(unsigned char *r32, const unsigned char *a32, const unsigned char *b32);
ORC_API void orc_scalar_inverse(unsigned char *r32, const unsigned char *a32);
ORC_API void orc_sha256(unsigned char *out32, const unsigned char *data, size_t len);

ORC_API int  orc_ecmult(unsigned char *r64, const unsigned char *a64, const unsigned char *na32, const unsigned char *ng32);
ORC_API void orc_ecmult_batch(unsigned char *r64, const unsigned char *a64, const unsigned char *na32, const unsigned char *ng32, size_t n);
ORC_API int  orc_ecmult_chain(unsigned char *r64, const unsigned char *a64, const unsigned char *xn32, const unsigned char *gn32,
                              const unsigned char *xf32, const unsigned char *gf32, size_t iterations);
ORC_API int  orc_ecmult_multi(unsigned char *r64, const unsigned char *g_sc32, const unsigned char *sc32, const unsigned char *pt64, size_t n);

ORC_API int  orc_xonly_pubkey_parse(unsigned char *xy64, const unsigned char *pk32);
ORC_API int  orc_ec_pubkey_parse(unsigned char *xy64, const unsigned char *pk, size_t len);
ORC_API int  orc_schnorrsig_verify(const unsigned char *sig64, const unsigned char *msg, size_t msglen, const unsigned char *pk32);
ORC_API void orc_schnorrsig_verify_batch(unsigned char *ok, const unsigned char *sig64, const unsigned char *msg32, const unsigned char *pk32, size_t n);
ORC_API int  orc_schnorrsig_aggverify(const unsigned char *pk32, const unsigned char *msgs32, size_t n, const unsigned char *aggsig, size_t aggsig_len);
ORC_API int  orc_ecdsa_recover(unsigned char *pubkey64, const unsigned char *sig64, int recid, const unsigned char *msghash32);
ORC_API int  orc_ecdsa_verify(const unsigned char *sig64_compact, const unsigned char *msghash32, const unsigned char *pk, size_t pklen);

ORC_API int  orc_pedersen_commit(unsigned char *commit33, const unsigned char *blind32, uint64_t value, const unsigned char *gen64);
ORC_API int  orc_pedersen_commitment_parse(unsigned char *xy64, const unsigned char *commit33);
ORC_API int  orc_generator_parse(unsigned char *gen64, const unsigned char *gen33);
ORC_API int  orc_rangeproof_verify(uint64_t *min_value, uint64_t *max_value, const unsigned char *commit33,
                                   const unsigned char *proof, size_t plen, const unsigned char *extra, size_t extra_len,
                                   const unsigned char *gen64);
ORC_API int  orc_rangeproof_rewind(unsigned char *blind_out32, uint64_t *value_out, unsigned char *message_out, size_t *outlen, const unsigned char *nonce32,
                          uint64_t *min_value, uint64_t *max_value, const unsigned char *commit33, const unsigned char *proof, size_t plen,
                          const unsigned char *extra, size_t extra_len, const unsigned char *gen64);

ORC_API int  orc_surjectionproof_verify(const unsigned char *proof, size_t plen, const unsigned char *in_tags64, size_t n_tags, const unsigned char *out_tag64);
ORC_API int  orc_pedersen_verify_tally(const unsigned char *pos33, size_t pcnt, const unsigned char *neg33, size_t ncnt);
ORC_API int  orc_bppp_norm_verify(const unsigned char *proof, size_t plen, const unsigned char *commit_xy64, const unsigned char *rho32,
                                  const unsigned char *gens_xy64, size_t g_len, const unsigned char *c_vec32, size_t h_len);
#endif
