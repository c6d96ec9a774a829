// launch.h -- internal host-side declarations of the kernel launchers (one per .cu file).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200 {

struct bip340_midstate_h { uint32_t s[8]; };

struct sign_consts_h { uint32_t mid_challenge[8]; uint32_t mid_nonce[8]; uint8_t zero_mask[32]; };

struct gtable_dev {
    const uint32_t* data;
    int gw;
    int windows;
};

// every launcher returns the number of kernel launches it issued through *launches (added)
cudaError_t launch_fe_op(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, int iters, cudaStream_t s, uint64_t* launches);
cudaError_t launch_imad_probe(uint64_t* out, int blocks, int threads, int iters, cudaStream_t s, uint64_t* launches);
cudaError_t launch_fe_op_portable(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, int iters, cudaStream_t s, uint64_t* launches);
cudaError_t launch_gtable_build(uint32_t* d_tab, uint32_t* d_bases, int gw, int windows, cudaStream_t s, uint64_t* launches);
cudaError_t launch_ecmult(uint8_t* out64, const uint8_t* a64, const uint8_t* na32, const uint8_t* ng32, size_t n,
                          const gtable_dev& gt, cudaStream_t s, uint64_t* launches);
cudaError_t launch_schnorr_verify(uint8_t* ok, const uint8_t* sig64, const uint8_t* msg, uint32_t msglen, const uint8_t* pk, int pkfmt,
                                  size_t n, const bip340_midstate_h& mid, const gtable_dev& gt, cudaStream_t s, uint64_t* launches);
cudaError_t launch_ecdsa_recover(uint8_t* ok, uint8_t* out64, const uint8_t* sig64, const uint8_t* recid, const uint8_t* msg32, size_t n,
                                 const gtable_dev& gt, cudaStream_t s, uint64_t* launches);
cudaError_t launch_ecdsa_verify(uint8_t* ok, const uint8_t* sig64, int sigfmt, const uint8_t* msg32, const uint8_t* pk, int pkfmt,
                                size_t n, const gtable_dev& gt, cudaStream_t s, uint64_t* launches);

cudaError_t launch_schnorr_sign(uint8_t* sig64, uint8_t* pk32, uint8_t* ok, const uint8_t* sk32, const uint8_t* msg, uint32_t msglen, size_t n,
                                const sign_consts_h& c, const gtable_dev& gt, cudaStream_t s, uint64_t* launches);

cudaError_t launch_rangeproof_verify(uint8_t* ok, uint64_t* minv, uint64_t* maxv, const uint8_t* commit33, const uint8_t* proofs, size_t stride,
                                     const uint32_t* plen, const uint8_t* gen64, size_t gen_stride, const uint8_t* extra, size_t extra_stride,
                                     const uint32_t* extra_len, size_t n, const gtable_dev& gt, cudaStream_t s, uint64_t* launches, uint8_t* ev_out = nullptr,
                                     cudaMemPool_t pool = nullptr /* scratch of the three launches; nullptr = the device's default pool */);
cudaError_t launch_rangeproof_rewind(uint8_t* ok, uint8_t* blind32, uint64_t* value, uint8_t* msg, size_t msg_stride, uint32_t* msg_len,
                                     const uint8_t* nonce32, const uint8_t* commit33, const uint8_t* proofs, size_t stride, const uint32_t* plen,
                                     const uint8_t* gen64, size_t gen_stride, const uint8_t* ev, uint8_t* scratch, size_t n, const gtable_dev& gt,
                                     cudaStream_t s, uint64_t* launches);
cudaError_t launch_pedersen_commit(uint8_t* ok, uint8_t* out33, const uint8_t* blind32, const uint64_t* value, const uint8_t* gen64, size_t gen_stride,
                                   size_t n, const gtable_dev& gt, cudaStream_t s, uint64_t* launches);

cudaError_t launch_codec(uint8_t* ok, uint8_t* out64, const uint8_t* in, size_t in_stride, uint32_t in_len, int fmt, size_t n,
                         cudaStream_t s, uint64_t* launches);

cudaError_t launch_bppp_norm_verify(uint8_t* ok, const uint8_t* proofs, size_t stride, const uint32_t* plen, const uint8_t* transcripts,
                                    const uint8_t* rho32, const uint8_t* gens64, uint32_t g_len, const uint8_t* c_vec32, uint32_t h_len,
                                    const uint8_t* commits64, size_t n, const gtable_dev& gt, cudaStream_t s, uint64_t* launches, cudaMemPool_t pool = nullptr /* generator tables of large batches */);
cudaError_t launch_surjection_verify(uint8_t* ok, const uint8_t* proofs, size_t stride, const uint32_t* plen, const uint8_t* in_tags64,
                                     const uint32_t* tag_off, const uint8_t* out_tags64, size_t n, const gtable_dev& gt, cudaStream_t s, uint64_t* launches);
cudaError_t launch_pedersen_tally(uint8_t* ok, const uint8_t* pos33, const uint32_t* pos_off, const uint8_t* neg33, const uint32_t* neg_off, size_t n,
                                  cudaStream_t s, uint64_t* launches);
size_t halfagg_workspace_bytes(size_t n);
cudaError_t launch_halfagg_verify(uint8_t* d_ok, const uint8_t* d_pk32, const uint8_t* d_msgs32, size_t n, const uint8_t* d_aggsig,
                                  const uint8_t* d_z32_in, const uint32_t* hmid, const bip340_midstate_h& mid, uint8_t* ws, size_t ws_bytes,
                                  cudaStream_t s, uint64_t* launches);
size_t msm_workspace_bytes(size_t n_terms);
// out_mode 0: out = 64 bytes affine x||y big-endian; 1: out = 96 bytes Jacobian (x, y, z LE limbs, z = 0 <=> infinity) for launch_gej_sum.
// window_bits 0 = chosen from n (msm_choose_c).  shards > 1: only the bucket windows of shard `shard` are summed (msm_window_owned; window sharding:
// every shard sees ALL terms; the Jacobian partial results of the shards add up to the MSM)
cudaError_t launch_msm(uint8_t* out, int out_mode, const uint8_t* g_scalar32, const uint8_t* scalars32, const uint8_t* points64, size_t n,
                       uint8_t* workspace, size_t workspace_bytes, int window_bits, cudaStream_t s, uint64_t* launches, uint32_t shard = 0, uint32_t shards = 1);
cudaError_t launch_unit_op(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, cudaStream_t s, uint64_t* launches);
cudaError_t launch_gej_sum(uint8_t* out64, const uint8_t* partials96, size_t count, cudaStream_t s, uint64_t* launches);

}  // namespace b200
