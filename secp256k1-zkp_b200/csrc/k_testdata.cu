// k_testdata.cu -- kernels behind the secp256k1_b200_testdata_* entry points (synthetic inputs only).
#include "testdata.cuh"
#include "launch.h"

namespace b200 {

__global__ void __launch_bounds__(128) k_schnorr_sign(uint8_t* __restrict__ sig64, uint8_t* __restrict__ pk32, uint8_t* __restrict__ ok,
        const uint8_t* __restrict__ sk32, const uint8_t* __restrict__ msg, uint32_t msglen, size_t n, bip340_sign_consts c, ecmult_gtable gt) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t sk[32], sig[64], pk[32];
    for (int k = 0; k < 32; k++) sk[k] = sk32[32 * i + k];
    int r = schnorr_sign_item(sig, pk, sk, msg + (size_t)msglen * i, msglen, c, gt);
    for (int k = 0; k < 64; k++) sig64[64 * i + k] = sig[k];
    for (int k = 0; k < 32; k++) pk32[32 * i + k] = pk[k];
    if (ok) ok[i] = (uint8_t)r;
}

cudaError_t launch_schnorr_sign(uint8_t* sig64, uint8_t* pk32, uint8_t* ok, const uint8_t* sk32, const uint8_t* msg, uint32_t msglen, size_t n,
                                const sign_consts_h& c, const gtable_dev& gt, cudaStream_t s, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    bip340_sign_consts cc;
    for (int i = 0; i < 8; i++) { cc.mid_challenge[i] = c.mid_challenge[i]; cc.mid_nonce[i] = c.mid_nonce[i]; }
    for (int i = 0; i < 32; i++) cc.zero_mask[i] = c.zero_mask[i];
    ecmult_gtable g; g.data = gt.data; g.gw = gt.gw; g.windows = gt.windows;
    k_schnorr_sign<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(sig64, pk32, ok, sk32, msg, msglen, n, cc, g);
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

}  // namespace b200
