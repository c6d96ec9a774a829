// k_verify.cu -- one-thread-per-item kernels: ecmult batch, BIP-340 verify batch, ECDSA verify batch.
// Block = 128 threads, 5 blocks/SM at 96 registers; the per-thread tables of odd multiples (see ecmult.cuh) are in local memory,
// dynamic shared memory (8 KiB) only holds the product trees of the block-wide inversions -- see the note at VERIFY_BLOCK below.
#include "verify.cuh"
#include "launch.h"

namespace b200 {

#define VERIFY_BLOCK 128
// The 8-entry odd-multiples table of each thread (512 bytes) lives in LOCAL memory (per-thread, interleaved across the warp by the
// hardware, cached in L1), not in shared memory: with the table in shared memory an SM holds 3 blocks (12 warps at 158 registers);
// in local memory the limit is the register file, and 5 blocks of 96-register threads (20 warps, ~1 KB of spills per thread) hide the
// latency of the multiply chains better.  Measured (tools/ab_bench.py, 2^20 verifies): shared/3 blocks 40.0 M/s; local/3 40.2;
// local/4 (128 regs) 41.9; local/5 (96) 43.2; local/6 (80) 42.6; local/7 41.7; local/8 40.6.
// Shared memory is only used by the block-wide inversions (2 * VERIFY_BLOCK field elements).
#ifdef B200_TAB_IN_SHARED         // the r1a-r1g layout, kept for A/B runs
#define VERIFY_SMEM (VERIFY_BLOCK * 8 * 16 * 4)
#define VERIFY_TAB(name) tab_shared name = {smem + threadIdx.x, (int)blockDim.x}
#define B200_VERIFY_MINBLOCKS 3
#else
#define VERIFY_SMEM (VERIFY_BLOCK * 2 * 32)
#define VERIFY_TAB(name) tab_local name
#ifndef B200_VERIFY_MINBLOCKS
#define B200_VERIFY_MINBLOCKS 5
#endif
#endif

__device__ __forceinline__ void load_bytes_global(uint8_t* dst, const uint8_t* src, int nbytes) {
    // 16-byte vector path when aligned (contiguous per-item records), byte path otherwise
    if (((reinterpret_cast<uintptr_t>(src) | (uintptr_t)nbytes) & 15) == 0) {
        const uint4* q = reinterpret_cast<const uint4*>(src);
        uint4* d = reinterpret_cast<uint4*>(dst);
        for (int i = 0; i < nbytes / 16; i++) d[i] = __ldg(q + i);
    } else {
        for (int i = 0; i < nbytes; i++) dst[i] = __ldg(src + i);
    }
}

// One field inversion per thread block (Montgomery's trick as a product tree in shared memory).  Every thread of the
// block must call it (z = 1 for threads without work); returns 1/z.  `nodes` needs 2*blockDim.x field elements and may
// alias the per-thread tables (they are dead by now).  Cost per 128-thread block: ~24 warp-level multiplications + ONE
// 270-multiplication inversion executed by a single warp, instead of 4 warps x 270.
__device__ __forceinline__ void fe_block_inverse(fe& zinv, const fe& z, fe* nodes) {
    const uint32_t t = threadIdx.x, nt = blockDim.x;      // nt is a power of two
    __syncthreads();                                        // table use finished everywhere
    nodes[nt + t] = z;
    __syncthreads();
    for (uint32_t s = nt >> 1; s >= 1; s >>= 1) {
        if (t < s) { fe r; fe_mul(r, nodes[2 * (s + t)], nodes[2 * (s + t) + 1]); nodes[s + t] = r; }
        __syncthreads();
    }
    if (t == 0) { fe r; fe_inv(r, nodes[1]); nodes[1] = r; }
    __syncthreads();
    for (uint32_t s = 1; s < nt; s <<= 1) {
        if (t < s) {
            fe inv = nodes[s + t], l = nodes[2 * (s + t)], r = nodes[2 * (s + t) + 1], a, b;
            fe_mul(a, inv, r);
            fe_mul(b, inv, l);
            nodes[2 * (s + t)] = a;
            nodes[2 * (s + t) + 1] = b;
        }
        __syncthreads();
    }
    zinv = nodes[nt + t];
}

// same tree for scalars mod n (ECDSA's s^-1): every thread calls it (1 for threads without work)
__device__ __forceinline__ void sc_block_inverse(sc& sinv, const sc& x, sc* nodes) {
    const uint32_t t = threadIdx.x, nt = blockDim.x;
    __syncthreads();
    nodes[nt + t] = x;
    __syncthreads();
    for (uint32_t s = nt >> 1; s >= 1; s >>= 1) {
        if (t < s) { sc r; sc_mul(r, nodes[2 * (s + t)], nodes[2 * (s + t) + 1]); nodes[s + t] = r; }
        __syncthreads();
    }
    if (t == 0) { sc r; sc_inverse(r, nodes[1]); nodes[1] = r; }
    __syncthreads();
    for (uint32_t s = 1; s < nt; s <<= 1) {
        if (t < s) {
            sc inv = nodes[s + t], l = nodes[2 * (s + t)], r = nodes[2 * (s + t) + 1], a, b;
            sc_mul(a, inv, r);
            sc_mul(b, inv, l);
            nodes[2 * (s + t)] = a;
            nodes[2 * (s + t) + 1] = b;
        }
        __syncthreads();
    }
    sinv = nodes[nt + t];
    __syncthreads();                                        // the region is about to become the tables
}

__global__ void __launch_bounds__(VERIFY_BLOCK, B200_VERIFY_MINBLOCKS) k_ecmult(uint8_t* __restrict__ out64, const uint8_t* __restrict__ a64,
        const uint8_t* __restrict__ na32, const uint8_t* __restrict__ ng32, size_t n, ecmult_gtable gt) {
    extern __shared__ uint32_t smem[];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    VERIFY_TAB(tab);
    gej R; gej_set_infinity(R);
    if (i < n) {
        __align__(16) uint8_t in[128];
        load_bytes_global(in, a64 + 64 * i, 64);
        load_bytes_global(in + 64, na32 + 32 * i, 32);
        if (ng32) load_bytes_global(in + 96, ng32 + 32 * i, 32);
        ge a;
        int z = 1;
        for (int k = 0; k < 64; k++) z &= (in[k] == 0);
        if (z) ge_set_infinity(a);
        else { fe_set_b32(a.x, in); fe_set_b32(a.y, in + 32); a.inf = 0; }
        sc na, ng;
        sc_set_b32(na, in + 64);
        if (ng32) sc_set_b32(ng, in + 96); else sc_set_zero(ng);
        gej aj;
        gej_set_ge(aj, a);
        ecmult(R, aj, na, ng, gt, tab);
    }
    fe zz, zinv;
    // a64 is taken as given (like the reference's static secp256k1_ecmult); an off-curve input can reach Z = 0 with inf = 0.  Such a result
    // is reported as infinity and must not enter the inversion this block shares (a zero factor would zero every other item's 1/z).
    if (!R.inf && fe_is_zero(R.z)) gej_set_infinity(R);
    if (!R.inf) zz = R.z; else fe_set_int(zz, 1);
    fe_block_inverse(zinv, zz, reinterpret_cast<fe*>(smem));
    if (i >= n) return;
    __align__(16) uint8_t o[64];
    if (R.inf) {
        for (int k = 0; k < 64; k++) o[k] = 0;
    } else {
        ge Ra; ge_set_gej_zinv(Ra, R, zinv);
        fe_get_b32(o, Ra.x); fe_get_b32(o + 32, Ra.y);
    }
    uint4* q = reinterpret_cast<uint4*>(out64 + 64 * i);
    const uint4* s4 = reinterpret_cast<const uint4*>(o);
    if ((reinterpret_cast<uintptr_t>(out64) & 15) == 0) { q[0] = s4[0]; q[1] = s4[1]; q[2] = s4[2]; q[3] = s4[3]; }
    else for (int k = 0; k < 64; k++) out64[64 * i + k] = o[k];
}

__global__ void __launch_bounds__(VERIFY_BLOCK, B200_VERIFY_MINBLOCKS) k_schnorr_verify(uint8_t* __restrict__ ok, const uint8_t* __restrict__ sig64,
        const uint8_t* __restrict__ msg, uint32_t msglen, const uint8_t* __restrict__ pk, int pkfmt, size_t n,
        bip340_midstate mid, ecmult_gtable gt) {
    extern __shared__ uint32_t smem[];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    VERIFY_TAB(tab);
    gej R; fe rx, z, zinv;
    int live = 0;
    if (i < n) {
        __align__(16) uint8_t in[64 + 64];
        load_bytes_global(in, sig64 + 64 * i, 64);
        int pklen = pkfmt == PK_OPAQUE64 ? 64 : (pkfmt == PK_XONLY32 ? 32 : 33);
        load_bytes_global(in + 64, pk + (size_t)pklen * i, pklen);
        // messages of arbitrary (batch-uniform) length are hashed straight from global memory
        live = schnorr_verify_prepare(R, rx, in, msg + (size_t)msglen * i, msglen, in + 64, pkfmt, mid, gt, tab);
    }
    if (live && fe_is_zero(R.z)) live = 0;                    // defence in depth: never feed a zero into the shared inversion
    if (live) z = R.z; else fe_set_int(z, 1);
    fe_block_inverse(zinv, z, reinterpret_cast<fe*>(smem));   // all threads of the block take part
    if (i < n) ok[i] = (uint8_t)(live && schnorr_verify_finish(R, rx, zinv));
}

__global__ void __launch_bounds__(VERIFY_BLOCK, B200_VERIFY_MINBLOCKS) k_ecdsa_verify(uint8_t* __restrict__ ok, const uint8_t* __restrict__ sig64, int sigfmt,
        const uint8_t* __restrict__ msg32, const uint8_t* __restrict__ pk, int pkfmt, size_t n, ecmult_gtable gt) {
    extern __shared__ uint32_t smem[];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    VERIFY_TAB(tab);
    ecdsa_state st;
    int live = 0;
    if (i < n) {
        __align__(16) uint8_t in[64 + 32 + 64];
        load_bytes_global(in, sig64 + 64 * i, 64);
        load_bytes_global(in + 64, msg32 + 32 * i, 32);
        int pklen = pkfmt == PK_OPAQUE64 ? 64 : (pkfmt == PK_XONLY32 ? 32 : 33);
        load_bytes_global(in + 96, pk + (size_t)pklen * i, pklen);
        live = ecdsa_verify_parse(st, in, sigfmt, in + 64, in + 96, pkfmt);
    }
    sc x, sn;
    if (live) x = st.s; else sc_set_int(x, 1);
    sc_block_inverse(sn, x, reinterpret_cast<sc*>(smem));     // one inversion mod n per block
    if (i < n) ok[i] = (uint8_t)(live && ecdsa_verify_finish(st, sn, gt, tab));
}

// out64[i] = recovered public key (affine x||y big-endian; zeros when ok[i] == 0)
__global__ void __launch_bounds__(VERIFY_BLOCK, B200_VERIFY_MINBLOCKS) k_ecdsa_recover(uint8_t* __restrict__ ok, uint8_t* __restrict__ out64, const uint8_t* __restrict__ sig64,
        const uint8_t* __restrict__ recid, const uint8_t* __restrict__ msg32, size_t n, ecmult_gtable gt) {
    extern __shared__ uint32_t smem[];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    VERIFY_TAB(tab);
    ecdsa_recover_state st;
    int live = 0;
    if (i < n) {
        __align__(16) uint8_t in[64 + 32];
        load_bytes_global(in, sig64 + 64 * i, 64);
        load_bytes_global(in + 64, msg32 + 32 * i, 32);
        int rid = recid[i];
        live = rid <= 3 && ecdsa_recover_parse(st, in, rid, in + 64);
    }
    sc x, rn;
    if (live) x = st.r; else sc_set_int(x, 1);
    sc_block_inverse(rn, x, reinterpret_cast<sc*>(smem));     // one inversion mod n per block
    gej Q; gej_set_infinity(Q);
    if (live) live = ecdsa_recover_finish(Q, st, rn, gt, tab);
    fe z, zinv;
    if (live && fe_is_zero(Q.z)) live = 0;
    if (live) z = Q.z; else fe_set_int(z, 1);
    fe_block_inverse(zinv, z, reinterpret_cast<fe*>(smem));   // one field inversion per block
    if (i >= n) return;
    __align__(16) uint8_t o[64];
    if (!live) {
        for (int k = 0; k < 64; k++) o[k] = 0;
    } else {
        ge Qa; ge_set_gej_zinv(Qa, Q, zinv);
        fe_get_b32(o, Qa.x); fe_get_b32(o + 32, Qa.y);
    }
    ok[i] = (uint8_t)live;
    uint4* q = reinterpret_cast<uint4*>(out64 + 64 * i);
    const uint4* s4 = reinterpret_cast<const uint4*>(o);
    if ((reinterpret_cast<uintptr_t>(out64) & 15) == 0) { q[0] = s4[0]; q[1] = s4[1]; q[2] = s4[2]; q[3] = s4[3]; }
    else for (int k = 0; k < 64; k++) out64[64 * i + k] = o[k];
}

template <class K>
static cudaError_t set_smem(K kernel) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VERIFY_SMEM);
}
static inline ecmult_gtable to_gt(const gtable_dev& g) { ecmult_gtable t; t.data = g.data; t.gw = g.gw; t.windows = g.windows; return t; }

cudaError_t launch_ecmult(uint8_t* out64, const uint8_t* a64, const uint8_t* na32, const uint8_t* ng32, size_t n,
                          const gtable_dev& gt, cudaStream_t s, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    cudaError_t err = set_smem(k_ecmult);
    if (err != cudaSuccess) return err;
    k_ecmult<<<(unsigned)((n + VERIFY_BLOCK - 1) / VERIFY_BLOCK), VERIFY_BLOCK, VERIFY_SMEM, s>>>(out64, a64, na32, ng32, n, to_gt(gt));
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}
cudaError_t launch_schnorr_verify(uint8_t* ok, const uint8_t* sig64, const uint8_t* msg, uint32_t msglen, const uint8_t* pk, int pkfmt,
                                  size_t n, const bip340_midstate_h& mid, const gtable_dev& gt, cudaStream_t s, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    cudaError_t err = set_smem(k_schnorr_verify);
    if (err != cudaSuccess) return err;
    bip340_midstate m;
    for (int i = 0; i < 8; i++) m.s[i] = mid.s[i];
    k_schnorr_verify<<<(unsigned)((n + VERIFY_BLOCK - 1) / VERIFY_BLOCK), VERIFY_BLOCK, VERIFY_SMEM, s>>>(ok, sig64, msg, msglen, pk, pkfmt, n, m, to_gt(gt));
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}
cudaError_t launch_ecdsa_verify(uint8_t* ok, const uint8_t* sig64, int sigfmt, const uint8_t* msg32, const uint8_t* pk, int pkfmt,
                                size_t n, const gtable_dev& gt, cudaStream_t s, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    cudaError_t err = set_smem(k_ecdsa_verify);
    if (err != cudaSuccess) return err;
    k_ecdsa_verify<<<(unsigned)((n + VERIFY_BLOCK - 1) / VERIFY_BLOCK), VERIFY_BLOCK, VERIFY_SMEM, s>>>(ok, sig64, sigfmt, msg32, pk, pkfmt, n, to_gt(gt));
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

cudaError_t launch_ecdsa_recover(uint8_t* ok, uint8_t* out64, const uint8_t* sig64, const uint8_t* recid, const uint8_t* msg32, size_t n,
                                 const gtable_dev& gt, cudaStream_t s, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    cudaError_t err = set_smem(k_ecdsa_recover);
    if (err != cudaSuccess) return err;
    k_ecdsa_recover<<<(unsigned)((n + VERIFY_BLOCK - 1) / VERIFY_BLOCK), VERIFY_BLOCK, VERIFY_SMEM, s>>>(ok, out64, sig64, recid, msg32, n, to_gt(gt));
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

}  // namespace b200
