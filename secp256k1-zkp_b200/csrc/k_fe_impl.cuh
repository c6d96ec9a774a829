// k_fe_impl.cuh -- batched field-op kernel body; included by k_fe.cu (PTX limb code) and by
// k_fe_portable.cu (compiled with -DB200_PORTABLE_BN -DB200_NS=b200_portable) so that both limb
// implementations exist in one library and can be compared ON THE DEVICE.
#include "modinv.cuh"
#include "launch.h"

namespace B200_NS {

__device__ __forceinline__ void fe_load_be32_global(fe& r, const uint8_t* p) {
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 hi = __ldg(q), lo = __ldg(q + 1);   // big-endian: first 16 bytes are the most significant
        r.v[7] = __byte_perm(hi.x, 0, 0x0123); r.v[6] = __byte_perm(hi.y, 0, 0x0123);
        r.v[5] = __byte_perm(hi.z, 0, 0x0123); r.v[4] = __byte_perm(hi.w, 0, 0x0123);
        r.v[3] = __byte_perm(lo.x, 0, 0x0123); r.v[2] = __byte_perm(lo.y, 0, 0x0123);
        r.v[1] = __byte_perm(lo.z, 0, 0x0123); r.v[0] = __byte_perm(lo.w, 0, 0x0123);
    } else {
        bn_from_be32(r.v, p);
    }
}
__device__ __forceinline__ void fe_store_be32_global(uint8_t* p, const fe& a) {
    fe t = a; fe_normalize(t);
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        uint4 hi, lo;
        hi.x = __byte_perm(t.v[7], 0, 0x0123); hi.y = __byte_perm(t.v[6], 0, 0x0123);
        hi.z = __byte_perm(t.v[5], 0, 0x0123); hi.w = __byte_perm(t.v[4], 0, 0x0123);
        lo.x = __byte_perm(t.v[3], 0, 0x0123); lo.y = __byte_perm(t.v[2], 0, 0x0123);
        lo.z = __byte_perm(t.v[1], 0, 0x0123); lo.w = __byte_perm(t.v[0], 0, 0x0123);
        uint4* q = reinterpret_cast<uint4*>(p);
        q[0] = hi; q[1] = lo;
    } else {
        bn_to_be32(p, t.v);
    }
}

template <int OP>
__global__ void __launch_bounds__(256) k_fe_op(uint8_t* __restrict__ out, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x, y, r;
    fe_load_be32_global(x, a + 32 * i);
    if (OP == 0 || OP == 2 || OP == 3) fe_load_be32_global(y, b + 32 * i); else fe_set_zero(y);
    r = x;
#pragma unroll 1
    for (int k = 0; k < iters; k++) {
        if (OP == 0) fe_mul(r, r, y);
        else if (OP == 1) fe_sqr(r, r);
        else if (OP == 2) fe_add(r, r, y);
        else if (OP == 3) fe_sub(r, r, y);
        else if (OP == 4) fe_half(r, r);
        else if (OP == 5) { fe t; fe_inv(t, r); r = t; }
        else if (OP == 6) { fe t; fe_sqrt(t, r); r = t; }
        else if (OP == 7) fe_neg(r, r);
        else if (OP == 8) { fe t; fe_inv_fermat(t, r); r = t; }
        else if (OP == 9) fe_set_int(r, (uint32_t)fe_is_square(r));
        else if (OP == 10) fe_set_int(r, (uint32_t)fe_is_square_sqrt(r));
    }
    fe_store_be32_global(out + 32 * i, r);
}

static cudaError_t launch_fe_op_impl(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, int iters, cudaStream_t s, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    dim3 block(256), grid((unsigned)((n + 255) / 256));
    switch (op) {
        case 0: k_fe_op<0><<<grid, block, 0, s>>>(out, a, b, n, iters); break;
        case 1: k_fe_op<1><<<grid, block, 0, s>>>(out, a, b, n, iters); break;
        case 2: k_fe_op<2><<<grid, block, 0, s>>>(out, a, b, n, iters); break;
        case 3: k_fe_op<3><<<grid, block, 0, s>>>(out, a, b, n, iters); break;
        case 4: k_fe_op<4><<<grid, block, 0, s>>>(out, a, b, n, iters); break;
        case 5: k_fe_op<5><<<grid, block, 0, s>>>(out, a, b, n, iters); break;
        case 6: k_fe_op<6><<<grid, block, 0, s>>>(out, a, b, n, iters); break;
        case 7: k_fe_op<7><<<grid, block, 0, s>>>(out, a, b, n, iters); break;
        case 8: k_fe_op<8><<<grid, block, 0, s>>>(out, a, b, n, iters); break;
        case 9: k_fe_op<9><<<grid, block, 0, s>>>(out, a, b, n, iters); break;
        case 10: k_fe_op<10><<<grid, block, 0, s>>>(out, a, b, n, iters); break;
        default: return cudaErrorInvalidValue;
    }
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

}  // namespace B200_NS
