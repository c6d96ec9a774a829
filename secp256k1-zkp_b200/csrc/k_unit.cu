// k_unit.cu -- unit kernels of the scalar and group primitives, one item per thread (or per group of 4 lanes for the ge_g4.cuh forms).
// Not on any hot path: they exist so that the PTX limb code of sc_mul / sc_inverse / sc_split_lambda / gej_double / gej_add /
// gej_add_ge -- and the lane-parallel gej_*_g4 -- can be compared with the reference's own functions (scalar_4x64_impl.h, scalar_impl.h:142,
// group_impl.h:503,:534,:598) input by input, including the degenerate branches (a = b, a = -b, infinity on either side).
#include "ge_g4.cuh"
#include "ge_w10.cuh"
#include "sc.cuh"
#include "launch.h"

namespace b200 {

// group-op operand: 96 bytes  x || y (affine, big-endian; all-zero = infinity) || s (big-endian, reduced mod p; 0 is taken as 1)
// the kernel works on the Jacobian point (x s^2, y s^3, s), i.e. the same group element with a non-trivial z.
__device__ void unit_load_gej(gej& r, const uint8_t* in96) {
    uint32_t z = 0;
    for (int k = 0; k < 64; k++) z |= in96[k];
    if (z == 0) { gej_set_infinity(r); return; }
    fe x, y, s, s2, s3;
    fe_set_b32(x, in96); fe_set_b32(y, in96 + 32); fe_set_b32(s, in96 + 64);
    if (fe_is_zero(s)) fe_set_int(s, 1);
    fe_sqr(s2, s); fe_mul(s3, s2, s);
    fe_mul(r.x, x, s2); fe_mul(r.y, y, s3); r.z = s; r.inf = 0;
}
__device__ void unit_store_gej(uint8_t* out64, const gej& r) {
    if (r.inf) { for (int k = 0; k < 64; k++) out64[k] = 0; return; }
    ge a; ge_set_gej(a, r);
    fe_get_b32(out64, a.x); fe_get_b32(out64 + 32, a.y);
}

__global__ void __launch_bounds__(128) k_unit_op(int op, uint8_t* __restrict__ out, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool g4 = op >= 24;                     // lane-parallel forms: 4 lanes per item (24-26), a whole warp per item (27-29)
    const int lanes_log2 = op >= 27 ? 5 : 2;
    size_t i = g4 ? (t >> lanes_log2) : t;
    const bool valid = i < n;
    if (!valid) { if (!g4) return; i = n - 1; }   // these forms need whole warps: surplus groups redo the last item and store nothing
    if (op < 8) {
        sc x, y, r;
        sc_set_b32(x, a + 32 * i);
        if (b) sc_set_b32(y, b + 32 * i); else sc_set_zero(y);
        switch (op) {
        case 0: sc_mul(r, x, y); sc_get_b32(out + 32 * i, r); break;
        case 1: sc_add(r, x, y); sc_get_b32(out + 32 * i, r); break;
        case 2: sc_negate(r, x); sc_get_b32(out + 32 * i, r); break;
        case 3: sc_inverse(r, x); sc_get_b32(out + 32 * i, r); break;
        case 4: { sc r1, r2; sc_split_lambda(r1, r2, x); sc_get_b32(out + 64 * i, r1); sc_get_b32(out + 64 * i + 32, r2); break; }
        default: sc_get_b32(out + 32 * i, x); break;          // 5: a mod n (sc_set_b32 + reduce)
        }
        return;
    }
    gej A, B, R;
    unit_load_gej(A, a + 96 * i);
    if (b) unit_load_gej(B, b + 96 * i); else gej_set_infinity(B);
    ge Ba;                                    // affine form of b for the mixed additions (s ignored)
    {
        uint32_t z = 0;
        if (b) for (int k = 0; k < 64; k++) z |= b[96 * i + k];
        if (z == 0) ge_set_infinity(Ba); else { fe_set_b32(Ba.x, b + 96 * i); fe_set_b32(Ba.y, b + 96 * i + 32); Ba.inf = 0; }
    }
    switch (op) {
    case 8: gej_double(R, A); break;
    case 9: gej_add(R, A, B); break;
    case 10: gej_add_ge(R, A, Ba); break;
    case 24: gej_double_g4(R, A); break;
    case 25: gej_add_g4(R, A, B); break;
    case 26: gej_add_ge_g4(R, A, Ba); break;
    case 27: gej_double_n_w10(R, A, 1); break;
    case 28: gej_double_n_w10(R, A, 16); break;
    default: gej_add_w10(R, A, B); break;
    }
    if (valid && (!g4 || (threadIdx.x & ((1u << lanes_log2) - 1)) == 0)) unit_store_gej(out + 64 * i, R);
}

cudaError_t launch_unit_op(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, cudaStream_t s, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    const bool ok = (op >= 0 && op <= 5) || (op >= 8 && op <= 10) || (op >= 24 && op <= 29);
    if (!ok) return cudaErrorInvalidValue;
    size_t threads = op >= 27 ? 32 * n : op >= 24 ? 4 * n : n;
    k_unit_op<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(op, out, a, b, n);
    if (launches) __atomic_fetch_add(launches, (uint64_t)1, __ATOMIC_RELAXED);
    return cudaGetLastError();
}

}  // namespace b200
