// rp_getheader.cuh -- the range-proof header decoder, ONE definition for the device kernels (rangeproof.cuh) and the host API
// (secp256k1_rangeproof_info in api_compat.cu).  Wire format of secp256k1_rangeproof_getheader_impl (modules/rangeproof/rangeproof_impl.h:487-538):
//   byte 0: bit 7 reserved (must be 0), bit 6 "has a non-zero range", bit 5 "has a minimum value", bits 0-4 exponent;
//   [byte: mantissa - 1]  [8 bytes big-endian: minimum value]
// The outputs are written in the reference's order, so that a caller sees the same partial results when a later check fails.
#pragma once
#include <stddef.h>
#include <stdint.h>
#ifndef HD
#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif
#endif

HD int rp_getheader(size_t& offset, int& exp, int& mantissa, uint64_t& scale, uint64_t& min_value, uint64_t& max_value, const uint8_t* proof, size_t plen) {
    offset = 0;
    if (plen < 65 || (proof[0] & 128) != 0) return 0;
    const int has_nz_range = proof[0] & 64, has_min = proof[0] & 32;
    exp = -1;
    mantissa = 0;
    if (has_nz_range) {
        exp = proof[0] & 31;
        offset += 1;
        if (exp > 18) return 0;
        mantissa = proof[offset] + 1;
        if (mantissa > 64) return 0;
        max_value = UINT64_MAX >> (64 - mantissa);
    } else {
        max_value = 0;
    }
    offset += 1;
    scale = 1;
    for (int i = 0; i < exp; i++) {
        if (max_value > UINT64_MAX / 10) return 0;
        max_value *= 10;
        scale *= 10;
    }
    min_value = 0;
    if (has_min) {
        if (plen - offset < 8) return 0;
        for (int i = 0; i < 8; i++) min_value = (min_value << 8) | proof[offset + i];
        offset += 8;
    }
    if (max_value > UINT64_MAX - min_value) return 0;
    max_value += min_value;
    return 1;
}
