// sha256.cuh -- FIPS 180-4 SHA-256 usable from host and device code
// (reference: hash_impl.h:51-194 compression/write/finalize, :196-205 tagged init).
#pragma once
#include "bn.cuh"

namespace B200_NS {

struct sha256 {
    uint32_t s[8];
    uint8_t buf[64];
    uint64_t bytes;
};

HD uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

HDNI void sha256_compress(uint32_t* s, const uint32_t* win) {   // ~2500 instructions: never inline
    const uint32_t K[64] = {
        0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
        0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
        0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
        0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
        0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
        0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
        0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
        0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = win[i];
    uint32_t a = s[0], b = s[1], c = s[2], d = s[3], e = s[4], f = s[5], g = s[6], h = s[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t wi;
        if (i < 16) {
            wi = w[i];
        } else {
            uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
            uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            wi = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
            w[i & 15] = wi;
        }
        uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = h + S1 + ch + K[i] + wi;
        uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    s[0] += a; s[1] += b; s[2] += c; s[3] += d; s[4] += e; s[5] += f; s[6] += g; s[7] += h;
}

HD void sha256_init(sha256& h) {
    h.s[0] = 0x6a09e667u; h.s[1] = 0xbb67ae85u; h.s[2] = 0x3c6ef372u; h.s[3] = 0xa54ff53au;
    h.s[4] = 0x510e527fu; h.s[5] = 0x9b05688cu; h.s[6] = 0x1f83d9abu; h.s[7] = 0x5be0cd19u;
    h.bytes = 0;
}
HD void sha256_block_from_buf(sha256& h) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint8_t* p = h.buf + 4 * i;
        w[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
    }
    sha256_compress(h.s, w);
}
HD void sha256_write(sha256& h, const uint8_t* data, size_t len) {
    size_t fill = (size_t)(h.bytes & 63);
    h.bytes += len;
    for (size_t i = 0; i < len; i++) {
        h.buf[fill++] = data[i];
        if (fill == 64) { sha256_block_from_buf(h); fill = 0; }
    }
}
HD void sha256_finalize(sha256& h, uint8_t* out32) {
    uint64_t bits = h.bytes * 8;
    size_t fill = (size_t)(h.bytes & 63);
    h.buf[fill++] = 0x80;
    if (fill > 56) {
        while (fill < 64) h.buf[fill++] = 0;
        sha256_block_from_buf(h);
        fill = 0;
    }
    while (fill < 56) h.buf[fill++] = 0;
    for (int i = 0; i < 8; i++) h.buf[56 + i] = (uint8_t)(bits >> (56 - 8 * i));
    sha256_block_from_buf(h);
    for (int i = 0; i < 8; i++) {
        out32[4 * i] = (uint8_t)(h.s[i] >> 24); out32[4 * i + 1] = (uint8_t)(h.s[i] >> 16);
        out32[4 * i + 2] = (uint8_t)(h.s[i] >> 8); out32[4 * i + 3] = (uint8_t)h.s[i];
    }
}
// SHA256(SHA256(tag) || SHA256(tag) || ...) prefix state (hash_impl.h:196-205)
HD void sha256_init_tagged(sha256& h, const uint8_t* tag, size_t taglen) {
    uint8_t th[32];
    sha256_init(h);
    sha256_write(h, tag, taglen);
    sha256_finalize(h, th);
    sha256_init(h);
    sha256_write(h, th, 32);
    sha256_write(h, th, 32);
}

}  // namespace B200_NS
