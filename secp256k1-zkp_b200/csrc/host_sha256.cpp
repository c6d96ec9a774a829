// host_sha256.cpp -- host-side walker of the half-aggregation randomizer chain (halfagg.cuh, halfagg_z_chain): the chain is ONE
// running SHA-256 over all signatures (schnorrsig_halfagg/main_impl.h:144-155), serial by construction, so for host-buffer calls it
// runs on the host CPU while the uploads are in flight.  Uses the x86 SHA extensions when the CPU has them (runtime check), else the
// portable compression function of sha256.cuh.  Plain C++ (g++), no CUDA.
#include "halfagg.cuh"
#include <cstring>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define B200_HAVE_SHANI 1
#else
#define B200_HAVE_SHANI 0
#endif

namespace {

typedef void (*compress_fn)(uint32_t* state, const uint8_t* block64);

void compress_portable(uint32_t* s, const uint8_t* b) {
    uint32_t w[16];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)b[4 * i] << 24) | ((uint32_t)b[4 * i + 1] << 16) | ((uint32_t)b[4 * i + 2] << 8) | (uint32_t)b[4 * i + 3];
    b200::sha256_compress(s, w);
}

#if B200_HAVE_SHANI
alignas(16) static const uint32_t K256[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

// One block with SHA256RNDS2 / SHA256MSG1 / SHA256MSG2.  The two state registers hold (A,B,E,F) and (C,D,G,H) as the
// instructions expect; four rounds per step, the message schedule kept in a ring of four 4-word registers.
__attribute__((target("sha,sse4.1,ssse3"))) void compress_shani(uint32_t* s, const uint8_t* b) {
    const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i t = _mm_loadu_si128((const __m128i*)s);            // D C B A (high..low = s[3] s[2] s[1] s[0])
    __m128i st1 = _mm_loadu_si128((const __m128i*)(s + 4));     // H G F E
    t = _mm_shuffle_epi32(t, 0xB1);                             // C D A B
    st1 = _mm_shuffle_epi32(st1, 0x1B);                         // E F G H
    __m128i st0 = _mm_alignr_epi8(t, st1, 8);                   // A B E F
    st1 = _mm_blend_epi16(st1, t, 0xF0);                        // C D G H
    const __m128i save0 = st0, save1 = st1;
    __m128i m[4];
    for (int i = 0; i < 4; i++) m[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(b + 16 * i)), bswap);
    for (int r = 0; r < 16; r++) {                              // r-th group of four rounds
        __m128i w = m[r & 3];
        __m128i wk = _mm_add_epi32(w, _mm_load_si128((const __m128i*)(K256 + 4 * r)));
        st1 = _mm_sha256rnds2_epu32(st1, st0, wk);
        st0 = _mm_sha256rnds2_epu32(st0, st1, _mm_shuffle_epi32(wk, 0x0E));
        if (r < 12) {                                           // schedule words 16+4r .. 19+4r into the slot being vacated
            __m128i x = _mm_sha256msg1_epu32(m[r & 3], m[(r + 1) & 3]);
            x = _mm_add_epi32(x, _mm_alignr_epi8(m[(r + 3) & 3], m[(r + 2) & 3], 4));
            m[r & 3] = _mm_sha256msg2_epu32(x, m[(r + 3) & 3]);
        }
    }
    st0 = _mm_add_epi32(st0, save0);
    st1 = _mm_add_epi32(st1, save1);
    t = _mm_shuffle_epi32(st0, 0x1B);                           // F E B A
    st1 = _mm_shuffle_epi32(st1, 0xB1);                         // D C H G
    st0 = _mm_blend_epi16(t, st1, 0xF0);                        // D C B A
    st1 = _mm_alignr_epi8(st1, t, 8);                           // H G F E
    _mm_storeu_si128((__m128i*)s, st0);
    _mm_storeu_si128((__m128i*)(s + 4), st1);
}
#endif

compress_fn pick_compress(int force_portable) {
#if B200_HAVE_SHANI
    if (!force_portable && __builtin_cpu_supports("sha") && __builtin_cpu_supports("sse4.1") && __builtin_cpu_supports("ssse3")) return compress_shani;
#endif
    (void)force_portable;
    return compress_portable;
}

void final_block(uint8_t* blk, size_t fill, uint64_t total_bytes) {   // blk[0..fill) holds data; pads in place (fill <= 55)
    blk[fill++] = 0x80;
    while (fill < 56) blk[fill++] = 0;
    uint64_t bits = total_bytes * 8;
    for (int i = 0; i < 8; i++) blk[56 + i] = (uint8_t)(bits >> (56 - 8 * i));
}
void state_to_be(uint8_t* out, const uint32_t* s) {
    for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(s[i] >> 24); out[4 * i + 1] = (uint8_t)(s[i] >> 16); out[4 * i + 2] = (uint8_t)(s[i] >> 8); out[4 * i + 3] = (uint8_t)s[i]; }
}

}  // namespace

// Same output as b200::halfagg_z_chain.  Triples are 96 bytes and the prefix is 64, so the stream alternates between "block boundary
// after r||pk" and "32 bytes pending": 2.5 compressions per signature including the finalisation of the running copy.
// force_portable != 0 skips the SHA-extension path (tests compare the two).  Returns 1 if the SHA extensions were used.
extern "C" int b200_host_halfagg_z_chain(uint8_t* z32, const uint32_t* mid, const uint8_t* aggsig, const uint8_t* pk32, const uint8_t* msgs32,
                                         size_t n, int force_portable) {
    compress_fn comp = pick_compress(force_portable);
    uint32_t st[8], c[8];
    for (int k = 0; k < 8; k++) st[k] = mid[k];
    uint8_t blk[64];
    for (size_t i = 0; i < n; i++) {
        const uint8_t *r = aggsig + 32 * i, *pk = pk32 + 32 * i, *m = msgs32 + 32 * i;
        uint64_t total = 64 + 96 * (uint64_t)(i + 1);
        size_t fill;
        if ((i & 1) == 0) {                       // buffer empty: r||pk is a block, m stays pending
            memcpy(blk, r, 32); memcpy(blk + 32, pk, 32);
            comp(st, blk);
            memcpy(blk, m, 32);
            fill = 32;
        } else {                                  // pending m_{i-1}: m_{i-1}||r, then pk||m
            memcpy(blk, msgs32 + 32 * (i - 1), 32); memcpy(blk + 32, r, 32);
            comp(st, blk);
            memcpy(blk, pk, 32); memcpy(blk + 32, m, 32);
            comp(st, blk);
            fill = 0;
        }
        if (i == 0) {
            for (int k = 0; k < 32; k++) z32[k] = (k == 31);
        } else {
            for (int k = 0; k < 8; k++) c[k] = st[k];
            final_block(blk, fill, total);
            comp(c, blk);
            state_to_be(z32 + 32 * i, c);
        }
    }
#if B200_HAVE_SHANI
    return comp == compress_shani;
#else
    return 0;
#endif
}
