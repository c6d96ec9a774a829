// bn.cuh -- 256-bit big-number primitives for sm_100a.
//
// Representation: 8 x uint32 little-endian limbs == the reference's packed
// `secp256k1_fe_storage` / `secp256k1_scalar` 4x64 form (field_5x52.h:45-47,
// scalar_4x64.h:13-15) viewed as 32-bit words.  The B200 integer datapath is
// 32-bit: one 64x64->128 product of the reference's 5x52 schedule costs four
// IMAD.WIDE.U32, so a 5x52 fe_mul is 124 MACs while the saturated
// 8x32 schedule below is 64 + 8 MACs.  Loads/stores and results are the
// reference's 4x64 forms; only the in-register limb schedule differs.
//
// Device path: inline PTX mad.lo.cc/madc.hi.cc chains which ptxas fuses into
// IMAD.WIDE.U32(.X) (checked with cuobjdump: 64 products -> 64 IMAD.WIDE).
// The "even/odd column" split keeps every chain free of 32-bit-misaligned
// carries.  A portable C path (uint64_t accumulators) is used when compiled
// for the host (tests/hostemu) or with -DB200_PORTABLE_BN (on-device A/B check).
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDNI static __host__ __device__ __noinline__
#else
#define HD inline
#define HDNI static
#endif

#ifndef B200_NS
#define B200_NS b200
#endif

#if defined(__CUDA_ARCH__) && !defined(B200_PORTABLE_BN)
#define BN_PTX 1
#else
#define BN_PTX 0
#endif

namespace B200_NS {

#if BN_PTX
// acc[0..8] += (a0 + a1*2^64 + a2*2^128 + a3*2^192) * b   (a_i, b 32-bit; four aligned 64-bit products)
// CO = false: the caller knows acc[7] is still zero (a fresh limb), so the top word  hi(a3*b) + carry <= 2^32 - 1  cannot overflow and the
// carry-out into acc[8] is omitted (one instruction -- a SEL or an IMAD.X on the multiply pipe -- and one zero-initialised register less).
template <bool CO = true>
__device__ __forceinline__ void bn_cmad4(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    if (CO) {
        asm("mad.lo.cc.u32  %0, %9,  %13, %0;\n\t"
            "madc.hi.cc.u32 %1, %9,  %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32       %8, %8, 0;\n\t"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
    } else {
        asm("mad.lo.cc.u32  %0, %8,  %12, %0;\n\t"
            "madc.hi.cc.u32 %1, %8,  %12, %1;\n\t"
            "madc.lo.cc.u32 %2, %9,  %12, %2;\n\t"
            "madc.hi.cc.u32 %3, %9,  %12, %3;\n\t"
            "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
            "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
            "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
            "madc.hi.u32    %7, %11, %12, %7;\n\t"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
    }
}
template <bool CO = true>
__device__ __forceinline__ void bn_cmad3(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b) {
    if (CO) {
        asm("mad.lo.cc.u32  %0, %7, %10, %0;\n\t"
            "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
            "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
            "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
            "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
            "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
            "addc.u32       %6, %6, 0;\n\t"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6])
            : "r"(a0), "r"(a1), "r"(a2), "r"(b));
    } else {
        asm("mad.lo.cc.u32  %0, %6, %9, %0;\n\t"
            "madc.hi.cc.u32 %1, %6, %9, %1;\n\t"
            "madc.lo.cc.u32 %2, %7, %9, %2;\n\t"
            "madc.hi.cc.u32 %3, %7, %9, %3;\n\t"
            "madc.lo.cc.u32 %4, %8, %9, %4;\n\t"
            "madc.hi.u32    %5, %8, %9, %5;\n\t"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5])
            : "r"(a0), "r"(a1), "r"(a2), "r"(b));
    }
}
template <bool CO = true>
__device__ __forceinline__ void bn_cmad2(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t b) {
    if (CO) {
        asm("mad.lo.cc.u32  %0, %5, %7, %0;\n\t"
            "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
            "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
            "addc.u32       %4, %4, 0;\n\t"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4])
            : "r"(a0), "r"(a1), "r"(b));
    } else {
        asm("mad.lo.cc.u32  %0, %4, %6, %0;\n\t"
            "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
            "madc.lo.cc.u32 %2, %5, %6, %2;\n\t"
            "madc.hi.u32    %3, %5, %6, %3;\n\t"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3])
            : "r"(a0), "r"(a1), "r"(b));
    }
}
template <bool CO = true>
__device__ __forceinline__ void bn_cmad1(uint32_t* acc, uint32_t a0, uint32_t b) {
    if (CO) {
        asm("mad.lo.cc.u32  %0, %3, %4, %0;\n\t"
            "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
            "addc.u32       %2, %2, 0;\n\t"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2])
            : "r"(a0), "r"(b));
    } else {
        asm("mad.lo.cc.u32  %0, %2, %3, %0;\n\t"
            "madc.hi.u32    %1, %2, %3, %1;\n\t"
            : "+r"(acc[0]), "+r"(acc[1])
            : "r"(a0), "r"(b));
    }
}
#endif

// r[0..15] = a[0..7] * b[0..7]
HD void bn_mul_wide(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if BN_PTX && defined(BN_MUL_SPLIT4)
    // four accumulator sets (rows 0-3 and rows 4-7 apart): the serial IMAD.WIDE.X dependency per accumulator halves
    // (16 instead of 32 links) at the price of ~20 more adds on the (idle) ALU pipe
    uint32_t e0[13], o0[13], e1[13], o1[13];
#pragma unroll
    for (int i = 0; i < 13; i++) { e0[i] = 0; o0[i] = 0; e1[i] = 0; o1[i] = 0; }
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
        bn_cmad4(e0 + i, a[0], a[2], a[4], a[6], b[i]);
        bn_cmad4(e1 + i, a[0], a[2], a[4], a[6], b[4 + i]);
        bn_cmad4(o0 + i, a[1], a[3], a[5], a[7], b[i]);
        bn_cmad4(o1 + i, a[1], a[3], a[5], a[7], b[4 + i]);
        bn_cmad4(o0 + i, a[0], a[2], a[4], a[6], b[i + 1]);
        bn_cmad4(o1 + i, a[0], a[2], a[4], a[6], b[4 + i + 1]);
        bn_cmad4(e0 + i + 2, a[1], a[3], a[5], a[7], b[i + 1]);
        bn_cmad4(e1 + i + 2, a[1], a[3], a[5], a[7], b[4 + i + 1]);
    }
    // lo = e0 + (o0 << 32)  (12 limbs: a * b[0..3] < 2^384), hi likewise
    uint32_t lo[12], hi[12], c;
    lo[0] = e0[0]; hi[0] = e1[0];
    asm("add.cc.u32  %0, %12, %23;\n\taddc.cc.u32 %1, %13, %24;\n\taddc.cc.u32 %2, %14, %25;\n\taddc.cc.u32 %3, %15, %26;\n\t"
        "addc.cc.u32 %4, %16, %27;\n\taddc.cc.u32 %5, %17, %28;\n\taddc.cc.u32 %6, %18, %29;\n\taddc.cc.u32 %7, %19, %30;\n\t"
        "addc.cc.u32 %8, %20, %31;\n\taddc.cc.u32 %9, %21, %32;\n\taddc.u32 %10, %22, %33;\n\tmov.u32 %11, 0;\n\t"
        : "=r"(lo[1]), "=r"(lo[2]), "=r"(lo[3]), "=r"(lo[4]), "=r"(lo[5]), "=r"(lo[6]), "=r"(lo[7]), "=r"(lo[8]), "=r"(lo[9]), "=r"(lo[10]), "=r"(lo[11]), "=r"(c)
        : "r"(e0[1]), "r"(e0[2]), "r"(e0[3]), "r"(e0[4]), "r"(e0[5]), "r"(e0[6]), "r"(e0[7]), "r"(e0[8]), "r"(e0[9]), "r"(e0[10]), "r"(e0[11]),
          "r"(o0[0]), "r"(o0[1]), "r"(o0[2]), "r"(o0[3]), "r"(o0[4]), "r"(o0[5]), "r"(o0[6]), "r"(o0[7]), "r"(o0[8]), "r"(o0[9]), "r"(o0[10]));
    asm("add.cc.u32  %0, %12, %23;\n\taddc.cc.u32 %1, %13, %24;\n\taddc.cc.u32 %2, %14, %25;\n\taddc.cc.u32 %3, %15, %26;\n\t"
        "addc.cc.u32 %4, %16, %27;\n\taddc.cc.u32 %5, %17, %28;\n\taddc.cc.u32 %6, %18, %29;\n\taddc.cc.u32 %7, %19, %30;\n\t"
        "addc.cc.u32 %8, %20, %31;\n\taddc.cc.u32 %9, %21, %32;\n\taddc.u32 %10, %22, %33;\n\tmov.u32 %11, 0;\n\t"
        : "=r"(hi[1]), "=r"(hi[2]), "=r"(hi[3]), "=r"(hi[4]), "=r"(hi[5]), "=r"(hi[6]), "=r"(hi[7]), "=r"(hi[8]), "=r"(hi[9]), "=r"(hi[10]), "=r"(hi[11]), "=r"(c)
        : "r"(e1[1]), "r"(e1[2]), "r"(e1[3]), "r"(e1[4]), "r"(e1[5]), "r"(e1[6]), "r"(e1[7]), "r"(e1[8]), "r"(e1[9]), "r"(e1[10]), "r"(e1[11]),
          "r"(o1[0]), "r"(o1[1]), "r"(o1[2]), "r"(o1[3]), "r"(o1[4]), "r"(o1[5]), "r"(o1[6]), "r"(o1[7]), "r"(o1[8]), "r"(o1[9]), "r"(o1[10]));
    // r = lo + (hi << 128)
    r[0] = lo[0]; r[1] = lo[1]; r[2] = lo[2]; r[3] = lo[3];
    asm("add.cc.u32  %0, %12, %20;\n\taddc.cc.u32 %1, %13, %21;\n\taddc.cc.u32 %2, %14, %22;\n\taddc.cc.u32 %3, %15, %23;\n\t"
        "addc.cc.u32 %4, %16, %24;\n\taddc.cc.u32 %5, %17, %25;\n\taddc.cc.u32 %6, %18, %26;\n\taddc.cc.u32 %7, %19, %27;\n\t"
        "addc.cc.u32 %8, %28, 0;\n\taddc.cc.u32 %9, %29, 0;\n\taddc.cc.u32 %10, %30, 0;\n\taddc.u32 %11, %31, 0;\n\t"
        : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]), "r"(lo[8]), "r"(lo[9]), "r"(lo[10]), "r"(lo[11]),
          "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]),
          "r"(hi[8]), "r"(hi[9]), "r"(hi[10]), "r"(hi[11]));
    (void)c;
#elif BN_PTX
    uint32_t e[17], o[17];
#pragma unroll
    for (int i = 0; i < 17; i++) { e[i] = 0; o[i] = 0; }
    // e[k] holds weight 2^(32k) for products whose column is even, o[k] weight 2^(32(k+1)) for odd columns
    // Carry-outs that cannot happen are not emitted (r2): a chain whose top target limb has never been written ends in  hi + carry  of a
    // fresh zero, which fits 32 bits.  That is the case for the very first chain on e, for every first chain of a pair on o (o[i+7] is fresh:
    // the previous pair reached o[i+5] and put its carries into o[i+6]) and for every second chain on e (e[i+9] is fresh): 9 of the 16 chains.
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        if (i == 0) bn_cmad4<false>(e + i, a[0], a[2], a[4], a[6], b[i]);   // columns i+0,2,4,6   (even)
        else bn_cmad4<true>(e + i, a[0], a[2], a[4], a[6], b[i]);
        bn_cmad4<false>(o + i, a[1], a[3], a[5], a[7], b[i]);               // columns i+1,3,5,7   (odd)
        bn_cmad4<true>(o + i, a[0], a[2], a[4], a[6], b[i + 1]);            // columns i+1+0,2,..  (odd)
        bn_cmad4<false>(e + i + 2, a[1], a[3], a[5], a[7], b[i + 1]);       // columns i+2,4,..    (even)
    }
    r[0] = e[0];
    uint32_t cmid, cdummy;
    asm("add.cc.u32  %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;\n\t"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(cmid)
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]),
          "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]));
    // the carry is handed over in a register (cmid) and re-materialised with add.cc cmid + 0xffffffff,
    // so no statement depends on the CC flag surviving between two asm blocks.
    asm("add.cc.u32  %7, %22, 0xffffffff;\n\t"
        "addc.cc.u32 %0, %8,  %15;\n\t"
        "addc.cc.u32 %1, %9,  %16;\n\t"
        "addc.cc.u32 %2, %10, %17;\n\t"
        "addc.cc.u32 %3, %11, %18;\n\t"
        "addc.cc.u32 %4, %12, %19;\n\t"
        "addc.cc.u32 %5, %13, %20;\n\t"
        "addc.u32    %6, %14, %21;\n\t"
        : "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(cdummy)
        : "r"(e[9]), "r"(e[10]), "r"(e[11]), "r"(e[12]), "r"(e[13]), "r"(e[14]), "r"(e[15]),
          "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(cmid));
    (void)cdummy;
#else
    // column sums of low/high halves kept apart so nothing overflows 64 bits
    uint32_t out[16];
    uint64_t carry = 0;
    for (int k = 0; k < 15; k++) {
        uint64_t lo = carry & 0xffffffffu, hi = carry >> 32;
        for (int i = 0; i < 8; i++) {
            int j = k - i;
            if (j < 0 || j > 7) continue;
            uint64_t p = (uint64_t)a[i] * b[j];
            lo += p & 0xffffffffu;
            hi += p >> 32;
        }
        out[k] = (uint32_t)lo;
        carry = hi + (lo >> 32);
    }
    out[15] = (uint32_t)carry;
    for (int i = 0; i < 16; i++) r[i] = out[i];
#endif
}

// r[0..15] = a[0..7]^2
HD void bn_sqr_wide(uint32_t* r, const uint32_t* a) {
#if BN_PTX
    // off-diagonal products a_i*a_j (i<j), 28 instead of 56, in the same even/odd column split
    uint32_t e[17], o[17];
#pragma unroll
    for (int i = 0; i < 17; i++) { e[i] = 0; o[i] = 0; }
    // <false>: the chain's top target limb has never been written, so no carry can leave it (see bn_mul_wide)
    bn_cmad3<false>(e + 2, a[2], a[4], a[6], a[0]);   bn_cmad4<false>(o + 0, a[1], a[3], a[5], a[7], a[0]);
    bn_cmad3<false>(e + 4, a[3], a[5], a[7], a[1]);   bn_cmad3<true>(o + 2, a[2], a[4], a[6], a[1]);
    bn_cmad2<true>(e + 6, a[4], a[6], a[2]);          bn_cmad3<false>(o + 4, a[3], a[5], a[7], a[2]);
    bn_cmad2<false>(e + 8, a[5], a[7], a[3]);         bn_cmad2<true>(o + 6, a[4], a[6], a[3]);
    bn_cmad1<true>(e + 10, a[6], a[4]);               bn_cmad2<false>(o + 8, a[5], a[7], a[4]);
    bn_cmad1<false>(e + 12, a[7], a[5]);              bn_cmad1<true>(o + 10, a[6], a[5]);
    bn_cmad1<false>(o + 12, a[7], a[6]);
    // s = e + (o << 32)   (s[0] = s[1]... e[0], e[1] are zero by construction)
    uint32_t s[16], cmid, cdummy;
    s[0] = 0;
    asm("add.cc.u32  %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;\n\t"
        : "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(s[8]), "=r"(cmid)
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]),
          "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]));
    asm("add.cc.u32  %7, %22, 0xffffffff;\n\t"
        "addc.cc.u32 %0, %8,  %15;\n\t"
        "addc.cc.u32 %1, %9,  %16;\n\t"
        "addc.cc.u32 %2, %10, %17;\n\t"
        "addc.cc.u32 %3, %11, %18;\n\t"
        "addc.cc.u32 %4, %12, %19;\n\t"
        "addc.cc.u32 %5, %13, %20;\n\t"
        "addc.u32    %6, %14, %21;\n\t"
        : "=r"(s[9]), "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15]), "=r"(cdummy)
        : "r"(e[9]), "r"(e[10]), "r"(e[11]), "r"(e[12]), "r"(e[13]), "r"(e[14]), "r"(e[15]),
          "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(cmid));
    // d = 2*s (funnel shifts, no carry chain)
    uint32_t d[16];
    d[0] = s[0] << 1;
#pragma unroll
    for (int k = 1; k < 16; k++) d[k] = __funnelshift_l(s[k - 1], s[k], 1);
    // diagonal squares
    uint32_t q[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        asm("mul.lo.u32 %0, %2, %2;\n\tmul.hi.u32 %1, %2, %2;\n\t" : "=r"(q[2 * i]), "=r"(q[2 * i + 1]) : "r"(a[i]));
    }
    // r = d + q
    asm("add.cc.u32  %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;\n\t"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(cmid)
        : "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]),
          "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]));
    asm("add.cc.u32  %8, %25, 0xffffffff;\n\t"
        "addc.cc.u32 %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.u32    %7, %16, %24;\n\t"
        : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(cdummy)
        : "r"(d[8]), "r"(d[9]), "r"(d[10]), "r"(d[11]), "r"(d[12]), "r"(d[13]), "r"(d[14]), "r"(d[15]),
          "r"(q[8]), "r"(q[9]), "r"(q[10]), "r"(q[11]), "r"(q[12]), "r"(q[13]), "r"(q[14]), "r"(q[15]), "r"(cmid));
    (void)cdummy;
#else
    bn_mul_wide(r, a, a);
#endif
}

// r = a + b, returns carry (0/1)
HD uint32_t bn_add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if BN_PTX
    uint32_t c;
    asm("add.cc.u32  %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;\n\t"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
#endif
}

// r = a - b, returns borrow (0/1)
HD uint32_t bn_sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if BN_PTX
    uint32_t c;
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;\n\t"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c & 1u;   // subc.u32 0-0-borrow = 0xffffffff when borrow
#else
    int64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (int64_t)a[i] - (int64_t)b[i];
        r[i] = (uint32_t)c;
        c >>= 32;  // arithmetic shift: 0 or -1
    }
    return (uint32_t)(c & 1);
#endif
}

// compare a >= b (both 8 limbs)
HD int bn_gte8(const uint32_t* a, const uint32_t* b) {
    uint32_t t[8];
    return bn_sub8(t, a, b) == 0;
}

HD int bn_is_zero8(const uint32_t* a) {
    uint32_t x = a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7];
    return x == 0;
}

HD int bn_eq8(const uint32_t* a, const uint32_t* b) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= a[i] ^ b[i];
    return x == 0;
}

// 32 big-endian bytes -> 8 LE limbs
HD void bn_from_be32(uint32_t* r, const uint8_t* b) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint8_t* p = b + 4 * (7 - i);
        r[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
    }
}
HD void bn_to_be32(uint8_t* b, const uint32_t* a) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint8_t* p = b + 4 * (7 - i);
        p[0] = (uint8_t)(a[i] >> 24); p[1] = (uint8_t)(a[i] >> 16); p[2] = (uint8_t)(a[i] >> 8); p[3] = (uint8_t)a[i];
    }
}

}  // namespace B200_NS
