// chacha_pre.cuh -- the rand-0.3 ChaChaRng block function (chacha.cuh) with the counter-independent part of its first round
// taken out of the per-block work.  Used by the paired-tile share generation (packed_tc2.cuh) and by the mask kernels
// (sharegen.cu) wherever a key serves many blocks whose counters stay below 2^32.
#pragma once
#include <cstddef>
#include <cstdint>

#include "chacha.cuh"

namespace sda {

// (rotates as wide multiplies on the FMA pipe, 1 or 2 of the 4, cost +4 % / +19 % cycles: IMAD.WIDE holds the issue port
// for ~4.8 cycles -- profiles/r01_pipes.md, profiles/r02_k2.md)
#define SDA_QR2(a, b, c, d)                                     \
    a += b; d ^= a; d = __funnelshift_l(d, d, 16);              \
    c += d; b ^= c; b = __funnelshift_l(b, b, 12);              \
    a += b; d ^= a; d = __funnelshift_l(d, d, 8);               \
    c += d; b ^= c; b = __funnelshift_l(b, b, 7);

// Columns 1..3 of the first round do not depend on the block counter while it stays below 2^32 (state words 13..15
// are zero then): they are per-participant constants, computed once per launch by chacha_prepare_kernel and loaded
// instead of being recomputed for every block -- 3 of a block's 8 ROUNDS / 2 quarter rounds.
struct ChaChaPre {
    uint32_t w[12];     // x1,x5,x9,x13, x2,x6,x10,x14, x3,x7,x11,x15 after the first column round, counter high word 0
};

static __global__ void chacha_prepare_kernel(const ChaChaKey *__restrict__ keys, size_t P, ChaChaPre *__restrict__ pre) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const uint32_t c[4] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    const ChaChaKey k = keys[p];
    for (int col = 1; col < 4; col++) {
        uint32_t a = c[col], b = k.w[col], cc = k.w[4 + col], d = 0;
        SDA_QR2(a, b, cc, d)
        pre[p].w[4 * (col - 1) + 0] = a;
        pre[p].w[4 * (col - 1) + 1] = b;
        pre[p].w[4 * (col - 1) + 2] = cc;
        pre[p].w[4 * (col - 1) + 3] = d;
    }
}

// one keystream block (counter b0 < 2^32) from the key and its precomputed first-round columns
template <int ROUNDS>
__device__ __forceinline__ void chacha_block2(const uint32_t (&k)[8], const uint32_t (&pre)[12], uint32_t b0, uint32_t (&o)[16]) {
    const uint32_t c0 = 0x61707865u, c1 = 0x3320646eu, c2 = 0x79622d32u, c3 = 0x6b206574u;
    uint32_t x0 = c0, x4 = k[0], x8 = k[4], x12 = b0;
    uint32_t x1 = pre[0], x5 = pre[1], x9 = pre[2], x13 = pre[3];
    uint32_t x2 = pre[4], x6 = pre[5], x10 = pre[6], x14 = pre[7];
    uint32_t x3 = pre[8], x7 = pre[9], x11 = pre[10], x15 = pre[11];
    SDA_QR2(x0, x4, x8, x12)                   // the one column of round 1 that sees the counter
#ifndef SDA_TC2_UNROLL
#define SDA_TC2_UNROLL 9     // full unrolling of the 9 double rounds: -2.2 % cycles against 3 (no loop-carried register moves)
#endif
#define SDA_TC2_PRAGMA_(x) _Pragma(#x)
#define SDA_TC2_PRAGMA(x) SDA_TC2_PRAGMA_(x)
    SDA_TC2_PRAGMA(unroll SDA_TC2_UNROLL)
    for (int i = 0; i < ROUNDS / 2 - 1; i++) {  // diagonal round, then the next column round
        SDA_QR2(x0, x5, x10, x15)
        SDA_QR2(x1, x6, x11, x12)
        SDA_QR2(x2, x7, x8, x13)
        SDA_QR2(x3, x4, x9, x14)
        SDA_QR2(x0, x4, x8, x12)
        SDA_QR2(x1, x5, x9, x13)
        SDA_QR2(x2, x6, x10, x14)
        SDA_QR2(x3, x7, x11, x15)
    }
    SDA_QR2(x0, x5, x10, x15)
    SDA_QR2(x1, x6, x11, x12)
    SDA_QR2(x2, x7, x8, x13)
    SDA_QR2(x3, x4, x9, x14)
    o[0] = x0 + c0;     o[1] = x1 + c1;     o[2] = x2 + c2;      o[3] = x3 + c3;
    o[4] = x4 + k[0];   o[5] = x5 + k[1];   o[6] = x6 + k[2];    o[7] = x7 + k[3];
    o[8] = x8 + k[4];   o[9] = x9 + k[5];   o[10] = x10 + k[6];  o[11] = x11 + k[7];
    o[12] = x12 + b0;   o[13] = x13;        o[14] = x14;         o[15] = x15;
}


// host twin of chacha_prepare_kernel for a single key (kernels that take their key by value)
inline ChaChaPre chacha_prepare_host(const ChaChaKey &k) {
    auto rotl = [](uint32_t x, int r) { return (x << r) | (x >> (32 - r)); };
    const uint32_t c[4] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    ChaChaPre pre;
    for (int col = 1; col < 4; col++) {
        uint32_t a = c[col], b = k.w[col], cc = k.w[4 + col], d = 0;
        a += b; d ^= a; d = rotl(d, 16);
        cc += d; b ^= cc; b = rotl(b, 12);
        a += b; d ^= a; d = rotl(d, 8);
        cc += d; b ^= cc; b = rotl(b, 7);
        pre.w[4 * (col - 1) + 0] = a;
        pre.w[4 * (col - 1) + 1] = b;
        pre.w[4 * (col - 1) + 2] = cc;
        pre.w[4 * (col - 1) + 3] = d;
    }
    return pre;
}

}  // namespace sda
