// chacha.cuh -- the rand-0.3 `ChaChaRng` keystream: RFC-7539 block function
// with a 128-bit little-endian block counter in words 12..15 and no nonce, `ROUNDS` rounds.
// u64 draw e (no rejections) = (word[2e] << 32) | word[2e+1]  -> block e/8, word pair e%8.
// One thread computes one whole block: 8 draws.
#pragma once
#include <cstdint>

namespace sda {

struct ChaChaKey {
    uint32_t w[8];
};

inline ChaChaKey key_from_seed_bytes(const uint8_t seed[32]) {
    ChaChaKey k;
    for (int i = 0; i < 8; i++)
        k.w[i] = (uint32_t)seed[4 * i] | (uint32_t)seed[4 * i + 1] << 8 | (uint32_t)seed[4 * i + 2] << 16 |
                 (uint32_t)seed[4 * i + 3] << 24;
    return k;
}
inline ChaChaKey key_from_words(const uint32_t *words, size_t n) {
    ChaChaKey k{};
    for (size_t i = 0; i < n && i < 8; i++) k.w[i] = words[i];
    return k;
}

#define SDA_QR(a, b, c, d)                                      \
    a += b; d ^= a; d = __funnelshift_l(d, d, 16);              \
    c += d; b ^= c; b = __funnelshift_l(b, b, 12);              \
    a += b; d ^= a; d = __funnelshift_l(d, d, 8);               \
    c += d; b ^= c; b = __funnelshift_l(b, b, 7);

#if defined(__CUDACC__)
template <int ROUNDS>
__device__ __forceinline__ void chacha_block(const ChaChaKey &key, uint64_t block, uint32_t out[16]) {
    const uint32_t c0 = 0x61707865u, c1 = 0x3320646eu, c2 = 0x79622d32u, c3 = 0x6b206574u;
    const uint32_t b0 = (uint32_t)block, b1 = (uint32_t)(block >> 32);
    uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3;
    uint32_t x4 = key.w[0], x5 = key.w[1], x6 = key.w[2], x7 = key.w[3];
    uint32_t x8 = key.w[4], x9 = key.w[5], x10 = key.w[6], x11 = key.w[7];
    uint32_t x12 = b0, x13 = b1, x14 = 0, x15 = 0;
#pragma unroll
    for (int i = 0; i < ROUNDS / 2; i++) {
        SDA_QR(x0, x4, x8, x12)
        SDA_QR(x1, x5, x9, x13)
        SDA_QR(x2, x6, x10, x14)
        SDA_QR(x3, x7, x11, x15)
        SDA_QR(x0, x5, x10, x15)
        SDA_QR(x1, x6, x11, x12)
        SDA_QR(x2, x7, x8, x13)
        SDA_QR(x3, x4, x9, x14)
    }
    out[0] = x0 + c0;   out[1] = x1 + c1;   out[2] = x2 + c2;   out[3] = x3 + c3;
    out[4] = x4 + key.w[0];  out[5] = x5 + key.w[1];  out[6] = x6 + key.w[2];  out[7] = x7 + key.w[3];
    out[8] = x8 + key.w[4];  out[9] = x9 + key.w[5];  out[10] = x10 + key.w[6]; out[11] = x11 + key.w[7];
    out[12] = x12 + b0; out[13] = x13 + b1; out[14] = x14;      out[15] = x15;
}

// the 8 u64 draws of one block, in stream order
template <int ROUNDS>
__device__ __forceinline__ void chacha_draws8(const ChaChaKey &key, uint64_t block, uint64_t v[8]) {
    uint32_t w[16];
    chacha_block<ROUNDS>(key, block, w);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = ((uint64_t)w[2 * i] << 32) | w[2 * i + 1];
}
#endif

}  // namespace sda
