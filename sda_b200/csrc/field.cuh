// field.cuh -- arithmetic in Z_m for 0 < m < 2^63 (the reference's `i64` modulus), shared by
// every kernel.  Two reduction back ends:
//   * MERSENNE61: m = 2^61 - 1, folds with shifts and adds (2^61 == 1);
//   * GENERIC:    any m, Moeller-Granlund 2-by-1 division by a precomputed reciprocal.
// Host code builds FieldParams once per call; kernels take it by value (constant bank).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SDA_HD __host__ __device__ __forceinline__
#define SDA_D __device__ __forceinline__
#else
#define SDA_HD inline
#define SDA_D inline
#endif

namespace sda {

constexpr uint64_t P61 = (1ull << 61) - 1;

enum FieldKind : uint32_t { FIELD_GENERIC = 0, FIELD_MERSENNE61 = 1 };

struct FieldParams {
    uint64_t m;      // modulus
    uint64_t d;      // m << s, top bit set
    uint64_t v;      // floor((2^128 - 1) / d) - 2^64
    uint32_t s;      // leading zeros of m (>= 1 because m < 2^63)
    uint32_t kind;   // FieldKind
};

// rand-0.3 `Range<u64>`: accept v < zone, sample = v % range   
struct DrawParams {
    FieldParams f;   // f.m = range
    uint64_t zone;   // u64::MAX - u64::MAX % range
    uint32_t kind;   // 0 generic, 1 range = 2^61-1, 2 range = 2^61-2
};
enum DrawKind : uint32_t { DRAW_GENERIC = 0, DRAW_M61 = 1, DRAW_M61_MINUS1 = 2 };

inline FieldParams make_field(uint64_t m) {
    FieldParams f{};
    f.m = m;
    f.s = (uint32_t)__builtin_clzll(m);
    f.d = m << f.s;
    unsigned __int128 num = ((unsigned __int128)(~f.d) << 64) | ~0ull;
    f.v = (uint64_t)(num / f.d);
    f.kind = (m == P61) ? FIELD_MERSENNE61 : FIELD_GENERIC;
    return f;
}
inline DrawParams make_draw(uint64_t range) {
    DrawParams d{};
    d.f = make_field(range);
    d.zone = ~0ull - (~0ull % range);
    d.kind = range == P61 ? DRAW_M61 : (range == P61 - 1 ? DRAW_M61_MINUS1 : DRAW_GENERIC);
    return d;
}

#if defined(__CUDACC__)

// ---- generic: (u1:u0) mod m, requires u1 < m --------------------------------------------
SDA_D uint64_t reduce128_generic(const FieldParams &f, uint64_t u1, uint64_t u0) {
    const uint32_t s = f.s;                       // 1..63
    uint64_t n1 = (u1 << s) | (u0 >> (64 - s));
    uint64_t n0 = u0 << s;
    uint64_t q0 = f.v * n1;
    uint64_t q1 = __umul64hi(f.v, n1);
    q0 += n0;
    q1 += n1 + (q0 < n0) + 1;
    uint64_t r = n0 - q1 * f.d;
    if (r > q0) r += f.d;
    if (r >= f.d) r -= f.d;
    return r >> s;
}
SDA_D uint64_t reduce64_generic(const FieldParams &f, uint64_t x) { return reduce128_generic(f, 0, x); }

// ---- Mersenne 2^61-1 --------------------------------------------------------------------
// x < 2^64  ->  [0, p)
SDA_D uint64_t reduce64_m61(uint64_t x) {
    uint64_t r = (x & P61) + (x >> 61);           // < 2^61 + 8
    return r >= P61 ? r - P61 : r;
}
// (u1:u0) < 2^125  ->  [0, p)
SDA_D uint64_t reduce128_m61(uint64_t u1, uint64_t u0) {
    uint64_t mid = (u1 << 3) | (u0 >> 61);        // bits 61..124
    uint64_t r = (u0 & P61) + (mid & P61) + (mid >> 61);   // < 2^62 + 8
    r = (r & P61) + (r >> 61);
    return r >= P61 ? r - P61 : r;
}

template <bool M61>
SDA_D uint64_t reduce64(const FieldParams &f, uint64_t x) {
    if (M61) return reduce64_m61(x);
    return reduce64_generic(f, x);
}

// canonical residue of an arbitrary i64 (truncating % then lift, client/src/receive.rs:13-21)
template <bool M61>
SDA_D uint64_t canon(const FieldParams &f, int64_t v) {
    if ((uint64_t)v < f.m) return (uint64_t)v;    // already canonical: the common case
    uint64_t a = v < 0 ? 0ull - (uint64_t)v : (uint64_t)v;
    uint64_t r = reduce64<M61>(f, a);
    return (v < 0 && r) ? f.m - r : r;
}

SDA_D uint64_t addmod(uint64_t a, uint64_t b, uint64_t m) {   // a, b < m < 2^63
    uint64_t r = a + b;
    return r >= m ? r - m : r;
}
SDA_D uint64_t submod(uint64_t a, uint64_t b, uint64_t m) {
    return a >= b ? a - b : a + m - b;
}

// signed 128-bit accumulator for column sums of arbitrary i64 (no per-element reduction)
struct Acc128 {
    uint64_t lo;
    int64_t hi;
    SDA_D void init() { lo = 0; hi = 0; }
    SDA_D void add(int64_t v) {
        uint64_t nl = lo + (uint64_t)v;
        hi += (v >> 63) + (nl < lo);
        lo = nl;
    }
};
// S + bias mod m, where bias = m * ceil(rows * 2^63 / m) >= |S| is host-computed (128 bit)
// and bias_hi + rows < m is NOT assumed: the high word is reduced first.
template <bool M61>
SDA_D uint64_t reduce_acc(const FieldParams &f, Acc128 a, uint64_t bias_hi, uint64_t bias_lo) {
    uint64_t lo = a.lo + bias_lo;
    uint64_t hi = (uint64_t)a.hi + bias_hi + (lo < a.lo);
    if (M61) {
        // hi < 2^61 is guaranteed for rows < 2^60; fold 2^64 == 8
        uint64_t h = reduce64_m61(hi);
        return reduce128_m61(h, lo);
    }
    uint64_t h = reduce64_generic(f, hi);
    return reduce128_generic(f, h, lo);
}

// gen_range(0, range): returns the sample; sets rejected when the word is outside the zone
template <uint32_t DK>
SDA_D uint64_t draw_reduce(const DrawParams &d, uint64_t v, bool &rejected) {
    rejected = v >= d.zone;
    if (DK == DRAW_M61) return reduce64_m61(v);
    if (DK == DRAW_M61_MINUS1) {                  // 2^61 == 2 (mod 2^61 - 2)
        uint64_t r = (v & ((1ull << 61) - 1)) + 2 * (v >> 61);   // < 2^61 + 14
        return r >= P61 - 1 ? r - (P61 - 1) : r;
    }
    return reduce64_generic(d.f, v);
}

#endif  // __CUDACC__

}  // namespace sda
