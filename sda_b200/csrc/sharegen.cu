// sharegen.cu -- K1 (additive split, Full / ChaCha masks) and K2 (packed-Shamir share
// generation): the participant side of the hot path.
//
//   additive split   client/src/crypto/sharing/additive.rs:32-51  (+ batched.rs:18-53)
//   packed Shamir    client/src/crypto/sharing/packed_shamir.rs:40-43 -> tss 0.2 `share`
//                    (+ batched.rs:18-53: zero-padded last batch, clerk-major scatter)
//   Full mask        client/src/crypto/masking/full.rs:21-35
//   ChaCha mask      client/src/crypto/masking/chacha.rs:24-54 (participant), :56-77 (recipient)
//
// Randomness never touches HBM: draw q of a participant's stream is the q-th u64 of
// ChaCha(key, rounds) (rand-0.3 word order) and a thread owns whole 64-byte keystream blocks
// = 8 draws.  A thread therefore processes a *unit* of G = 8/gcd(d,8) consecutive elements
// (d = draws per element: n-1 for additive, t for packed, 1 for masks), which makes its
// input a contiguous run of G*k i64 and its output n contiguous runs of G i64.
// gen_range's rejection (probability ~2^-60 for the supported moduli) shifts the whole
// stream, so a rejected word raises `flag` and the host redoes the call on the exact path
// (draws pre-computed by draw_exact into scratch, kernels instantiated with FROM_MEM).
#include "chacha_pre.cuh"
#include "kernels.h"
#include "vecio.cuh"

namespace sda {

namespace {

constexpr int CTA = 128;

constexpr int gcd_c(int a, int b) { return b == 0 ? a : gcd_c(b, a % b); }
template <int D>
struct Unit {
    static constexpr int G = 8 / gcd_c(D, 8);    // elements (batches) per thread
    static constexpr int NB = D * G / 8;         // keystream blocks per thread
};

__device__ __forceinline__ ChaChaKey load_key(const ChaChaKey *keys, size_t p) {
    ChaChaKey k;
    const uint4 *src = reinterpret_cast<const uint4 *>(keys + p);
    uint4 a = __ldg(src), b = __ldg(src + 1);
    k.w[0] = a.x; k.w[1] = a.y; k.w[2] = a.z; k.w[3] = a.w;
    k.w[4] = b.x; k.w[5] = b.y; k.w[6] = b.z; k.w[7] = b.w;
    return k;
}

// the 8 u64 draws of block `block` (< 2^32) of a key whose first-round constants were precomputed (chacha_pre.cuh)
template <int ROUNDS>
__device__ __forceinline__ void chacha_words_pre(const ChaChaKey &key, const ChaChaPre &pre, uint32_t block, uint32_t (&w)[16]) {
    chacha_block2<ROUNDS>(key.w, pre.w, block, w);
}
__device__ __forceinline__ ChaChaPre load_pre(const ChaChaPre *pres, size_t p) {
    ChaChaPre r;
    const uint4 *src = reinterpret_cast<const uint4 *>(pres + p);
    const uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
    r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
    r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
    r.w[8] = c.x; r.w[9] = c.y; r.w[10] = c.z; r.w[11] = c.w;
    return r;
}

// ------------------------------------------------------------------------------------------
// additive split, D = n - 1 draws per element, in-kernel randomness
// ------------------------------------------------------------------------------------------
template <bool M61, uint32_t DK, int ROUNDS, int D>
__global__ void __launch_bounds__(CTA)
additive_split_kernel(const int64_t *__restrict__ secrets, size_t ld, size_t dim, const ChaChaKey *__restrict__ keys,
                      int64_t *__restrict__ out, FieldParams f, DrawParams dr, int in_lanes, int out_lanes,
                      unsigned *flag, const ChaChaPre *__restrict__ pres = nullptr) {
    constexpr int G = Unit<D>::G, NB = Unit<D>::NB;
    const size_t p = blockIdx.y;
    const size_t u = (size_t)blockIdx.x * CTA + threadIdx.x;
    const size_t e0 = u * G;
    if (e0 >= dim) return;
    const int nvalid = (int)min((size_t)G, dim - e0);
    const ChaChaKey key = load_key(keys, p);
    // pres != nullptr (Mersenne path, every block counter below 2^32): the participant's first-round constants
    ChaChaPre pre{};
    if (M61 && pres != nullptr) pre = load_pre(pres, p);
    auto next_block = [&](size_t b, uint64_t (&dst)[8]) {
        if (M61 && pres != nullptr) {
            uint32_t w[16];
            chacha_words_pre<ROUNDS>(key, pre, (uint32_t)b, w);
#pragma unroll
            for (int i = 0; i < 8; i++) dst[i] = ((uint64_t)w[2 * i] << 32) | w[2 * i + 1];
        } else {
            chacha_draws8<ROUNDS>(key, b, dst);
        }
    };

    int64_t x[G];
    load_run<G>(secrets + p * ld + e0, x, nvalid, in_lanes);

    int64_t sh[D + 1][G];
    uint64_t blk[8];
    bool rej = false;
    if constexpr (M61) {
        // Mersenne path with as few ALU-pipe instructions as the arithmetic allows (the kernel is bound by
        // ChaCha on that pipe, profiles/r01_pipes.md): a draw is (v & p) + (v >> 61) with NO compare -- that is
        // gen_range's v % p unless the low 61 bits are within 8 of 2^61, which is detected through a carry into
        // bit 29 of (hi + 1) and settled exactly in a rare branch -- and the last share is reduced once from
        // x + D p - sum s_j instead of by D conditional subtractions (additive.rs:47, same residue).
        constexpr uint32_t LOW29 = 0x1fffffffu;
#pragma unroll
        for (int e = 0; e < G; e++) {
            uint64_t xe = (uint64_t)x[e];
            if (xe >> 61) xe = canon<true>(f, x[e]);          // any i64 is a legal secret; [0, 2^61) needs no work
            uint64_t t = xe + (uint64_t)D * P61;              // < 2^61 + 4 p
#pragma unroll
            for (int j = 0; j < D; j++) {
                const int q = e * D + j;
                if (q % 8 == 0) next_block(u * NB + q / 8, blk);
                const uint64_t v = blk[q % 8];
                const uint32_t w0 = (uint32_t)(v >> 32), hi = w0 & LOW29;
                uint64_t s = (((uint64_t)hi << 32) | (uint32_t)v) + (w0 >> 29);
                if ((hi + 1u) >> 29) {                        // bits 32..60 all ones: 2^-29 per draw
                    if (v >= dr.zone) rej |= e < nvalid;      // a word gen_range rejects: the stream shifts
                    if (s >= P61) s -= P61;
                }
                sh[j][e] = (int64_t)s;                        // additive.rs:42-44
                t -= s;
            }
            uint64_t r = (t & P61) + (t >> 61);               // <= p + 4
            r = r >= P61 ? r - P61 : r;
            sh[D][e] = (int64_t)r;
        }
    } else {
#pragma unroll
        for (int e = 0; e < G; e++) {
            uint64_t acc = canon<M61>(f, x[e]);
#pragma unroll
            for (int j = 0; j < D; j++) {
                const int q = e * D + j;
                if (q % 8 == 0) chacha_draws8<ROUNDS>(key, u * NB + q / 8, blk);
                bool r;
                const uint64_t s = draw_reduce<DK>(dr, blk[q % 8], r);
                rej |= r && e < nvalid;
                sh[j][e] = (int64_t)s;                      // additive.rs:42-44
                acc = submod(acc, s, f.m);                  // additive.rs:47
            }
            sh[D][e] = (int64_t)acc;
        }
    }
    int64_t *o = out + (p * (D + 1)) * dim + e0;
#pragma unroll
    for (int j = 0; j <= D; j++) store_run<G>(o + (size_t)j * dim, sh[j], nvalid, out_lanes);
    if (rej) atomicOr(flag, 1u);
}

// additive split for ANY share count, in-kernel randomness (additive.rs:32-51 with n - 1 = D a run-time value): a thread
// owns 8 adjacent elements, i.e. exactly D whole keystream blocks (draw q = e D + j of the participant's stream is share j
// of element e).  The running "secret - sum of draws" of the 8 elements lives in shared memory ([element][thread], conflict
// free) because the element a draw belongs to is a run-time index; shares are stored as they are drawn.
template <bool M61, uint32_t DK, int ROUNDS>
__global__ void __launch_bounds__(CTA)
additive_split_any_kernel(const int64_t *__restrict__ secrets, size_t ld, size_t dim, int D, const ChaChaKey *__restrict__ keys,
                          int64_t *__restrict__ out, FieldParams f, DrawParams dr, unsigned *flag) {
    __shared__ uint64_t acc[8][CTA];
    const size_t p = blockIdx.y;
    const size_t u = (size_t)blockIdx.x * CTA + threadIdx.x;
    const size_t e0 = u * 8;
    if (e0 >= dim) return;
    const int nvalid = (int)min((size_t)8, dim - e0);
    const ChaChaKey key = load_key(keys, p);
#pragma unroll
    for (int e = 0; e < 8; e++) acc[e][threadIdx.x] = e < nvalid ? canon<M61>(f, secrets[p * ld + e0 + e]) : 0;
    int64_t *o = out + p * (size_t)(D + 1) * dim + e0;
    bool rej = false;
    int e = 0, j = 0;
    for (int b = 0; b < D; b++) {
        uint64_t blk[8];
        chacha_draws8<ROUNDS>(key, u * (size_t)D + b, blk);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            bool r;
            const uint64_t s = draw_reduce<DK>(dr, blk[i], r);
            if (e < nvalid) {
                rej |= r;
                o[(size_t)j * dim + e] = (int64_t)s;                            // additive.rs:42-44
                acc[e][threadIdx.x] = submod(acc[e][threadIdx.x], s, f.m);      // additive.rs:47
            }
            if (++j == D) {
                j = 0;
                e++;
            }
        }
    }
    for (int k = 0; k < nvalid; k++) o[(size_t)D * dim + k] = (int64_t)acc[k][threadIdx.x];
    if (rej) atomicOr(flag, 1u);
}

// any n, draws read from memory (exact path): one element per thread
template <bool M61>
__global__ void __launch_bounds__(CTA)
additive_split_mem_kernel(const int64_t *__restrict__ secrets, size_t ld, size_t dim, int n,
                          const uint64_t *__restrict__ draws, int64_t *__restrict__ out, FieldParams f) {
    const size_t p = blockIdx.y;
    const size_t e = (size_t)blockIdx.x * CTA + threadIdx.x;
    if (e >= dim) return;
    uint64_t acc = canon<M61>(f, secrets[p * ld + e]);
    const uint64_t *d = draws + (p * dim + e) * (size_t)(n - 1);
    int64_t *o = out + p * (size_t)n * dim + e;
    for (int j = 0; j < n - 1; j++) {
        const uint64_t s = d[j];
        o[(size_t)j * dim] = (int64_t)s;
        acc = submod(acc, s, f.m);
    }
    o[(size_t)(n - 1) * dim] = (int64_t)acc;
}

// ------------------------------------------------------------------------------------------
// masks: one draw per element
// ------------------------------------------------------------------------------------------
// FLOAT_IN: the secrets are real values; they are brought to fixed point here (q = rint(x 2^frac_bits) as a canonical
// residue, the definition of sda_fixed_encode_dev) instead of being read as i64 -- the fused encode + mask of a model
// update (4 B in, 8 B out per element, no intermediate i64 vector).
template <bool M61, uint32_t DK, int ROUNDS, bool FROM_MEM, bool FLOAT_IN = false>
__global__ void __launch_bounds__(CTA)
mask_kernel(const int64_t *__restrict__ secrets, size_t dim, ChaChaKey key, const uint64_t *__restrict__ draws,
            int64_t *__restrict__ mask_out, int64_t *__restrict__ masked_out, FieldParams f, DrawParams dr,
            int lanes, unsigned *flag, const float *__restrict__ fx = nullptr, double scale = 1.0, ChaChaPre pre = ChaChaPre{},
            int use_pre = 0) {
    constexpr int G = 8;
    const size_t u = (size_t)blockIdx.x * CTA + threadIdx.x;
    const size_t e0 = u * G;
    if (e0 >= dim) return;
    const int nvalid = (int)min((size_t)G, dim - e0);
    int64_t x[G], mk[G], md[G];
    if constexpr (FLOAT_IN) {
        float v[G];
        if (nvalid == G && (reinterpret_cast<uintptr_t>(fx + e0) & 15) == 0) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(fx + e0)), b = __ldg(reinterpret_cast<const float4 *>(fx + e0) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < G; e++) v[e] = e < nvalid ? __ldg(fx + e0 + e) : 0.f;
        }
        if constexpr (M61) {
            // x 2^frac_bits is exact in float (a power-of-two scale; overflow saturates the conversion exactly as the
            // double product would), and |q| < 2^60 -- every realistic update -- needs one masked add to become a residue
            const float sf = (float)scale;
#pragma unroll
            for (int e = 0; e < G; e++) {
                const long long q = __float2ll_rn(v[e] * sf);
                uint64_t r = (uint64_t)q + (P61 & (uint64_t)(q >> 63));
                if (((uint64_t)q + (1ull << 60)) >> 61) r = canon<true>(f, (int64_t)q);          // |q| >= 2^60: the general way
                x[e] = (int64_t)r;
            }
        } else {
#pragma unroll
            for (int e = 0; e < G; e++) x[e] = (int64_t)canon<M61>(f, __double2ll_rn((double)v[e] * scale));
        }
    } else {
        load_run<G>(secrets + e0, x, nvalid, lanes);
    }
    uint64_t blk[8];
    bool rej = false;
    if (FROM_MEM) {
#pragma unroll
        for (int e = 0; e < G; e++) blk[e] = e < nvalid ? draws[e0 + e] : 0;
    } else if (use_pre) {                        // every block counter below 2^32: first-round constants from the host
        uint32_t w[16];
        chacha_words_pre<ROUNDS>(key, pre, (uint32_t)u, w);
#pragma unroll
        for (int e = 0; e < 8; e++) blk[e] = ((uint64_t)w[2 * e] << 32) | w[2 * e + 1];
    } else {
        chacha_draws8<ROUNDS>(key, u, blk);
    }
    if constexpr (M61 && DK == DRAW_M61 && !FROM_MEM) {
        // modulus 2^61 - 1, in-kernel draws: the arithmetic of the additive split's Mersenne path (above) -- a draw is
        // (v & p) + (v >> 61) with no compare, secrets below 2^61 are taken as they are, and the sum is folded once.
        // The two exceptions (a draw whose low 61 bits are within 8 of 2^61: 2^-29 per draw; a secret outside
        // [0, 2^61)) are detected for the whole thread with a running maximum / a running OR and settled below.
        constexpr uint32_t LOW29 = 0x1fffffffu;
        uint32_t suspect = 0, wide = 0;
#pragma unroll
        for (int e = 0; e < G; e++) {
            const uint64_t v = blk[e];
            const uint32_t w0 = (uint32_t)(v >> 32), hi = w0 & LOW29;
            const uint64_t sd = (((uint64_t)hi << 32) | (uint32_t)v) + (w0 >> 29);
            suspect = max(suspect, hi);
            wide |= (uint32_t)((uint64_t)x[e] >> 61);
            // secret <= p (below 2^61) and, off the suspect path, sd < p: the sum is below 2 p, one conditional subtraction
            const uint64_t t = (uint64_t)x[e] + sd;
            const uint64_t r = t >= P61 ? t - P61 : t;
            mk[e] = (int64_t)sd;                                        // full.rs:24-27
            md[e] = (int64_t)r;                                         // full.rs:28-31
        }
        if (suspect == LOW29 || wide != 0) {
#pragma unroll
            for (int e = 0; e < G; e++) {
                bool r;
                const uint64_t sd = draw_reduce<DK>(dr, blk[e], r);
                rej |= r && e < nvalid;
                mk[e] = (int64_t)sd;
                md[e] = (int64_t)addmod(canon<M61>(f, x[e]), sd, f.m);
            }
        }
    } else {
#pragma unroll
        for (int e = 0; e < G; e++) {
            uint64_t s;
            if (FROM_MEM) {
                s = blk[e];
            } else {
                bool r;
                s = draw_reduce<DK>(dr, blk[e], r);
                rej |= r && e < nvalid;
            }
            mk[e] = (int64_t)s;                                            // full.rs:24-27
            md[e] = (int64_t)addmod(canon<M61>(f, x[e]), s, f.m);          // full.rs:28-31
        }
    }
    if (mask_out != nullptr) store_run<G>(mask_out + e0, mk, nvalid, lanes);
    store_run<G>(masked_out + e0, md, nvalid, lanes);
    if (!FROM_MEM && rej) atomicOr(flag, 1u);
}

// recipient-side ChaCha mask combine (chacha.rs:60-73): sum over P seeds of the re-expanded
// masks.  Compute-bound by construction (P keystream blocks per 8 outputs, no input traffic).
// grid.y slices the seed axis; slice partials (canonical) go to out + slice*out_ld.
template <bool M61, uint32_t DK>
__global__ void __launch_bounds__(CTA)
chacha_mask_combine_kernel(const ChaChaKey *__restrict__ keys, size_t P, size_t seeds_per_slice, size_t dim,
                           int64_t *__restrict__ out, size_t out_ld, FieldParams f, DrawParams dr, int lanes,
                           unsigned *flag, const ChaChaPre *__restrict__ pres = nullptr) {
    constexpr int G = 8;
    const size_t u = (size_t)blockIdx.x * CTA + threadIdx.x;
    const size_t e0 = u * G;
    if (e0 >= dim) return;
    const int nvalid = (int)min((size_t)G, dim - e0);
    const size_t p0 = (size_t)blockIdx.y * seeds_per_slice;
    const size_t p1 = min(P, p0 + seeds_per_slice);
    uint64_t acc[G];
#pragma unroll
    for (int e = 0; e < G; e++) acc[e] = 0;
    bool rej = false;
    if constexpr (M61 && DK == DRAW_M61) {
        if (pres != nullptr) {
            // 2^61 - 1 with precomputed first-round constants (the launcher passes them when every block counter is below
            // 2^32): a draw is (v & p) + (v >> 61) <= p + 7 with no compare, summed as it is and folded every 7 seeds; a
            // rejected word (v >= 2^64 - 8: high word all ones) is looked for only when the block's largest high word says so
            constexpr uint32_t LOW29 = 0x1fffffffu;
            uint32_t since_fold = 0;
            for (size_t p = p0; p < p1; p++) {
                const ChaChaKey key = load_key(keys, p);
                const ChaChaPre pre = load_pre(pres, p);
                uint32_t w[16];
                chacha_words_pre<20>(key, pre, (uint32_t)u, w);
                uint32_t top = 0;
#pragma unroll
                for (int e = 0; e < G; e++) {
                    top = max(top, w[2 * e]);
                    acc[e] += (((uint64_t)(w[2 * e] & LOW29) << 32) | w[2 * e + 1]) + (w[2 * e] >> 29);
                }
                if (top == 0xffffffffu) {
#pragma unroll
                    for (int e = 0; e < G; e++) rej |= e < nvalid && w[2 * e] == 0xffffffffu && w[2 * e + 1] >= 0xfffffff8u;
                }
                if (++since_fold == 7) {
                    since_fold = 0;
#pragma unroll
                    for (int e = 0; e < G; e++) acc[e] = (acc[e] & P61) + (acc[e] >> 61);
                }
            }
#pragma unroll
            for (int e = 0; e < G; e++) {
                uint64_t a = (acc[e] & P61) + (acc[e] >> 61);
                a = (a & P61) + (a >> 61);
                acc[e] = a >= P61 ? a - P61 : a;
            }
        }
    }
    for (size_t p = (M61 && DK == DRAW_M61 && pres != nullptr) ? p1 : p0; p < p1; p++) {
        const ChaChaKey key = load_key(keys, p);
        uint64_t blk[8];
        chacha_draws8<20>(key, u, blk);
#pragma unroll
        for (int e = 0; e < G; e++) {
            bool r;
            const uint64_t s = draw_reduce<DK>(dr, blk[e], r);
            rej |= r && e < nvalid;
            acc[e] = addmod(acc[e], s, f.m);
        }
    }
    int64_t r[G];
#pragma unroll
    for (int e = 0; e < G; e++) r[e] = (int64_t)acc[e];
    store_run<G>(out + (size_t)blockIdx.y * out_ld + e0, r, nvalid, lanes);
    if (rej) atomicOr(flag, 1u);
}

// ------------------------------------------------------------------------------------------
// packed Shamir: shares = M . [secrets ; randomness] per batch
// ------------------------------------------------------------------------------------------
template <int N, int W>
struct MatSplit {            // generic field: canonical entries as two 32-bit words
    uint32_t lo[N * W];
    uint32_t hi[N * W];
};

// generic field: 128-bit accumulation, folded every `lazy` terms (lazy * (m-1)^2 + m < m 2^64)
template <int N, int W>
__device__ __forceinline__ uint64_t dot_generic(const FieldParams &f, int lazy, const MatSplit<N, W> &m, int j,
                                                const uint64_t (&x)[W]) {
    uint64_t lo = 0, hi = 0;
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < W; i++) {
        const uint64_t a = ((uint64_t)m.hi[j * W + i] << 32) | m.lo[j * W + i];
        const uint64_t pl = a * x[i], ph = __umul64hi(a, x[i]);
        lo += pl;
        hi += ph + (lo < pl);
        if (++cnt == lazy && i + 1 < W) {
            lo = reduce128_generic(f, hi, lo);
            hi = 0;
            cnt = 0;
        }
    }
    return reduce128_generic(f, hi, lo);
}

template <int K, int T, int N, bool M61, uint32_t DK, int ROUNDS>
__global__ void __launch_bounds__(CTA)
packed_share_kernel(const int64_t *__restrict__ secrets, size_t ld, size_t dim, size_t B,
                    const ChaChaKey *__restrict__ keys, int64_t *__restrict__ out, FieldParams f, DrawParams dr,
                    int lazy, MatSplit<N, K + T> mat, int in_lanes, int out_lanes,
                    unsigned *flag) {
    constexpr int W = K + T;
    constexpr int G = Unit<T>::G, NB = Unit<T>::NB;
    const size_t p = blockIdx.y;
    const size_t u = (size_t)blockIdx.x * CTA + threadIdx.x;
    const size_t b0 = u * G;
    if (b0 >= B) return;
    const int nb = (int)min((size_t)G, B - b0);
    const ChaChaKey key = load_key(keys, p);

    // contiguous run of G*K secrets; the last batch is zero padded (batched.rs:38-43)
    int64_t s[G * K];
    const size_t s0 = b0 * K;
    const size_t avail = dim - s0;
    load_run<G * K>(secrets + p * ld + s0, s, (int)min((size_t)(G * K), avail), in_lanes);

    int64_t sh[N][G];
    uint64_t blk[8];
    bool rej = false;
#pragma unroll
    for (int g = 0; g < G; g++) {
        uint64_t x[W];
#pragma unroll
        for (int i = 0; i < K; i++) x[i] = canon<M61>(f, s[g * K + i]);
#pragma unroll
        for (int i = 0; i < T; i++) {
            const int q = g * T + i;
            if (q % 8 == 0) chacha_draws8<ROUNDS>(key, u * NB + q / 8, blk);
            bool r;
            x[K + i] = draw_reduce<DK>(dr, blk[q % 8], r);   // tss share(): Range::new(0, prime - 1)
            rej |= r && g < nb;
        }
#pragma unroll
        for (int j = 0; j < N; j++) sh[j][g] = (int64_t)dot_generic<N, W>(f, lazy, mat, j, x);
    }
    int64_t *o = out + (p * N) * B + b0;
#pragma unroll
    for (int j = 0; j < N; j++) store_run<G>(o + (size_t)j * B, sh[j], nb, out_lanes);
    if (rej) atomicOr(flag, 1u);
}

// any (k, t, n): one batch per thread, matrix in shared memory, draws from memory
template <bool M61>
__global__ void __launch_bounds__(CTA)
packed_share_mem_kernel(const int64_t *__restrict__ secrets, size_t ld, size_t dim, size_t B, int k, int t, int n,
                        const uint64_t *__restrict__ mat_g, const uint64_t *__restrict__ draws,
                        int64_t *__restrict__ out, FieldParams f, int lazy) {
    __shared__ uint64_t mat[MAX_N * MAX_W];
    const int w = k + t;
    for (int i = threadIdx.x; i < n * w; i += CTA) mat[i] = mat_g[i];
    __syncthreads();
    const size_t p = blockIdx.y;
    const size_t b = (size_t)blockIdx.x * CTA + threadIdx.x;
    if (b >= B) return;
    uint64_t x[MAX_W];
    for (int i = 0; i < k; i++) {
        const size_t e = b * k + i;
        x[i] = e < dim ? canon<M61>(f, secrets[p * ld + e]) : 0;
    }
    const uint64_t *d = draws + (p * B + b) * (size_t)t;
    for (int i = 0; i < t; i++) x[k + i] = d[i];
    for (int j = 0; j < n; j++) {
        uint64_t lo = 0, hi = 0;
        int cnt = 0;
        for (int i = 0; i < w; i++) {
            const uint64_t a = mat[j * w + i];
            const uint64_t pl = a * x[i], ph = __umul64hi(a, x[i]);
            lo += pl;
            hi += ph + (lo < pl);
            if (++cnt == lazy) {
                lo = reduce128_generic(f, hi, lo);
                hi = 0;
                cnt = 0;
            }
        }
        out[(p * n + j) * B + b] = (int64_t)reduce128_generic(f, hi, lo);
    }
}

// ---- host-side dispatch -------------------------------------------------------------------

template <int N, int W>
MatSplit<N, W> split_matrix(const Matrix &m) {
    MatSplit<N, W> s;
    for (int i = 0; i < N * W; i++) {
        s.lo[i] = (uint32_t)m.e[i];
        s.hi[i] = (uint32_t)(m.e[i] >> 32);
    }
    return s;
}

template <int K, int T, int N, bool M61, uint32_t DK, int ROUNDS>
cudaError_t packed_launch(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, int lazy,
                          const Matrix &mtx, const int64_t *secrets, size_t ld, size_t P, size_t dim,
                          const ChaChaKey *keys, int64_t *out, unsigned *flag) {
    constexpr int G = Unit<T>::G;
    const size_t B = (dim + K - 1) / K;
    const size_t units = (B + G - 1) / G;
    dim3 grid((unsigned)((units + CTA - 1) / CTA), (unsigned)P);
    const int in_lanes = pick_lanes(secrets, ld, G * K);
    const int out_lanes = pick_lanes(out, B, G);
    packed_share_kernel<K, T, N, M61, DK, ROUNDS><<<grid, CTA, 0, lc.stream>>>(
        secrets, ld, dim, B, keys, out, f, dr, lazy, split_matrix<N, K + T>(mtx), in_lanes, out_lanes, flag);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

template <int K, int T, int N>
cudaError_t packed_dispatch(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, int rounds, int lazy,
                            const Matrix &mtx, const int64_t *secrets, size_t ld, size_t P, size_t dim,
                            const ChaChaKey *keys, int64_t *out, unsigned *flag) {
#define SDA_PL(M61, DK, R) \
    return packed_launch<K, T, N, M61, DK, R>(lc, f, dr, lazy, mtx, secrets, ld, P, dim, keys, out, flag)
    if (rounds == 8) SDA_PL(false, DRAW_GENERIC, 8);
    if (rounds == 12) SDA_PL(false, DRAW_GENERIC, 12);
    SDA_PL(false, DRAW_GENERIC, 20);
#undef SDA_PL
}

template <bool M61, uint32_t DK, int ROUNDS, int D>
cudaError_t additive_launch(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, const int64_t *secrets,
                            size_t ld, size_t P, size_t dim, const ChaChaKey *keys, int64_t *out, unsigned *flag,
                            uint32_t *d_pre) {
    constexpr int G = Unit<D>::G;
    const size_t units = (dim + G - 1) / G;
    dim3 grid((unsigned)((units + CTA - 1) / CTA), (unsigned)P);
    const int in_lanes = pick_lanes(secrets, ld, G);
    const int out_lanes = pick_lanes(out, dim, G);
    ChaChaPre *pres = nullptr;
    if (M61 && d_pre != nullptr && (((dim * (size_t)D + 7) / 8 + 8) >> 32) == 0) {
        pres = reinterpret_cast<ChaChaPre *>(d_pre);
        chacha_prepare_kernel<<<(unsigned)((P + 127) / 128), 128, 0, lc.stream>>>(keys, P, pres);
        ++*lc.nlaunch;
    }
    additive_split_kernel<M61, DK, ROUNDS, D><<<grid, CTA, 0, lc.stream>>>(secrets, ld, dim, keys, out, f, dr,
                                                                            in_lanes, out_lanes, flag, pres);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

template <int D>
cudaError_t additive_dispatch(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, int rounds,
                              const int64_t *secrets, size_t ld, size_t P, size_t dim, const ChaChaKey *keys,
                              int64_t *out, unsigned *flag, uint32_t *d_pre) {
#define SDA_AL(M61, DK, R) return additive_launch<M61, DK, R, D>(lc, f, dr, secrets, ld, P, dim, keys, out, flag, d_pre)
    if (f.kind == FIELD_MERSENNE61) {
        if (rounds == 8) SDA_AL(true, DRAW_M61, 8);
        if (rounds == 12) SDA_AL(true, DRAW_M61, 12);
        SDA_AL(true, DRAW_M61, 20);
    }
    if (rounds == 8) SDA_AL(false, DRAW_GENERIC, 8);
    if (rounds == 12) SDA_AL(false, DRAW_GENERIC, 12);
    SDA_AL(false, DRAW_GENERIC, 20);
#undef SDA_AL
}

// largest number of (m-1)^2 products that can be summed on top of a residue before the
// 128-bit accumulator's high word reaches m
int lazy_terms(uint64_t m) {
    unsigned __int128 cap = ((unsigned __int128)m << 64) - m;
    unsigned __int128 sq = (unsigned __int128)(m - 1) * (m - 1);
    if (sq == 0) return 1 << 20;
    unsigned __int128 q = cap / sq;
    if (q > (1u << 20)) q = 1u << 20;
    return q ? (int)q : 1;
}

}  // namespace

bool packed_share_has_fast_path(int k, int t, int n) {
    return (k == 3 && t == 2 && n == 5) || (k == 5 && t == 4 && n == 9) || (k == 3 && t == 4 && n == 7) ||
           (k == 3 && t == 4 && n == 8);
}
bool additive_split_has_fast_path(int n) { return n >= 2; }   // n <= 5: unrolled kernel; above: the run-time one

cudaError_t launch_additive_split(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, int rounds, int n,
                                  const int64_t *secrets, size_t ld, size_t P, size_t dim, const ChaChaKey *keys,
                                  const uint64_t *draws, int64_t *shares_out, unsigned *flag, uint32_t *d_key_scratch) {
    if (dim == 0 || P == 0) return cudaSuccess;
    if (P > 65535) return cudaErrorInvalidValue;
    if (draws == nullptr && n >= 2 && n <= 5) {
        *lc.kernel_name = f.kind == FIELD_MERSENNE61 ? "additive_split<in-kernel rng>/mersenne61"
                                                     : "additive_split<in-kernel rng>/generic";
        switch (n - 1) {
        case 1: return additive_dispatch<1>(lc, f, dr, rounds, secrets, ld, P, dim, keys, shares_out, flag, d_key_scratch);
        case 2: return additive_dispatch<2>(lc, f, dr, rounds, secrets, ld, P, dim, keys, shares_out, flag, d_key_scratch);
        case 3: return additive_dispatch<3>(lc, f, dr, rounds, secrets, ld, P, dim, keys, shares_out, flag, d_key_scratch);
        case 4: return additive_dispatch<4>(lc, f, dr, rounds, secrets, ld, P, dim, keys, shares_out, flag, d_key_scratch);
        }
    }
    if (draws == nullptr && n > 5) {
        const bool m61 = f.kind == FIELD_MERSENNE61;
        *lc.kernel_name = m61 ? "additive_split<in-kernel rng, run-time share count>/mersenne61"
                              : "additive_split<in-kernel rng, run-time share count>/generic";
        dim3 grid((unsigned)(((dim + 7) / 8 + CTA - 1) / CTA), (unsigned)P);
#define SDA_AA(M61, DK, R) additive_split_any_kernel<M61, DK, R><<<grid, CTA, 0, lc.stream>>>(secrets, ld, dim, n - 1, keys, shares_out, f, dr, flag)
        if (m61) {
            if (rounds == 8) SDA_AA(true, DRAW_M61, 8);
            else if (rounds == 12) SDA_AA(true, DRAW_M61, 12);
            else SDA_AA(true, DRAW_M61, 20);
        } else {
            if (rounds == 8) SDA_AA(false, DRAW_GENERIC, 8);
            else if (rounds == 12) SDA_AA(false, DRAW_GENERIC, 12);
            else SDA_AA(false, DRAW_GENERIC, 20);
        }
#undef SDA_AA
        ++*lc.nlaunch;
        return cudaGetLastError();
    }
    if (draws == nullptr && n > 1) return cudaErrorInvalidValue;   // caller must pre-draw
    *lc.kernel_name = "additive_split<draws from memory>";
    dim3 grid((unsigned)((dim + CTA - 1) / CTA), (unsigned)P);
    if (f.kind == FIELD_MERSENNE61)
        additive_split_mem_kernel<true><<<grid, CTA, 0, lc.stream>>>(secrets, ld, dim, n, draws, shares_out, f);
    else
        additive_split_mem_kernel<false><<<grid, CTA, 0, lc.stream>>>(secrets, ld, dim, n, draws, shares_out, f);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

cudaError_t launch_mask(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, int rounds,
                        const int64_t *secrets, size_t dim, const ChaChaKey &key, const uint64_t *draws,
                        int64_t *mask_out, int64_t *masked_out, unsigned *flag, const float *fx, int frac_bits) {
    if (dim == 0) return cudaSuccess;
    const size_t units = (dim + 7) / 8;
    const unsigned grid = (unsigned)((units + CTA - 1) / CTA);
    const ChaChaPre pre = chacha_prepare_host(key);
    const int use_pre = (units >> 32) == 0;
    if (fx != nullptr) {                     // fused fixed-point encode + mask, in-kernel draws only
        if (draws != nullptr) return cudaErrorInvalidValue;
        const double scale = ldexp(1.0, frac_bits);
        int fl = pick_lanes(masked_out, 4, 8);
        if (mask_out && pick_lanes(mask_out, 4, 8) < fl) fl = pick_lanes(mask_out, 4, 8);
#define SDA_MF(M61, DK, R) mask_kernel<M61, DK, R, false, true><<<grid, CTA, 0, lc.stream>>>(nullptr, dim, key, nullptr, mask_out, masked_out, f, dr, fl, flag, fx, scale, pre, use_pre)
        if (f.kind == FIELD_MERSENNE61) {
            if (rounds == 8) SDA_MF(true, DRAW_M61, 8);
            else if (rounds == 12) SDA_MF(true, DRAW_M61, 12);
            else SDA_MF(true, DRAW_M61, 20);
        } else {
            if (rounds == 8) SDA_MF(false, DRAW_GENERIC, 8);
            else if (rounds == 12) SDA_MF(false, DRAW_GENERIC, 12);
            else SDA_MF(false, DRAW_GENERIC, 20);
        }
#undef SDA_MF
        ++*lc.nlaunch;
        return cudaGetLastError();
    }
    int lanes = pick_lanes(secrets, 4, 8);
    const int l2 = pick_lanes(masked_out, 4, 8), l3 = mask_out ? pick_lanes(mask_out, 4, 8) : 4;
    if (l2 < lanes) lanes = l2;
    if (l3 < lanes) lanes = l3;
#define SDA_ML(M61, DK, R, MEM)                                                                               \
    mask_kernel<M61, DK, R, MEM><<<grid, CTA, 0, lc.stream>>>(secrets, dim, key, draws, mask_out, masked_out, f, \
                                                              dr, lanes, flag, nullptr, 1.0, pre, use_pre)
    const bool m61 = f.kind == FIELD_MERSENNE61;
    if (draws != nullptr) {
        if (m61) SDA_ML(true, DRAW_M61, 20, true);
        else SDA_ML(false, DRAW_GENERIC, 20, true);
    } else if (m61) {
        if (rounds == 8) SDA_ML(true, DRAW_M61, 8, false);
        else if (rounds == 12) SDA_ML(true, DRAW_M61, 12, false);
        else SDA_ML(true, DRAW_M61, 20, false);
    } else {
        if (rounds == 8) SDA_ML(false, DRAW_GENERIC, 8, false);
        else if (rounds == 12) SDA_ML(false, DRAW_GENERIC, 12, false);
        else SDA_ML(false, DRAW_GENERIC, 20, false);
    }
#undef SDA_ML
    ++*lc.nlaunch;
    return cudaGetLastError();
}

static size_t mask_combine_slices(int sm_count, size_t P, size_t dim) {
    const size_t ctas = (((dim + 7) / 8) + CTA - 1) / CTA;
    const size_t want = (size_t)sm_count * 8;
    if (ctas >= want || P < 4) return 1;
    size_t s = (want + ctas - 1) / ctas;
    if (s > P / 2) s = P / 2;
    if (s > 65535) s = 65535;
    return s ? s : 1;
}
size_t chacha_mask_combine_scratch_elems(int sm_count, size_t P, size_t dim) {
    const size_t s = mask_combine_slices(sm_count, P, dim);
    // slice partials, then one ChaChaPre per seed (16-byte aligned: the partial rows are multiples of 4 elements)
    return (s > 1 ? s * ((dim + 3) & ~(size_t)3) : 0) + (P * sizeof(ChaChaPre) + 7) / 8;
}

cudaError_t launch_chacha_mask_combine(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr,
                                       const ChaChaKey *keys, size_t P, size_t dim, int64_t *out, int64_t *scratch,
                                       size_t scratch_elems, unsigned *flag) {
    if (dim == 0) return cudaSuccess;
    const size_t dimp = (dim + 3) & ~(size_t)3;
    size_t slices = mask_combine_slices(lc.sm_count, P, dim);
    if (slices > 1 && (scratch == nullptr || scratch_elems < slices * dimp)) slices = 1;
    const size_t sps = P ? (P + slices - 1) / slices : 1;
    int64_t *dst = slices > 1 ? scratch : out;
    const size_t units = (dim + 7) / 8;
    dim3 grid((unsigned)((units + CTA - 1) / CTA), (unsigned)slices);
    const int lanes = pick_lanes(dst, dimp, 8);
    if (f.kind == FIELD_MERSENNE61) {
        // first-round constants per seed behind the slice partials, when the scratch has room and the counters allow
        const size_t used = slices > 1 ? slices * dimp : 0, need = (P * sizeof(ChaChaPre) + 7) / 8;
        ChaChaPre *pres = nullptr;
        if (scratch != nullptr && scratch_elems >= used + need && P > 0 && (units >> 32) == 0) {
            pres = reinterpret_cast<ChaChaPre *>(scratch + used);
            chacha_prepare_kernel<<<(unsigned)((P + 127) / 128), 128, 0, lc.stream>>>(keys, P, pres);
            ++*lc.nlaunch;
        }
        chacha_mask_combine_kernel<true, DRAW_M61><<<grid, CTA, 0, lc.stream>>>(keys, P, sps, dim, dst, dimp, f, dr,
                                                                                lanes, flag, pres);
    } else
        chacha_mask_combine_kernel<false, DRAW_GENERIC><<<grid, CTA, 0, lc.stream>>>(keys, P, sps, dim, dst, dimp, f,
                                                                                     dr, lanes, flag);
    ++*lc.nlaunch;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || slices == 1) return e;
    return launch_combine(lc, f, scratch, dimp, slices, dim, nullptr, out, nullptr, 0);
}

cudaError_t launch_packed_share(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, int rounds, int k,
                                int t, int n, const Matrix &mtx, const int64_t *secrets, size_t ld, size_t P,
                                size_t dim, const ChaChaKey *keys, const uint64_t *draws, const uint64_t *d_mat,
                                int64_t *shares_out, unsigned *flag) {
    if (dim == 0 || P == 0) return cudaSuccess;
    if (P > 65535) return cudaErrorInvalidValue;
    const int lazy = lazy_terms(f.m);
    if (draws == nullptr && f.kind == FIELD_MERSENNE61 && packed_share_has_fast_path(k, t, n)) {
        return launch_packed_share_m61(lc, rounds, k, t, n, mtx, secrets, ld, P, dim, keys, shares_out, flag);
    }
    if (draws == nullptr) {
#define SDA_CFG(K, T, N)                                                                                     \
    if (k == K && t == T && n == N) {                                                                        \
        *lc.kernel_name = "packed_share<" #K "," #T "," #N ">/generic";                                      \
        return packed_dispatch<K, T, N>(lc, f, dr, rounds, lazy, mtx, secrets, ld, P, dim, keys, shares_out, \
                                        flag);                                                               \
    }
        SDA_CFG(3, 2, 5)   // BASELINE config #3
        SDA_CFG(5, 4, 9)   // BASELINE config #4
        SDA_CFG(3, 4, 7)   // BASELINE config #5
        SDA_CFG(3, 4, 8)   // the reference's own test parameters (full_loop.rs:57-64)
#undef SDA_CFG
        return cudaErrorInvalidValue;   // caller must pre-draw for other shapes
    }
    *lc.kernel_name = "packed_share<draws from memory>";
    if (d_mat == nullptr) return cudaErrorInvalidValue;
    const size_t B = (dim + k - 1) / k;
    dim3 grid((unsigned)((B + CTA - 1) / CTA), (unsigned)P);
    const uint64_t *mat_g = d_mat;
    if (f.kind == FIELD_MERSENNE61)
        packed_share_mem_kernel<true><<<grid, CTA, 0, lc.stream>>>(secrets, ld, dim, B, k, t, n, mat_g, draws,
                                                                   shares_out, f, lazy);
    else
        packed_share_mem_kernel<false><<<grid, CTA, 0, lc.stream>>>(secrets, ld, dim, B, k, t, n, mat_g, draws,
                                                                    shares_out, f, lazy);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

}  // namespace sda
