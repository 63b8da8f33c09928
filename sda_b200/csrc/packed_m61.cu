// packed_m61.cu -- K2 for the Mersenne prime p = 2^61 - 1: packed-Shamir share generation
//   client/src/crypto/sharing/packed_shamir.rs:40-43 -> tss 0.2 `share`  (+ batched.rs:18-53)
// as shares_b = M . [secrets_b ; randomness_b] per batch b, with every u64 of randomness taken
// from the participant's ChaCha keystream at its rand-0.3 stream position.
//
// The kernel is bound by the SM's two integer pipes (ALU: LOP3/SHF/IADD3, FMA: IMAD), not by
// HBM, so it is organised to keep the instruction count per batch minimal and the hot code
// small enough for the instruction caches (profiles/r01_k2_baseline.md):
//
//   * a thread computes whole 64-byte keystream blocks (rolled round loop) and parks them in
//     shared memory; after a warp barrier the warp walks its 32*G batches G times, lane = batch,
//     so secrets are read as 8-byte words at lane stride 8K bytes and every share row is written
//     as 32 consecutive i64 -- coalesced without any vector-alignment variants;
//   * field elements are split into centred 31/30-bit limbs x = x0 + x1 2^31 + (2^30 + 2^60),
//     x0 in [-2^30, 2^30), |x1| <= 2^29, the matrix likewise (c = m0 + m1 2^31, d1 = 2 m1), and
//     one dot product is 4 IMAD.WIDE per term into two 64-bit accumulators
//         A = sum m0 x0 + d1 x1  (signed)      X = 2^63 + K'_row + sum m0 x1 + m1 x0  (unsigned)
//     (2^62 == 2 puts the top product in A).  Rows longer than 5 terms are renormalised
//     (A <- (A & p) + (A >> 61) + offset, same for X) before each further chunk of 4 terms.  The
//     constant K'_row (host, 128-bit) cancels every offset and adds 1;
//   * the fold  t = A + X 2^31 (mod p)  is 2 ALU + 3 IMAD.WIDE, and because t == result + 1 the
//     canonical value is (t + q - 1) & p with q = floor(t / p) = (t + (t >> 61)) >> 61:
//     3 ALU + 2 IMAD.WIDE, no compare/select chains;
//   * a draw is reduced to x == v mod (p - 1) without comparisons as (v & p) + 2 (v >> 61); the
//     one-in-2^57 words for which that differs from gen_range's answer (including the rejected
//     ones) raise `flag`, and the host redoes the call on the exact path like every other kernel.
#include "kernels.h"

namespace sda {

namespace {

constexpr int CTA = 128;
constexpr int WARPS = CTA / 32;

constexpr int gcd_c(int a, int b) { return b == 0 ? a : gcd_c(b, a % b); }
template <int T>
struct Unit {
    static constexpr int G = 8 / gcd_c(T, 8);    // passes over a warp's batches
    static constexpr int NB = T * G / 8;         // keystream blocks per thread
};

template <int N, int W>
struct M61Params {
    int32_t m0[N * W];
    int32_t m1[N * W];
    int32_t d1[N * W];
    uint64_t x_init[N];              // OX1 + K'_row
    uint32_t one, four, two31;       // multipliers ptxas must not strength-reduce into ALU adds
    uint32_t zero;                   // see SDA_QR
};

constexpr uint64_t OX1 = 1ull << 63;
constexpr uint32_t OA2_HI = 0xf0000000u;    // A is signed: renormalise around 0 (offset -2^60)
constexpr uint32_t OX2_HI = 0x60000000u;    // 6 * 2^60 (disjoint from the 29 kept bits: OR-able)
constexpr uint32_t LOW29 = 0x1fffffffu;

// c + a * b as mul.wide + add: the form ptxas keeps as one IMAD.WIDE per term with the running sum
// as addend (a chain of mad.wide is re-associated into IMAD.WIDE + 3-input IADD3 trees, which
// moves the sums onto the ALU pipe this kernel is bound by)
__device__ __forceinline__ uint64_t mac_s(int32_t a, int32_t b, uint64_t c) {
    uint64_t d;
    asm("{\n\t.reg .s64 t;\n\tmul.wide.s32 t, %1, %2;\n\tadd.s64 %0, %3, t;\n\t}" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mac_u(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t d;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %1, %2;\n\tadd.u64 %0, %3, t;\n\t}" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
__device__ __forceinline__ void unpack(uint64_t v, uint32_t &lo, uint32_t &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}

struct Consts {
    uint32_t one, four, two31;
};

// A (signed) -> A mod p in [-2^60 - 4, 2^60 + 3]
__device__ __forceinline__ uint64_t renorm_a(uint64_t a, const Consts &c) {
    uint32_t lo, hi;
    unpack(a, lo, hi);
    return mac_s((int32_t)hi >> 29, (int32_t)c.one, pack(lo, (hi & LOW29) + OA2_HI));
}
// X (unsigned) -> X mod p in [6 2^60, 8 2^60 + 7]
__device__ __forceinline__ uint64_t renorm_x(uint64_t x, const Consts &c) {
    uint32_t lo, hi;
    unpack(x, lo, hi);
    return mac_u(hi >> 29, c.one, pack(lo, (hi & LOW29) | OX2_HI));
}

// canonical (A + X 2^31 - 1) mod p for signed A, unsigned X
__device__ __forceinline__ uint64_t fold(uint64_t A, uint64_t X, const Consts &c) {
    uint32_t a_lo, a_hi, x_lo, x_hi;
    unpack(A, a_lo, a_hi);
    unpack(X, x_lo, x_hi);
    uint64_t t = mac_u(x_lo, c.two31, pack(a_lo, a_hi & LOW29));        // X_lo 2^31
    t = mac_u(x_hi, c.four, t);                                         // X_hi 2^63 == 4 X_hi
    t = mac_s((int32_t)a_hi >> 29, (int32_t)c.one, t);                  // A[61..63] 2^61 == itself
    uint32_t t_lo, t_hi, s_lo, s_hi;
    unpack(t, t_lo, t_hi);
    unpack(mac_u(t_hi >> 29, c.one, t), s_lo, s_hi);
    const int32_t qm1 = (int32_t)(s_hi >> 29) - 1;                      // floor(t / p) - 1
    uint32_t r_lo, r_hi;
    unpack(mac_s(qm1, (int32_t)c.one, t), r_lo, r_hi);
    return pack(r_lo, r_hi & LOW29);
}

__device__ __forceinline__ uint64_t canon_slow(int64_t v) {             // any i64 -> [0, p)
    uint64_t a = v < 0 ? 0ull - (uint64_t)v : (uint64_t)v;
    uint64_t r = (a & P61) + (a >> 61);
    r = r >= P61 ? r - P61 : r;
    return (v < 0 && r) ? P61 - r : r;
}

// One quarter round.  ptxas puts plain 32-bit adds on the FMA pipe (IMAD.IADD) because the round loop
// by itself is ALU-heavy; in this kernel the FMA pipe is the busier one (IMAD.WIDE occupies it for
// 4 cycles per warp), so `z` -- a zero the compiler cannot see through -- turns chosen adds into
// 3-input IADD3, which only the ALU pipe executes.
#define SDA_QR(a, b, c, d)                                          \
    a += b; d ^= a; d = __funnelshift_l(d, d, 16);                  \
    c = add3(c, d, z); b ^= c; b = __funnelshift_l(b, b, 12);       \
    a = add3(a, b, z); d ^= a; d = __funnelshift_l(d, d, 8);        \
    c = add3(c, d, z); b ^= c; b = __funnelshift_l(b, b, 7);

__device__ __forceinline__ uint32_t add3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("{\n\t.reg .u32 t;\n\tadd.u32 t, %1, %2;\n\tadd.u32 %0, t, %3;\n\t}" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int ROUNDS>
__device__ __forceinline__ void chacha_block_to_smem(const uint32_t (&k)[8], uint64_t block, uint32_t z, uint4 *dst) {
    const uint32_t c0 = 0x61707865u, c1 = 0x3320646eu, c2 = 0x79622d32u, c3 = 0x6b206574u;
    const uint32_t b0 = (uint32_t)block, b1 = (uint32_t)(block >> 32);
    uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3;
    uint32_t x4 = k[0], x5 = k[1], x6 = k[2], x7 = k[3];
    uint32_t x8 = k[4], x9 = k[5], x10 = k[6], x11 = k[7];
    uint32_t x12 = b0, x13 = b1, x14 = 0, x15 = 0;
#pragma unroll 1
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
    dst[0] = make_uint4(x0 + c0, x1 + c1, x2 + c2, x3 + c3);
    dst[1] = make_uint4(x4 + k[0], x5 + k[1], x6 + k[2], x7 + k[3]);
    dst[2] = make_uint4(x8 + k[4], x9 + k[5], x10 + k[6], x11 + k[7]);
    dst[3] = make_uint4(x12 + b0, x13 + b1, x14, x15);
}

template <int K, int T, int N, int ROUNDS>
__global__ void __launch_bounds__(CTA)
packed_share_m61_kernel(const int64_t *__restrict__ secrets, size_t ld, size_t dim, size_t B,
                        const ChaChaKey *__restrict__ keys, int64_t *__restrict__ out,
                        const __grid_constant__ M61Params<N, K + T> prm, unsigned *flag) {
    constexpr int W = K + T;
    constexpr int G = Unit<T>::G, NB = Unit<T>::NB;
    constexpr int UB = 32 * G;                  // batches per warp
    __shared__ __align__(16) uint4 stage[WARPS][NB * 32 * 4];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t p = blockIdx.y;
    const size_t unit = (size_t)blockIdx.x * WARPS + warp;
    const size_t b_base = unit * UB;
    if (b_base >= B) return;                    // warp-uniform

    // ---- randomness: NB blocks per lane into stream order ---------------------------------
    {
        uint32_t k[8];
        const uint4 *src = reinterpret_cast<const uint4 *>(keys + p);
        const uint4 ka = __ldg(src), kb = __ldg(src + 1);
        k[0] = ka.x; k[1] = ka.y; k[2] = ka.z; k[3] = ka.w;
        k[4] = kb.x; k[5] = kb.y; k[6] = kb.z; k[7] = kb.w;
#pragma unroll 1
        for (int nb = 0; nb < NB; nb++) {
            const int slot = nb * 32 + lane;
            chacha_block_to_smem<ROUNDS>(k, unit * (32 * NB) + slot, prm.zero, &stage[warp][slot * 4]);
        }
    }
    __syncwarp();

    const Consts cs{prm.one, prm.four, prm.two31};
    uint64_t xi[N];
#pragma unroll
    for (int r = 0; r < N; r++) xi[r] = prm.x_init[r];
    const int64_t *sec = secrets + p * ld;
    int64_t *o = out + p * (size_t)N * B;
    const uint2 *draws = reinterpret_cast<const uint2 *>(&stage[warp][0]);

#pragma unroll 1
    for (int g = 0; g < G; g++) {
        const int j = g * 32 + lane;
        const size_t b = b_base + j;
        if (b >= B) break;
        int32_t x0[W], x1[W];

        // secrets of batch b; the last batch is zero padded (batched.rs:38-43)
        {
            int64_t s[K];
            const size_t e0 = b * K;
            uint32_t top = 0;
#pragma unroll
            for (int i = 0; i < K; i++) {
                s[i] = e0 + i < dim ? __ldg(sec + e0 + i) : 0;
                top |= (uint32_t)((uint64_t)s[i] >> 32);
            }
            if (top >> 29) {                    // some value outside [0, 2^61): any i64 is legal input
#pragma unroll
                for (int i = 0; i < K; i++) s[i] = (int64_t)canon_slow(s[i]);
            }
#pragma unroll
            for (int i = 0; i < K; i++) {
                uint32_t lo, hi;
                unpack((uint64_t)s[i], lo, hi);
                x0[i] = (int32_t)(lo & 0x7fffffffu) - (1 << 30);
                x1[i] = (int32_t)__funnelshift_l(lo, hi, 1) - (1 << 29);
            }
        }
        // T draws of batch b: stream positions b*T ..; word pair (hi, lo) per draw
#pragma unroll
        for (int i = 0; i < T; i++) {
            const uint2 d = draws[j * T + i];
            const uint32_t w0 = d.x, w1 = d.y;
            const uint32_t l1 = __funnelshift_l(w1, w0, 1) & 0x3fffffffu;
            const uint32_t h = w0 >> 29;
            if (l1 == 0x3fffffffu) {
                // v mod 2^61 >= 2^61 - 2^31: tss's Range::new(0, p - 1) may reject the word or wrap it
                const uint64_t v = ((uint64_t)w0 << 32) | w1;
                if ((v & P61) + 2 * (v >> 61) >= P61 - 1 || v >= 0xfffffffffffffff0ull) atomicOr(flag, 1u);
            }
            x0[K + i] = (int32_t)((w1 & 0x7fffffffu) + 2 * h) - (1 << 30);
            x1[K + i] = (int32_t)l1 - (1 << 29);
        }

        int64_t *ob = o + b;
#pragma unroll
        for (int r = 0; r < N; r++) {
            uint64_t A = 0, X = xi[r];
#pragma unroll
            for (int i = 0; i < W; i++) {
                if (i >= 5 && (i - 5) % 4 == 0) {
                    A = renorm_a(A, cs);
                    X = renorm_x(X, cs);
                }
                const int32_t m0 = prm.m0[r * W + i], m1 = prm.m1[r * W + i], d1 = prm.d1[r * W + i];
                A = mac_s(m0, x0[i], A);
                A = mac_s(d1, x1[i], A);
                X = mac_s(m0, x1[i], X);
                X = mac_s(m1, x0[i], X);
            }
            ob[(size_t)r * B] = (int64_t)fold(A, X, cs);
        }
    }
}

template <int N, int W>
M61Params<N, W> make_params(const Matrix &m) {
    typedef unsigned __int128 u128;
    M61Params<N, W> s;
    int nb = 0;                                  // renormalisations per row
    for (int i = 5; i < W; i += 4) nb++;
    const u128 delta = ((u128)1 << 30) + ((u128)1 << 60);
    // everything the accumulators carry besides sum c_i v_i, as a residue mod p
    const u128 oa2 = P61 - (((u128)1 << 60) % P61);          // -2^60
    const u128 offs = ((u128)nb * oa2 +
                       (((u128)OX1 % P61 + (u128)nb * (((u128)OX2_HI << 32) % P61)) % P61) * (((u128)1 << 31) % P61)) % P61;
    for (int r = 0; r < N; r++) {
        u128 cd = 0;                             // sum c_i * delta mod p, with c_i the centred entry
        for (int i = 0; i < W; i++) {
            const uint64_t e = m.e[r * W + i];
            const int64_t c = e > P61 / 2 ? (int64_t)e - (int64_t)P61 : (int64_t)e;      // (-2^60, 2^60)
            const int64_t m1 = (c + (1ll << 30)) >> 31;
            s.m0[r * W + i] = (int32_t)(c - m1 * (1ll << 31));                           // [-2^30, 2^30)
            s.m1[r * W + i] = (int32_t)m1;                                               // |m1| <= 2^29
            s.d1[r * W + i] = (int32_t)(2 * m1);
            cd = (cd + (u128)e * (delta % P61)) % P61;
        }
        // K' 2^31 == sum c_i delta + 1 - offs   ->   K' = (...) 2^30   (2^61 == 1)
        const u128 k = (cd + 1 + P61 - offs) % P61;
        const u128 kp = k * (((u128)1 << 30) % P61) % P61;
        s.x_init[r] = OX1 + (uint64_t)kp;
    }
    s.one = 1u;
    s.four = 4u;
    s.two31 = 0x80000000u;
    s.zero = 0u;
    return s;
}

template <int K, int T, int N, int ROUNDS>
cudaError_t launch(const LaunchCtx &lc, const Matrix &mtx, const int64_t *secrets, size_t ld, size_t P, size_t dim,
                   const ChaChaKey *keys, int64_t *out, unsigned *flag) {
    constexpr int G = Unit<T>::G;
    const size_t B = (dim + K - 1) / K;
    const size_t per_cta = (size_t)WARPS * 32 * G;
    dim3 grid((unsigned)((B + per_cta - 1) / per_cta), (unsigned)P);
    packed_share_m61_kernel<K, T, N, ROUNDS><<<grid, CTA, 0, lc.stream>>>(secrets, ld, dim, B, keys, out,
                                                                         make_params<N, K + T>(mtx), flag);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

template <int K, int T, int N>
cudaError_t dispatch(const LaunchCtx &lc, int rounds, const Matrix &mtx, const int64_t *secrets, size_t ld, size_t P,
                     size_t dim, const ChaChaKey *keys, int64_t *out, unsigned *flag) {
    if (rounds == 8) return launch<K, T, N, 8>(lc, mtx, secrets, ld, P, dim, keys, out, flag);
    if (rounds == 12) return launch<K, T, N, 12>(lc, mtx, secrets, ld, P, dim, keys, out, flag);
    return launch<K, T, N, 20>(lc, mtx, secrets, ld, P, dim, keys, out, flag);
}

}  // namespace

// in-kernel-rng share generation over 2^61 - 1 for the shapes packed_share_has_fast_path() lists
cudaError_t launch_packed_share_m61(const LaunchCtx &lc, int rounds, int k, int t, int n, const Matrix &mtx,
                                    const int64_t *secrets, size_t ld, size_t P, size_t dim, const ChaChaKey *keys,
                                    int64_t *shares_out, unsigned *flag) {
#define SDA_CFG(K, T, N)                                                                          \
    if (k == K && t == T && n == N) {                                                             \
        *lc.kernel_name = "packed_share<" #K "," #T "," #N ">/mersenne61 warp-staged";           \
        return dispatch<K, T, N>(lc, rounds, mtx, secrets, ld, P, dim, keys, shares_out, flag); \
    }
    SDA_CFG(3, 2, 5)   // BASELINE config #3
    SDA_CFG(5, 4, 9)   // BASELINE config #4
    SDA_CFG(3, 4, 7)   // BASELINE config #5
    SDA_CFG(3, 4, 8)   // the reference's own test shape (full_loop.rs:57-64)
#undef SDA_CFG
    return cudaErrorInvalidValue;
}

}  // namespace sda
