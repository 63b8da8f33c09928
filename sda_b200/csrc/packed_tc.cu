// packed_tc.cu -- K2 for p = 2^61 - 1 on the 5th-generation tensor cores: packed-Shamir share
// generation (client/src/crypto/sharing/packed_shamir.rs:40-43 -> tss 0.2 `share`, + batched.rs:18-53)
// cast as a dense integer matrix product.
//
//   shares_b = M . x_b  (mod p),  x_b = [secrets_b ; randomness_b]  for every batch b
//
// is linear in the BYTES of x_b:  share_j = sum_{i,c} byte_c(x_i) * (M[j][i] 2^{8c} mod p), and each
// 61-bit constant is itself 8 bytes, so with
//     A[b][(i,c)]      = byte c of x_i of batch b                      (u8, one 64..96-byte row per batch)
//     B[(j,s)][(i,c)]  = byte s of (M[j][i] 2^{8c} mod p)              (u8, constant)
// the s32 tile  D = A . B^T  holds, for batch b and share j, eight limb sums D[b][(j,s)] < 2^23 with
//     share_j == sum_s D[b][(j,s)] 2^{8s}   (mod p).
// One tcgen05.mma.kind::i8 (M = 128 batches, N = 8 * share_count rounded up to 16, K = 32 bytes) per
// 32 bytes of row computes D into TMEM; a thread then owns one batch (TMEM lane), reads its limb
// sums with tcgen05.ld and only has to carry-propagate and reduce: ~16 integer instructions per share
// instead of the 4 IMAD.WIDE per matrix entry + fold of the CUDA-core kernel (packed_m61.cu), whose
// IMAD.WIDE stream is what bounds it (profiles/r01_k2_cuda.md).
//
// The A rows are nothing but the operands as they lie in memory: a secret's 8 little-endian bytes
// ARE its byte limbs, so the secrets go from global memory to the shared-memory tile unchanged, and
// a draw is stored as (v & p) + 2 (v >> 61) == v mod (p - 1) (tss draws from [0, p - 1)).  Negative
// secrets (legal i64 inputs) are canonicalised on the way; any u64 bit pattern is a valid row.
//
// Shared-memory operand layout (no swizzle, K-major, as the UMMA descriptor defines it): 8 rows x
// 16 bytes form a contiguous 128-byte core matrix; LBO = 128 bytes steps to the next 16-byte K chunk,
// SBO steps to the next 8 rows.  tools/tc_probe.cu pins the descriptor encoding against a host GEMM.
//
// Randomness: every u64 comes from the participant's ChaCha keystream at its rand-0.3 stream
// position exactly as in packed_m61.cu; a thread computes whole 64-byte blocks and scatters the
// reduced draws into the rows they belong to.
//
// A pass of a persistent CTA (4 per SM) = G tiles of 128 batches of one participant:
//   * its raw secrets (G x 128 x K x 8 bytes, contiguous in the vector) were brought into shared memory by one
//     cp.async.bulk on the TMA unit, issued a pass earlier by one thread and completed on an mbarrier; each thread
//     moves its row into the operand layout as 16-byte chunks (per-thread loads for ragged or unaligned passes);
//   * its draws were staged, double-buffered, while the previous pass's MMAs ran;
//   * the tiles go through the tensor core two at a time (two TMEM accumulators): one `full` wait, two
//     tcgen05.ld groups, one `drained` arrive and one MMA issue per pair, the next pair's MMAs running under the
//     compose arithmetic and the stores of the current one.
// The kernel is bound by instruction issue (ChaCha20's 12 instructions per quarter round are 55 % of it), not by HBM,
// latency or the MMA; profiles/r01_k2_variants.md records what was tried.
#include <algorithm>
#include <cstring>

#include "chacha_pre.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace sda {

namespace {

using namespace tc;

#ifndef SDA_TC_PREFETCH
#define SDA_TC_PREFETCH 2   // where the next pass's secrets are loaded: 1 under the last tile's compose, 2 before the tile loop
#endif

#ifndef SDA_TC_ACC_BUFS
#define SDA_TC_ACC_BUFS 2   // TMEM accumulators: 2 = the tiles of a pass go through the tensor core and the barriers in pairs
                            // (16.54 ms against 17.18 ms with 1 on the same box, profiles/r01_k2_variants.md)
#endif
#ifndef SDA_TC_BULK_IN
#define SDA_TC_BULK_IN 1    // secrets of a pass arrive in shared memory by one bulk copy (TMA unit) instead of 12 loads per thread
#endif
#ifndef SDA_TC_MINBLOCKS
#define SDA_TC_MINBLOCKS 1
#endif
#ifndef SDA_TC_UNROLL
#define SDA_TC_UNROLL 2
#endif

constexpr int kChaChaUnroll = SDA_TC_UNROLL;   // double rounds per iteration of the (otherwise rolled) keystream loop
constexpr int CTA = 128;             // threads = rows of one MMA tile = TMEM lanes

constexpr int gcd_c(int a, int b) { return b == 0 ? a : gcd_c(b, a % b); }

template <int K, int T, int N>
struct Shape {
    static_assert(T % 2 == 0 && T >= 2, "draws fill whole 16-byte chunks");
    static constexpr int kT = T;
    static constexpr int G = 8 / gcd_c(T, 8);              // 128-batch tiles per pass of a CTA
    static constexpr int NB = T / gcd_c(T, 8);             // keystream blocks per thread per pass
    // A row = [draws | secrets]; the two parts live in separate buffers (the draws double-buffered, so the
    // keystream of the next pass is computed under this pass's MMAs) and are consumed by separate K steps
    static constexpr int DC = T / 2;                       // 16-byte chunks of a row holding draws
    static constexpr int SC = (K + 1) / 2;                 // ... holding secrets
    static constexpr int NKD = (DC + 1) / 2, NKS = (SC + 1) / 2;
    static constexpr int NK = NKD + NKS;                   // MMAs per tile (32 bytes of K each)
    // an odd chunk count lets the second chunk of the last K step alias the next 8-row group: B is 0 there
    static constexpr uint32_t SBO_D = DC * 128, SBO_S = SC * 128;
    static constexpr uint32_t D_TILE = 16 * SBO_D, S_TILE = 16 * SBO_S;
    static constexpr uint32_t D_BYTES = G * D_TILE + 128, S_BYTES = G * S_TILE + 128;
    static constexpr int NMMA = (8 * N + 15) / 16 * 16;
    static constexpr uint32_t SBO_B = 2 * NK * 128;
    static constexpr uint32_t B_BYTES = NMMA / 8 * SBO_B;
    // the raw secrets of one pass, [G x 128 batches][K] i64 exactly as they lie in the participant's vector
    static constexpr uint32_t IN_BYTES = SDA_TC_BULK_IN ? G * 128 * K * 8 : 0;
    static constexpr uint32_t SMEM = 2 * D_BYTES + S_BYTES + B_BYTES + IN_BYTES;
    static constexpr int ACC_COLS = NMMA <= 32 ? 32 : NMMA <= 64 ? 64 : NMMA <= 128 ? 128 : 256;
    // optionally two TMEM accumulators (when that still leaves four CTAs per SM their columns): the tiles of a pass
    // are multiplied, waited for and drained two at a time, which halves the barrier traffic per tile
    static constexpr int ACC_BUFS = (SDA_TC_ACC_BUFS >= 2 && 2 * ACC_COLS <= 128 && G % 2 == 0) ? 2 : 1;
    static constexpr int TMEM_COLS = ACC_BUFS * ACC_COLS;
    static constexpr uint32_t IDESC = idesc_u8(NMMA);
    static_assert(4 % DC == 0, "a keystream block covers whole rows");
    static_assert(D_BYTES % 128 == 0 && S_BYTES % 128 == 0, "operand buffers stay 128-byte aligned");
    static_assert(B_BYTES % 16 == 0 && IN_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes to 16-byte aligned addresses");
};

#define SDA_QR(a, b, c, d)                                      \
    a += b; d ^= a; d = __funnelshift_l(d, d, 16);              \
    c += d; b ^= c; b = __funnelshift_l(b, b, 12);              \
    a += b; d ^= a; d = __funnelshift_l(d, d, 8);               \
    c += d; b ^= c; b = __funnelshift_l(b, b, 7);

template <int ROUNDS>
__device__ __forceinline__ void chacha_block(const uint32_t (&k)[8], uint64_t block, uint32_t (&o)[16]) {
    const uint32_t c0 = 0x61707865u, c1 = 0x3320646eu, c2 = 0x79622d32u, c3 = 0x6b206574u;
    const uint32_t b0 = (uint32_t)block, b1 = (uint32_t)(block >> 32);
    uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3;
    uint32_t x4 = k[0], x5 = k[1], x6 = k[2], x7 = k[3];
    uint32_t x8 = k[4], x9 = k[5], x10 = k[6], x11 = k[7];
    uint32_t x12 = b0, x13 = b1, x14 = 0, x15 = 0;
#pragma unroll kChaChaUnroll
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
    o[0] = x0 + c0;     o[1] = x1 + c1;     o[2] = x2 + c2;      o[3] = x3 + c3;
    o[4] = x4 + k[0];   o[5] = x5 + k[1];   o[6] = x6 + k[2];    o[7] = x7 + k[3];
    o[8] = x8 + k[4];   o[9] = x9 + k[5];   o[10] = x10 + k[6];  o[11] = x11 + k[7];
    o[12] = x12 + b0;   o[13] = x13 + b1;   o[14] = x14;         o[15] = x15;
}

// draw (hi word w0, lo word w1) -> v mod (p - 1) as (v & p) + 2 (v >> 61).  That is gen_range's answer
// unless v mod 2^61 >= 2^61 - 32 (a rejected word or a wrap-around).  The necessary condition "bits 32..60
// all ones" (2^-29 per draw) is accumulated as the running maximum of the masked high words (one three-input
// VIMNMX per two draws), and the caller settles the rare case `suspect >= 2^29 - 1` exactly.
__device__ __forceinline__ uint64_t reduce_draw(uint32_t w0, uint32_t w1, uint32_t &suspect) {
    uint32_t h;
    asm("shr.u32 %0, %1, 29;" : "=r"(h) : "r"(w0));       // kept apart from the doubling: (h << 1) + w1 is one LEA with carry
    const uint32_t hi = w0 & LOW29;
    suspect = max(suspect, hi);
    return pack(w1, hi) + (uint64_t)(h << 1);
}

// ---- any prime below 2^63 -----------------------------------------------------------------------------
// The byte-limb GEMM does not care about the modulus (the constants are reduced on the host); only the two
// scalar steps around it do: a draw is v mod (p - 1) by reciprocal multiplication with gen_range's rejection
// zone tested exactly, and the limb sums are composed into a 79-bit integer and reduced once (field.cuh).
struct GenericField {
    FieldParams f;       // p
    DrawParams dr;       // range p - 1
};

__device__ __forceinline__ uint64_t reduce_draw_generic(const DrawParams &dr, uint32_t w0, uint32_t w1, uint32_t &suspect) {
    const uint64_t v = pack(w1, w0);
    if (v >= dr.zone) suspect = 0xffffffffu;                // rejected by gen_range: the stream shifts, host redoes the call
    return reduce64_generic(dr.f, v);
}

__device__ __forceinline__ uint64_t compose_generic(const uint32_t (&d)[8], uint32_t two16, const FieldParams &f) {
    const uint32_t e0 = d[0] + (d[1] << 8), e1 = d[2] + (d[3] << 8);
    const uint32_t e2 = d[4] + (d[5] << 8), e3 = d[6] + (d[7] << 8);
    const uint64_t L = mac_u(e1, two16, e0), H = mac_u(e3, two16, e2);           // value = L + H 2^32 < 2^80
    const uint64_t lo = L + (H << 32);
    uint64_t hi = (H >> 32) + (lo < L);                                          // < 2^17
    if (f.m <= (1ull << 17)) hi = reduce64_generic(f, hi);                       // reduce128 needs hi < m (uniform branch)
    return reduce128_generic(f, hi, lo);
}

static __device__ __noinline__ uint64_t canon_negative_generic(const FieldParams &f, int64_t v) {
    return canon<false>(f, v);
}

// one thread: the NK MMAs of a 128-row tile, completion signalled on `full_bar`
template <class S>
__device__ __forceinline__ void issue_tile(uint32_t taddr, uint32_t d_tile, uint32_t s_tile, uint32_t b_base, uint32_t full_bar,
                                           bool commit = true) {
    const uint64_t dd = umma_desc(d_tile, S::SBO_D), ds = umma_desc(s_tile, S::SBO_S), db = umma_desc(b_base, S::SBO_B);
#pragma unroll
    for (int kk = 0; kk < S::NKD; kk++)
        umma_i8(taddr, dd + ((2 * LBO * kk) >> 4), db + ((2 * LBO * kk) >> 4), S::IDESC, kk > 0);
#pragma unroll
    for (int kk = 0; kk < S::NKS; kk++)
        umma_i8(taddr, ds + ((2 * LBO * kk) >> 4), db + ((2 * LBO * (S::NKD + kk)) >> 4), S::IDESC, 1);
    if (commit) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(full_bar) : "memory");
}

// the tiles q .. q + ACC_BUFS - 1 of a pass into the ACC_BUFS accumulators, one completion on `full_bar`
template <class S>
__device__ __forceinline__ void issue_tiles(uint32_t taddr, uint32_t d_cur, uint32_t s_base, uint32_t b_base, uint32_t full_bar, int q) {
#pragma unroll
    for (int a = 0; a < S::ACC_BUFS; a++)
        issue_tile<S>(taddr + a * S::ACC_COLS, d_cur + (q + a) * S::D_TILE, s_base + (q + a) * S::S_TILE, b_base, full_bar,
                      a == S::ACC_BUFS - 1);
}

// the keystream of one pass: NB blocks per thread, every draw reduced and scattered into the row it belongs to
template <class S, int ROUNDS, bool M61>
__device__ __forceinline__ void stage_draws(const ChaChaKey *__restrict__ keys, size_t p, size_t u, int tid, uint8_t *sD,
                                            const GenericField &gf, unsigned *flag) {
    uint32_t k[8];
    const uint4 *src = reinterpret_cast<const uint4 *>(keys + p);
    const uint4 ka = __ldg(src), kb = __ldg(src + 1);
    k[0] = ka.x; k[1] = ka.y; k[2] = ka.z; k[3] = ka.w;
    k[4] = kb.x; k[5] = kb.y; k[6] = kb.z; k[7] = kb.w;
#pragma unroll 1
    for (int nb = 0; nb < S::NB; nb++) {
        const uint32_t slot = nb * CTA + tid;                     // block of this pass, in stream order
        uint32_t w[16];
        chacha_block<ROUNDS>(k, u * (size_t)(CTA * S::NB) + slot, w);
        uint32_t suspect = 0;
#pragma unroll
        for (int cb = 0; cb < 4; cb++) {                          // 4 chunks of 2 draws
            const uint64_t xa = M61 ? reduce_draw(w[4 * cb], w[4 * cb + 1], suspect)
                                    : reduce_draw_generic(gf.dr, w[4 * cb], w[4 * cb + 1], suspect);
            const uint64_t xb = M61 ? reduce_draw(w[4 * cb + 2], w[4 * cb + 3], suspect)
                                    : reduce_draw_generic(gf.dr, w[4 * cb + 2], w[4 * cb + 3], suspect);
            const uint32_t gc = slot * 4 + cb;                    // chunk index of the pass
            const uint32_t batch = gc / S::DC, c = gc % S::DC;
            const uint32_t q = batch / CTA, row = batch % CTA;
            uint32_t xal, xah, xbl, xbh;
            unpack(xa, xal, xah);
            unpack(xb, xbl, xbh);
            *reinterpret_cast<uint4 *>(sD + q * S::D_TILE + (row >> 3) * S::SBO_D + c * LBO + (row & 7) * 16) =
                make_uint4(xal, xah, xbl, xbh);
        }
        if (suspect >= LOW29) {
            bool bad = !M61;
#pragma unroll
            for (int d = 0; d < 8; d++) bad |= (w[2 * d] & LOW29) == LOW29 && w[2 * d + 1] >= 0xffffffe0u;
            if (bad) atomicOr(flag, 1u);
        }
    }
}

// the K secrets of row `tid` of each of the G tiles of pass (p, u) into registers; the last batch of a
// vector is zero padded (batched.rs:38-43)
template <class S, int K>
__device__ __forceinline__ void load_secrets(const int64_t *__restrict__ secrets, size_t ld, size_t dim, size_t p, size_t u,
                                             int tid, uint4 (&s)[S::G][S::SC]) {
    // a row's secrets as the 16-byte chunks they are staged as: (x, y) = the even secret of the chunk, (z, w) = the odd
    const int64_t *sec = secrets + p * ld;
    const size_t b0 = u * (size_t)(S::G * CTA);
    const size_t e_first = (b0 + tid) * K;
    if ((b0 + (size_t)S::G * CTA) * K <= dim) {                                       // whole pass inside the vector
        const uint2 *src = reinterpret_cast<const uint2 *>(sec + e_first);
#pragma unroll
        for (int q = 0; q < S::G; q++)
#pragma unroll
            for (int c = 0; c < S::SC; c++) {
                const uint2 a = __ldg(src + q * (CTA * K) + 2 * c);
                const uint2 b = 2 * c + 1 < K ? __ldg(src + q * (CTA * K) + 2 * c + 1) : make_uint2(0, 0);
                s[q][c] = make_uint4(a.x, a.y, b.x, b.y);
            }
    } else {
#pragma unroll
        for (int q = 0; q < S::G; q++) {
            const size_t e0 = e_first + (size_t)q * (CTA * K);
#pragma unroll
            for (int c = 0; c < S::SC; c++) {
                uint2 a = make_uint2(0, 0), b = make_uint2(0, 0);
                if (e0 + 2 * c < dim) a = __ldg(reinterpret_cast<const uint2 *>(sec + e0 + 2 * c));
                if (2 * c + 1 < K && e0 + 2 * c + 1 < dim) b = __ldg(reinterpret_cast<const uint2 *>(sec + e0 + 2 * c + 1));
                s[q][c] = make_uint4(a.x, a.y, b.x, b.y);
            }
        }
    }
}

// The same rows into the pass's shared-memory staging buffer, by the threads themselves: the path of a pass that
// is not wholly inside the vector (zero padding, batched.rs:38-43) or whose source is not 16-byte aligned.  Every
// thread writes, and later reads, only its own K words per tile.
template <class S, int K>
__device__ __forceinline__ void fill_secrets(const int64_t *__restrict__ secrets, size_t ld, size_t dim, size_t p, size_t u,
                                             int tid, int64_t *sIn) {
    const int64_t *sec = secrets + p * ld;
    const size_t e_first = (u * (size_t)(S::G * CTA) + tid) * K;
#pragma unroll
    for (int q = 0; q < S::G; q++) {
        const size_t e0 = e_first + (size_t)q * (CTA * K);
#pragma unroll
        for (int i = 0; i < K; i++) sIn[(q * CTA + tid) * K + i] = e0 + i < dim ? __ldg(sec + e0 + i) : 0;
    }
}

// One thread: the whole pass -- G x 128 batches x K secrets, contiguous in the participant's vector -- into the
// staging buffer with one bulk copy; completion (by byte count) on `bar`.
template <class S, int K>
__device__ __forceinline__ void bulk_load_secrets(const int64_t *__restrict__ secrets, size_t ld, size_t p, size_t u,
                                                  uint32_t sin_addr, uint32_t bar) {
    const int64_t *src = secrets + p * ld + u * (size_t)(S::G * CTA) * K;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(S::IN_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(sin_addr), "l"(src), "r"(S::IN_BYTES), "r"(bar) : "memory");
}

// a negative secret (sign bit in its high word) -> its canonical residue
template <bool M61>
__device__ __forceinline__ void canon_pair(uint32_t &lo, uint32_t &hi, const FieldParams &f) {
    if ((int32_t)hi < 0) {
        const int64_t v = (int64_t)pack(lo, hi);
        unpack(M61 ? canon_negative(v) : canon_negative_generic(f, v), lo, hi);
    }
}

template <int K, int T, int N, int ROUNDS, bool M61>
__global__ void __launch_bounds__(CTA, SDA_TC_MINBLOCKS)
packed_share_tc_kernel(const int64_t *__restrict__ secrets, size_t ld, size_t dim, size_t B, size_t unit_begin, size_t units_per_p,
                       size_t units_total, const ChaChaKey *__restrict__ keys, const uint4 *__restrict__ b_image,
                       int64_t *__restrict__ out, uint32_t two16, const __grid_constant__ GenericField gf, unsigned *flag,
                       int bulk_ok) {
    typedef Shape<K, T, N> S;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sD = smem;                                    // 2 x (G tiles x 128 rows x draws)
    uint8_t *sS = smem + 2 * S::D_BYTES;                   // G tiles x 128 rows x secrets
    uint8_t *sB = sS + S::S_BYTES;                         // the constant operand
    int64_t *sIn = reinterpret_cast<int64_t *>(sB + S::B_BYTES);   // the coming pass's raw secrets (SDA_TC_BULK_IN)
    __shared__ __align__(8) uint64_t mbar[3];              // [0] full (MMAs done), [1] drained (TMEM read out), [2] secrets landed
    __shared__ uint32_t tmem_base;

    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- one-time setup: TMEM, barriers, constant operand ------------------------------------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base)), "n"(S::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar[1])), "n"(CTA) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar[2])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < S::B_BYTES / 16; i += CTA) reinterpret_cast<uint4 *>(sB)[i] = __ldg(b_image + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    const uint32_t my_taddr = taddr + ((uint32_t)(warp * 32) << 16);
    const uint32_t full_bar = smem_u32(&mbar[0]), drained_bar = smem_u32(&mbar[1]);
    const uint32_t d_base = smem_u32(sD), s_base = smem_u32(sS), b_base = smem_u32(sB);
    const uint32_t landed_bar = smem_u32(&mbar[2]), sin_addr = smem_u32(sIn);
    const size_t row_bytes = B * sizeof(int64_t);           // distance between the share rows of a participant
    uint32_t parity = 0, buf = 0, landed_parity = 0;
    // a pass that lies wholly inside its vector arrives by bulk copy when the source is 16-byte aligned (bulk_ok)
    auto by_bulk = [&](size_t uu) { return bulk_ok != 0 && (uu + 1) * (size_t)(S::G * CTA) * K <= dim; };

    // (participant, pass) of this CTA's current unit and of its next one
    // (32-bit division: a 64-bit one is a call, and a call anywhere in the kernel makes ptxas keep the global
    // memory descriptor in a vector register and copy it to a uniform one at every load and store)
    // a launch covers passes unit_begin .. unit_begin + units_per_p - 1 of every participant (the host entry point
    // walks a vector in slices so that its copies overlap the kernel; device callers pass the whole vector)
    const size_t unit_end = unit_begin + units_per_p;
    size_t p = blockIdx.x / (uint32_t)units_per_p, u = unit_begin + blockIdx.x % (uint32_t)units_per_p;
#if SDA_TC_BULK_IN
    if (blockIdx.x < units_total) {
        if (by_bulk(u)) {
            if (tid == 0) bulk_load_secrets<S, K>(secrets, ld, p, u, sin_addr, landed_bar);
        } else {
            fill_secrets<S, K>(secrets, ld, dim, p, u, tid, sIn);
        }
        stage_draws<S, ROUNDS, M61>(keys, p, u, tid, sD, gf, flag);
    }
#else
    uint4 s[S::G][S::SC];                                    // secrets of the coming pass, prefetched
    if (blockIdx.x < units_total) {
        load_secrets<S, K>(secrets, ld, dim, p, u, tid, s);
        stage_draws<S, ROUNDS, M61>(keys, p, u, tid, sD, gf, flag);
    }
#endif

    for (size_t unit = blockIdx.x; unit < units_total; unit += gridDim.x) {
        const size_t b_base_batch = u * (size_t)(S::G * CTA);
        // ---- secrets of row `tid` of every tile (loaded during the previous pass): their little-endian
        //      bytes are the limbs ----------------------------------------------------------------------
        {
#if SDA_TC_BULK_IN
            // the pass's raw secrets are in shared memory (bulk copy issued a pass ago, or this thread's own stores)
            if (by_bulk(u)) {
                mbar_wait(landed_bar, landed_parity);
                landed_parity ^= 1;
            }
            uint4 s[S::G][S::SC];
#pragma unroll
            for (int q = 0; q < S::G; q++) {
                const int64_t *row = sIn + (q * CTA + tid) * K;
#pragma unroll
                for (int c = 0; c < S::SC; c++) {
                    const uint2 a = *reinterpret_cast<const uint2 *>(row + 2 * c);
                    const uint2 b = 2 * c + 1 < K ? *reinterpret_cast<const uint2 *>(row + 2 * c + 1) : make_uint2(0, 0);
                    s[q][c] = make_uint4(a.x, a.y, b.x, b.y);
                }
            }
#endif
#pragma unroll
            for (int q = 0; q < S::G; q++) {
                uint32_t sign = 0;
#pragma unroll
                for (int c = 0; c < S::SC; c++) sign |= s[q][c].y | s[q][c].w;
                if ((int32_t)sign < 0) {
#pragma unroll
                    for (int c = 0; c < S::SC; c++) {
                        canon_pair<M61>(s[q][c].x, s[q][c].y, gf.f);
                        canon_pair<M61>(s[q][c].z, s[q][c].w, gf.f);
                    }
                }
#pragma unroll
                for (int c = 0; c < S::SC; c++)
                    *reinterpret_cast<uint4 *>(sS + q * S::S_TILE + (tid >> 3) * S::SBO_S + c * LBO + (tid & 7) * 16) = s[q][c];
            }
        }
        // rows complete: the secrets just written and the draws written during the previous pass
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_cur = d_base + buf * S::D_BYTES;
        size_t pn = p, un = u + gridDim.x;                   // this CTA's next unit
        while (un >= unit_end) {
            un -= units_per_p;
            pn++;
        }
        const bool more = unit + gridDim.x < units_total;
        if (tid == 0) {
            issue_tiles<S>(taddr, d_cur, s_base, b_base, full_bar, 0);
#if SDA_TC_BULK_IN
            // everyone is past the barrier, i.e. has read this pass's raw secrets: the next pass's may land
            if (more && by_bulk(un)) bulk_load_secrets<S, K>(secrets, ld, pn, un, sin_addr, landed_bar);
#endif
        }

        // ---- the next pass's keystream, under this pass's first MMAs -----------------------------------
        if (more) stage_draws<S, ROUNDS, M61>(keys, pn, un, tid, sD + (buf ^ 1) * S::D_BYTES, gf, flag);
#if SDA_TC_BULK_IN
        if (more && !by_bulk(un)) fill_secrets<S, K>(secrets, ld, dim, pn, un, tid, sIn);
#elif SDA_TC_PREFETCH == 2
        if (more) load_secrets<S, K>(secrets, ld, dim, pn, un, tid, s);
#endif

        // ---- per group of ACC_BUFS tiles: D = A . B^T on the tensor core, then compose the shares -----------
        // `full` completes when the group's MMAs have written TMEM; `drained` when all 128 threads have read
        // their lanes out of it, so thread 0 can launch the next group's MMAs under everyone's (and its own)
        // compose arithmetic of the group's last tile instead of after a CTA-wide barrier.
        // this thread's column of the share rows: batch b_first + q * CTA of tile q, live while it is below B
        const size_t b_first = b_base_batch + tid;
        char *ob = reinterpret_cast<char *>(out + p * (size_t)N * B + b_first);
        const size_t rows_left = b_first < B ? B - b_first : 0;
        const uint32_t live_rows = rows_left < (size_t)(S::G * CTA) ? (uint32_t)rows_left : (uint32_t)(S::G * CTA);
#pragma unroll 1
        for (int q = 0; q < S::G; q += S::ACC_BUFS) {
            mbar_wait(full_bar, parity);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int a = 0; a < S::ACC_BUFS; a++) {
                uint32_t d[N][8];
                tmem_ld_shares<N>(my_taddr + a * S::ACC_COLS, d);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (a == S::ACC_BUFS - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(drained_bar) : "memory");
#if !SDA_TC_BULK_IN && SDA_TC_PREFETCH == 1
                    // next pass's secrets: issued under the last tile's compose arithmetic, consumed after it
                    if (q + S::ACC_BUFS >= S::G && more) load_secrets<S, K>(secrets, ld, dim, pn, un, tid, s);
#endif
                    if (tid == 0 && q + S::ACC_BUFS < S::G) {
                        mbar_wait(drained_bar, parity);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        issue_tiles<S>(taddr, d_cur, s_base, b_base, full_bar, q + S::ACC_BUFS);
                    }
                }
                const bool live = (uint32_t)((q + a) * CTA) < live_rows;
#pragma unroll
                for (int j = 0; j < N; j++) {
                    const uint64_t r = M61 ? compose(d[j], two16) : compose_generic(d[j], two16, gf.f);
                    if (live) *reinterpret_cast<int64_t *>(ob + (size_t)j * row_bytes) = (int64_t)r;
                }
                ob += CTA * sizeof(int64_t);
            }
            parity ^= 1;
        }
        // every thread is past its TMEM loads of the last tile and every MMA of this pass has completed
        // (`full` was waited on), so the next pass may overwrite the secrets and reuse TMEM
        p = pn;
        u = un;
        buf ^= 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(S::TMEM_COLS) : "memory");
}

// =================================================================================================
// Fused participant -> clerk path (the reference's FIXME at client/src/clerk.rs:71-72):
//     out[j][b] = acc_in[j][b] + sum_p share_j(participant p, batch b)   (mod p)
// without ever materialising shares[P][n][B].  The GEMM formulation makes the sum free: a tile of 128
// batches is multiplied for participant after participant INTO THE SAME TMEM ACCUMULATOR (tcgen05.mma
// with accumulate), so the eight limb sums of every share simply keep growing -- at most 96 * 255^2 per
// participant, i.e. 256 participants fit the s32 accumulators -- and the per-share compose arithmetic runs
// once per 256 participants instead of once per participant.  What is left per participant is the
// keystream, the staging of the rows and one MMA.
//
// A CTA owns batch range r (128 batches) and walks the participants four at a time: warp w produces the
// keystream of participant p + w for the range (16 T blocks per tile = T / 2 per lane), thread tid stages
// the secrets of batch r * 128 + tid of all four, and the four tiles are four MMA groups on one accumulator.
template <int K, int T, int N>
struct FusedShape : Shape<K, T, N> {
    typedef Shape<K, T, N> S;
    static constexpr int GP = 4;                               // participants per pass (one per warp)
    static constexpr int NBW = T / 2;                          // keystream blocks per lane per pass
    static constexpr uint32_t D_BYTES = GP * S::D_TILE + 128, S_BYTES = GP * S::S_TILE + 128;
    // the raw secrets of one pass: GP participants x [128 batches][K] i64, each block contiguous in its vector
    static constexpr uint32_t IN_ROW_BYTES = 128 * K * 8;
    static constexpr uint32_t IN_BYTES = SDA_TC_BULK_IN ? GP * IN_ROW_BYTES : 0;
    static constexpr uint32_t SMEM = 2 * D_BYTES + S_BYTES + S::B_BYTES + IN_BYTES;
    static constexpr int TMEM_COLS = S::NMMA <= 32 ? 32 : S::NMMA <= 64 ? 64 : 128;
    static constexpr int MAX_ACCUM = 256;                      // participants per TMEM accumulation: 256 * 96 * 255^2 < 2^31
};

// limb sums d[s] < 2^31 -> canonical sum_s d[s] 2^{8s} mod p (runs once per 256 participants: clarity over speed)
__device__ __forceinline__ uint64_t compose_wide(const uint32_t (&d)[8]) {
    uint64_t lo = 0, hi = 0;                                   // value = lo + hi 2^32, both < 2^57
#pragma unroll
    for (int s = 0; s < 4; s++) {
        lo += (uint64_t)d[s] << (8 * s);
        hi += (uint64_t)d[4 + s] << (8 * s);
    }
    // hi 2^32 = (hi mod 2^29) 2^32 + (hi >> 29) 2^61 == (hi mod 2^29) 2^32 + (hi >> 29)
    uint64_t v = lo + ((hi & LOW29) << 32) + (hi >> 29);       // < 2^57 + 2^61 + 2^28
    v = (v & P61) + (v >> 61);
    return v >= P61 ? v - P61 : v;
}

// pres != nullptr: the participants' precomputed first-round constants (chacha_pre.cuh); the launcher passes them when
// every block counter of the launch is below 2^32
template <class F, int ROUNDS>
__device__ __forceinline__ void stage_draws_fused(const ChaChaKey *__restrict__ keys, const ChaChaPre *__restrict__ pres, size_t p,
                                                  size_t r, int warp, int lane, uint8_t *sD, unsigned *flag) {
    typedef typename F::S S;
    uint32_t k[8], pre[12];
    const uint4 *src = reinterpret_cast<const uint4 *>(keys + p);
    const uint4 ka = __ldg(src), kb = __ldg(src + 1);
    k[0] = ka.x; k[1] = ka.y; k[2] = ka.z; k[3] = ka.w;
    k[4] = kb.x; k[5] = kb.y; k[6] = kb.z; k[7] = kb.w;
    if (pres != nullptr) {
        const uint4 *ps = reinterpret_cast<const uint4 *>(pres + p);
        const uint4 pa = __ldg(ps), pb = __ldg(ps + 1), pc = __ldg(ps + 2);
        pre[0] = pa.x; pre[1] = pa.y; pre[2] = pa.z; pre[3] = pa.w;
        pre[4] = pb.x; pre[5] = pb.y; pre[6] = pb.z; pre[7] = pb.w;
        pre[8] = pc.x; pre[9] = pc.y; pre[10] = pc.z; pre[11] = pc.w;
    }
#pragma unroll 1
    for (int nb = 0; nb < F::NBW; nb++) {
        const uint32_t blk = nb * 32 + lane;                       // block of this tile, in stream order
        uint32_t w[16];
        if (pres != nullptr) chacha_block2<ROUNDS>(k, pre, (uint32_t)(r * (size_t)(16 * S::kT) + blk), w);
        else chacha_block<ROUNDS>(k, r * (size_t)(16 * S::kT) + blk, w);
        uint32_t suspect = 0;
#pragma unroll
        for (int cb = 0; cb < 4; cb++) {
            const uint64_t xa = reduce_draw(w[4 * cb], w[4 * cb + 1], suspect);
            const uint64_t xb = reduce_draw(w[4 * cb + 2], w[4 * cb + 3], suspect);
            const uint32_t gc = blk * 4 + cb;                      // chunk index of the tile
            const uint32_t row = gc / S::DC, c = gc % S::DC;
            uint32_t xal, xah, xbl, xbh;
            unpack(xa, xal, xah);
            unpack(xb, xbl, xbh);
            *reinterpret_cast<uint4 *>(sD + warp * S::D_TILE + (row >> 3) * S::SBO_D + c * LBO + (row & 7) * 16) =
                make_uint4(xal, xah, xbl, xbh);
        }
        if (suspect >= LOW29) {
            bool bad = false;
#pragma unroll
            for (int d = 0; d < 8; d++) bad |= (w[2 * d] & LOW29) == LOW29 && w[2 * d + 1] >= 0xffffffe0u;
            if (bad) atomicOr(flag, 1u);
        }
    }
}

// the K secrets of batch b of participants p0 .. p0 + GP - 1 (zero beyond P and beyond the vector), as the
// 16-byte chunks they are staged as
template <class F, int K>
__device__ __forceinline__ void load_secrets_fused(const int64_t *__restrict__ secrets, size_t ld, size_t dim, size_t P,
                                                   size_t p0, size_t e0, uint4 (&s)[F::GP][F::S::SC]) {
#pragma unroll
    for (int q = 0; q < F::GP; q++) {
        const int64_t *sec = secrets + (p0 + q) * ld;
#pragma unroll
        for (int c = 0; c < F::S::SC; c++) {
            uint2 a = make_uint2(0, 0), b = make_uint2(0, 0);
            if (p0 + q < P && e0 + 2 * c < dim) a = __ldg(reinterpret_cast<const uint2 *>(sec + e0 + 2 * c));
            if (2 * c + 1 < K && p0 + q < P && e0 + 2 * c + 1 < dim) b = __ldg(reinterpret_cast<const uint2 *>(sec + e0 + 2 * c + 1));
            s[q][c] = make_uint4(a.x, a.y, b.x, b.y);
        }
    }
}

// the same rows into the pass's staging buffer by the threads themselves (ragged last range, unaligned sources)
template <class F, int K>
__device__ __forceinline__ void fill_secrets_fused(const int64_t *__restrict__ secrets, size_t ld, size_t dim, size_t P,
                                                   size_t p0, size_t e0, int tid, int64_t *sIn) {
#pragma unroll
    for (int q = 0; q < F::GP; q++) {
        const int64_t *sec = secrets + (p0 + q) * ld;
#pragma unroll
        for (int i = 0; i < K; i++)
            sIn[q * (CTA * K) + tid * K + i] = (p0 + q < P && e0 + i < dim) ? __ldg(sec + e0 + i) : 0;
    }
}

// one thread: batch range r of participants p0 .. p0 + np - 1, one bulk copy each, completion by byte count on `bar`
template <class F, int K>
__device__ __forceinline__ void bulk_load_secrets_fused(const int64_t *__restrict__ secrets, size_t ld, size_t p0, int np,
                                                        size_t r, uint32_t sin_addr, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(np * F::IN_ROW_BYTES) : "memory");
    for (int q = 0; q < np; q++) {
        const int64_t *src = secrets + (p0 + q) * ld + r * (size_t)(CTA * K);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(sin_addr + q * F::IN_ROW_BYTES), "l"(src), "r"(F::IN_ROW_BYTES), "r"(bar) : "memory");
    }
}

template <int K, int T, int N, int ROUNDS>
__global__ void __launch_bounds__(CTA)
packed_share_combine_tc_kernel(const int64_t *__restrict__ secrets, size_t ld, size_t dim, size_t B, size_t P,
                               const ChaChaKey *__restrict__ keys, const uint4 *__restrict__ b_image,
                               const int64_t *acc_in, int64_t *out, unsigned *flag, int bulk_ok,   // acc_in may equal out
                               const ChaChaPre *__restrict__ pres) {
    typedef FusedShape<K, T, N> F;
    typedef Shape<K, T, N> S;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sD = smem;                                    // 2 x (4 participants x 128 rows x draws)
    uint8_t *sS = smem + 2 * F::D_BYTES;                   // 4 participants x 128 rows x secrets
    uint8_t *sB = sS + F::S_BYTES;
    int64_t *sIn = reinterpret_cast<int64_t *>(sB + S::B_BYTES);   // the coming pass's raw secrets (SDA_TC_BULK_IN)
    __shared__ __align__(8) uint64_t mbar, landed;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base)), "n"(F::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&landed)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < S::B_BYTES / 16; i += CTA) reinterpret_cast<uint4 *>(sB)[i] = __ldg(b_image + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    const uint32_t my_taddr = taddr + ((uint32_t)(warp * 32) << 16);
    const uint32_t full_bar = smem_u32(&mbar), landed_bar = smem_u32(&landed), sin_addr = smem_u32(sIn);
    const uint32_t d_base = smem_u32(sD), s_base = smem_u32(sS), b_base = smem_u32(sB);
    uint32_t parity = 0, buf = 0, landed_parity = 0;

    const size_t ranges = (B + CTA - 1) / CTA;
    const size_t groups = (P + F::GP - 1) / F::GP;
    // a batch range that lies wholly inside the vectors arrives by bulk copy when the sources are 16-byte aligned
    auto by_bulk = [&](size_t rr) { return bulk_ok != 0 && (rr + 1) * (size_t)(CTA * K) <= dim; };
    auto participants_of = [&](size_t pg) { return (int)(P - pg < (size_t)F::GP ? P - pg : F::GP); };
    // keystream and secrets of the first pass
#if SDA_TC_BULK_IN
    if (blockIdx.x < ranges) {
        if (by_bulk(blockIdx.x)) {
            if (tid == 0) bulk_load_secrets_fused<F, K>(secrets, ld, 0, participants_of(0), blockIdx.x, sin_addr, landed_bar);
        } else {
            fill_secrets_fused<F, K>(secrets, ld, dim, P, 0, ((size_t)blockIdx.x * CTA + tid) * K, tid, sIn);
        }
        if ((size_t)warp < P) stage_draws_fused<F, ROUNDS>(keys, pres, warp, blockIdx.x, warp, lane, sD, flag);
    }
#else
    uint4 s[F::GP][S::SC];
    if (blockIdx.x < ranges) {
        load_secrets_fused<F, K>(secrets, ld, dim, P, 0, ((size_t)blockIdx.x * CTA + tid) * K, s);
        if ((size_t)warp < P) stage_draws_fused<F, ROUNDS>(keys, pres, warp, blockIdx.x, warp, lane, sD, flag);
    }
#endif

    for (size_t r = blockIdx.x; r < ranges; r += gridDim.x) {
        const size_t b = r * CTA + tid;                    // this thread's batch
        const size_t e0 = b * K;
        uint64_t acc[N];
#pragma unroll
        for (int j = 0; j < N; j++) {
            int64_t a = (acc_in != nullptr && b < B) ? acc_in[(size_t)j * B + b] : 0;
            if (a < 0) a = (int64_t)canon_negative(a);
            acc[j] = (uint64_t)a >= P61 ? (((uint64_t)a & P61) + ((uint64_t)a >> 61)) % P61 : (uint64_t)a;
        }
        int in_tmem = 0;                                   // participants accumulated in TMEM since the last drain
        for (size_t g = 0; g < groups; g++) {
            const size_t p0 = g * F::GP;
            const int np = (int)(P - p0 < (size_t)F::GP ? P - p0 : F::GP);
            // ---- secrets of batch b of the np participants: their bytes are the limbs -----------
#if SDA_TC_BULK_IN
            if (by_bulk(r)) {
                mbar_wait(landed_bar, landed_parity);
                landed_parity ^= 1;
            }
            uint4 s[F::GP][S::SC];
#pragma unroll
            for (int q = 0; q < F::GP; q++) {
                const int64_t *row = sIn + q * (CTA * K) + tid * K;
#pragma unroll
                for (int c = 0; c < S::SC; c++) {
                    uint2 a = make_uint2(0, 0), b2 = make_uint2(0, 0);
                    if (q < np) {
                        a = *reinterpret_cast<const uint2 *>(row + 2 * c);
                        if (2 * c + 1 < K) b2 = *reinterpret_cast<const uint2 *>(row + 2 * c + 1);
                    }
                    s[q][c] = make_uint4(a.x, a.y, b2.x, b2.y);
                }
            }
#endif
#pragma unroll
            for (int q = 0; q < F::GP; q++) {
                if (q < np) {
                    uint32_t sign = 0;
#pragma unroll
                    for (int c = 0; c < S::SC; c++) sign |= s[q][c].y | s[q][c].w;
                    if ((int32_t)sign < 0) {
                        const FieldParams unused{};
#pragma unroll
                        for (int c = 0; c < S::SC; c++) {
                            canon_pair<true>(s[q][c].x, s[q][c].y, unused);
                            canon_pair<true>(s[q][c].z, s[q][c].w, unused);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < S::SC; c++)
                        *reinterpret_cast<uint4 *>(sS + q * S::S_TILE + (tid >> 3) * S::SBO_S + c * LBO + (tid & 7) * 16) = s[q][c];
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 0) {
                const uint32_t d_cur = d_base + buf * F::D_BYTES;
                const uint64_t db = umma_desc(b_base, S::SBO_B);
                for (int q = 0; q < np; q++) {
                    const uint64_t dd = umma_desc(d_cur + q * S::D_TILE, S::SBO_D), ds = umma_desc(s_base + q * S::S_TILE, S::SBO_S);
#pragma unroll
                    for (int kk = 0; kk < S::NKD; kk++)
                        umma_i8(taddr, dd + ((2 * LBO * kk) >> 4), db + ((2 * LBO * kk) >> 4), S::IDESC, (in_tmem + q + kk) > 0);
#pragma unroll
                    for (int kk = 0; kk < S::NKS; kk++)
                        umma_i8(taddr, ds + ((2 * LBO * kk) >> 4), db + ((2 * LBO * (S::NKD + kk)) >> 4), S::IDESC, 1);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(full_bar) : "memory");
            }
            in_tmem += np;
            // ---- the next pass's secrets and keystream (next participants of this range, or the next range) under the MMAs
            {
                size_t rn = r, pg = p0 + F::GP;
                if (g + 1 == groups) {
                    rn = r + gridDim.x;
                    pg = 0;
                }
                if (rn < ranges) {
#if SDA_TC_BULK_IN
                    // everyone is past the barrier, i.e. has read this pass's raw secrets: the next pass's may land
                    if (by_bulk(rn)) {
                        if (tid == 0) bulk_load_secrets_fused<F, K>(secrets, ld, pg, participants_of(pg), rn, sin_addr, landed_bar);
                    } else {
                        fill_secrets_fused<F, K>(secrets, ld, dim, P, pg, (rn * CTA + tid) * K, tid, sIn);
                    }
#else
                    load_secrets_fused<F, K>(secrets, ld, dim, P, pg, (rn * CTA + tid) * K, s);     // consumed after the keystream
#endif
                    if (pg + warp < P) stage_draws_fused<F, ROUNDS>(keys, pres, pg + warp, rn, warp, lane, sD + (buf ^ 1) * F::D_BYTES, flag);
                }
            }
            mbar_wait(full_bar, parity);                   // rows consumed: the next pass may overwrite the secrets
            parity ^= 1;
            buf ^= 1;
            // ---- drain TMEM when the accumulators are about to fill up, and at the end of the range -------
            if (in_tmem + F::GP > F::MAX_ACCUM || g + 1 == groups) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int j = 0; j < N; j++) {
                    uint32_t d[8];
                    tmem_ld8(my_taddr + 8 * j, d);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const uint64_t v = acc[j] + compose_wide(d);
                    acc[j] = v >= P61 ? v - P61 : v;
                }
                in_tmem = 0;
                // the next MMA (accumulate = 0) is issued after the next staging barrier, which every thread
                // reaches only after these loads
            }
        }
        if (b < B) {
#pragma unroll
            for (int j = 0; j < N; j++) out[(size_t)j * B + b] = (int64_t)acc[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(F::TMEM_COLS) : "memory");
}

// the constant operand as it lies in shared memory
template <int K, int T, int N>
void build_b_image(const Matrix &m, uint64_t p, uint8_t *img) {
    typedef Shape<K, T, N> S;
    typedef unsigned __int128 u128;
    memset(img, 0, S::B_BYTES);
    for (int j = 0; j < N; j++)
        for (int part = 0; part < 2; part++)                // 0: draws (K steps 0..NKD), 1: secrets (the rest)
            for (int c = 0; c < (part ? S::SC : S::DC); c++)
                for (int v = 0; v < 2; v++) {
                    const int idx = 2 * c + v;                  // draw / secret index within the batch
                    if (idx >= (part ? K : T)) continue;
                    const int xi = part ? idx : K + idx;        // index into x = [secrets ; randomness]
                    const int cg = part ? 2 * S::NKD + c : c;   // chunk along K of the whole row
                    for (int byte = 0; byte < 8; byte++) {
                        const uint64_t cst = (uint64_t)((u128)m.e[j * (K + T) + xi] * ((((u128)1) << (8 * byte)) % p) % p);
                        for (int s = 0; s < 8; s++) {
                            const int n = j * 8 + s;
                            img[(n / 8) * S::SBO_B + cg * LBO + (n % 8) * 16 + v * 8 + byte] = (uint8_t)(cst >> (8 * s));
                        }
                    }
                }
}

template <int K, int T, int N, int ROUNDS, bool M61>
cudaError_t launch(const LaunchCtx &lc, const GenericField &gf, const int64_t *secrets, size_t ld, size_t P, size_t dim,
                   size_t first_batch, size_t n_batches, const ChaChaKey *keys, const uint8_t *d_b_image, int64_t *out,
                   unsigned *flag) {
    typedef Shape<K, T, N> S;
    const size_t B = (dim + K - 1) / K;
    if (first_batch % (S::G * CTA) != 0 || first_batch > B) return cudaErrorInvalidValue;
    if (n_batches > B - first_batch) n_batches = B - first_batch;
    const size_t unit_begin = first_batch / (S::G * CTA);
    const size_t units_per_p = (n_batches + S::G * CTA - 1) / (S::G * CTA);
    const size_t units_total = units_per_p * P;
    if (units_total == 0) return cudaSuccess;
    if (units_per_p >> 32) return cudaErrorInvalidValue;
    auto kern = packed_share_tc_kernel<K, T, N, ROUNDS, M61>;
    // never more CTAs on an SM than can hold their TMEM columns (tc_common.cuh)
    const size_t smem = smem_capping_residency(S::SMEM, 512 / S::TMEM_COLS);
    static KernelSetup setup;
    int regs = 0;
    size_t static_smem = 0;
    const cudaError_t se = setup_kernel(setup, kern, smem, &regs, &static_smem);
    if (se != cudaSuccess) return se;
    const int per_sm = resident_ctas(regs, CTA, smem, static_smem, S::TMEM_COLS);
    size_t grid = (size_t)lc.sm_count * per_sm;
    if (grid > units_total) grid = units_total;
    // bulk copies need 16-byte aligned sources: every pass of every participant starts at an even element
    const int bulk_ok = SDA_TC_BULK_IN && reinterpret_cast<uintptr_t>(secrets) % 16 == 0 && (ld % 2 == 0 || P == 1);
    kern<<<(unsigned)grid, CTA, smem, lc.stream>>>(secrets, ld, dim, B, unit_begin, units_per_p, units_total, keys,
                                                   reinterpret_cast<const uint4 *>(d_b_image), out, 65536u, gf, flag, bulk_ok);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

template <int K, int T, int N>
cudaError_t dispatch(const LaunchCtx &lc, const GenericField &gf, int rounds, const int64_t *secrets, size_t ld, size_t P,
                     size_t dim, size_t first_batch, size_t n_batches, const ChaChaKey *keys, const uint8_t *d_b_image,
                     int64_t *out, unsigned *flag) {
#define SDA_L(R, M) return launch<K, T, N, R, M>(lc, gf, secrets, ld, P, dim, first_batch, n_batches, keys, d_b_image, out, flag)
    if (gf.f.kind == FIELD_MERSENNE61) {
        if (rounds == 8) SDA_L(8, true);
        if (rounds == 12) SDA_L(12, true);
        SDA_L(20, true);
    }
    if (rounds == 8) SDA_L(8, false);
    if (rounds == 12) SDA_L(12, false);
    SDA_L(20, false);
#undef SDA_L
}

}  // namespace

template <int K, int T, int N, int ROUNDS>
cudaError_t launch_fused(const LaunchCtx &lc, const int64_t *secrets, size_t ld, size_t P, size_t dim, const ChaChaKey *keys,
                         const uint8_t *d_b_image, const int64_t *acc_in, int64_t *out, unsigned *flag, uint32_t *d_pre) {
    typedef FusedShape<K, T, N> F;
    const size_t B = (dim + K - 1) / K;
    const size_t ranges = (B + CTA - 1) / CTA;
    const size_t smem = smem_capping_residency(F::SMEM, 512 / F::TMEM_COLS);
    auto kern = packed_share_combine_tc_kernel<K, T, N, ROUNDS>;
    static KernelSetup setup;
    int regs = 0;
    size_t static_smem = 0;
    const cudaError_t se = setup_kernel(setup, kern, smem, &regs, &static_smem);
    if (se != cudaSuccess) return se;
    const int per_sm = resident_ctas(regs, CTA, smem, static_smem, F::TMEM_COLS);
    const size_t grid = std::min<size_t>(ranges, (size_t)lc.sm_count * per_sm);
    const int bulk_ok = SDA_TC_BULK_IN && reinterpret_cast<uintptr_t>(secrets) % 16 == 0 && (ld % 2 == 0 || P == 1);
    // first-round constants per participant (d_pre: P ChaChaPre) when a participant's keystream stays below 2^32 blocks
    ChaChaPre *pres = nullptr;
#ifdef SDA_TC_FUSED_NO_PRE
    d_pre = nullptr;
#endif
    if (d_pre != nullptr && P > 0 && ((B * (size_t)T + 7) / 8 + 16 * (size_t)T) >> 32 == 0) {
        pres = reinterpret_cast<ChaChaPre *>(d_pre);
        chacha_prepare_kernel<<<(unsigned)((P + 127) / 128), 128, 0, lc.stream>>>(keys, P, pres);
        ++*lc.nlaunch;
    }
    kern<<<(unsigned)grid, CTA, smem, lc.stream>>>(secrets, ld, dim, B, P, keys, reinterpret_cast<const uint4 *>(d_b_image),
                                                   acc_in, out, flag, bulk_ok, pres);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

template <int K, int T, int N>
cudaError_t dispatch_fused(const LaunchCtx &lc, int rounds, const int64_t *secrets, size_t ld, size_t P, size_t dim,
                           const ChaChaKey *keys, const uint8_t *d_b_image, const int64_t *acc_in, int64_t *out, unsigned *flag,
                           uint32_t *d_pre) {
    if (rounds == 8) return launch_fused<K, T, N, 8>(lc, secrets, ld, P, dim, keys, d_b_image, acc_in, out, flag, d_pre);
    if (rounds == 12) return launch_fused<K, T, N, 12>(lc, secrets, ld, P, dim, keys, d_b_image, acc_in, out, flag, d_pre);
    return launch_fused<K, T, N, 20>(lc, secrets, ld, P, dim, keys, d_b_image, acc_in, out, flag, d_pre);
}

#define SDA_TC_SHAPES(X) X(3, 2, 5) X(5, 4, 9) X(3, 4, 7) X(3, 4, 8)

size_t packed_share_tc_image_bytes(int k, int t, int n) {
#define X(K, T, N) if (k == K && t == T && n == N) return Shape<K, T, N>::B_BYTES;
    SDA_TC_SHAPES(X)
#undef X
    return 0;
}

void packed_share_tc_build_image(int k, int t, int n, const Matrix &mtx, uint64_t p, uint8_t *img) {
#define X(K, T, N) if (k == K && t == T && n == N) return build_b_image<K, T, N>(mtx, p, img);
    SDA_TC_SHAPES(X)
#undef X
}

// batches one CTA pass covers: slices of a vector handed to launch_packed_share_tc start at multiples of it
size_t packed_share_tc_slice_batches(int k, int t, int n) {
#define X(K, T, N) if (k == K && t == T && n == N) return (size_t)Shape<K, T, N>::G * CTA;
    SDA_TC_SHAPES(X)
#undef X
    return 0;
}

// d_b_image: device copy of the image built above (packed_share_tc_image_bytes bytes, 16-byte aligned).
// Batches first_batch .. first_batch + n_batches - 1 of every participant are generated (n_batches is clipped to
// the vector; first_batch is a multiple of packed_share_tc_slice_batches); `secrets` and `shares_out` always
// address whole vectors.
cudaError_t launch_packed_share_tc(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, int rounds, int k, int t,
                                   int n, const int64_t *secrets, size_t ld, size_t P, size_t dim, size_t first_batch,
                                   size_t n_batches, const ChaChaKey *keys, const uint8_t *d_b_image, int64_t *shares_out,
                                   unsigned *flag) {
    const GenericField gf{f, dr};
    const bool m61 = f.kind == FIELD_MERSENNE61;
#define X(K, T, N)                                                                                              \
    if (k == K && t == T && n == N) {                                                                           \
        *lc.kernel_name = m61 ? "packed_share<" #K "," #T "," #N ">/mersenne61 tcgen05.mma.kind::i8"            \
                              : "packed_share<" #K "," #T "," #N ">/any prime tcgen05.mma.kind::i8";            \
        return dispatch<K, T, N>(lc, gf, rounds, secrets, ld, P, dim, first_batch, n_batches, keys, d_b_image,  \
                                 shares_out, flag);                                                             \
    }
    SDA_TC_SHAPES(X)
#undef X
    return cudaErrorInvalidValue;
}

// fused share generation + clerk accumulation over the participants: out[n][B] = acc_in + sum_p shares(p)
cudaError_t launch_packed_share_combine_tc(const LaunchCtx &lc, int rounds, int k, int t, int n, const int64_t *secrets,
                                           size_t ld, size_t P, size_t dim, const ChaChaKey *keys, const uint8_t *d_b_image,
                                           const int64_t *acc_in, int64_t *out, unsigned *flag, uint32_t *d_key_scratch) {
#define X(K, T, N)                                                                                              \
    if (k == K && t == T && n == N) {                                                                           \
        *lc.kernel_name = "packed_share_combine<" #K "," #T "," #N ">/mersenne61 tcgen05.mma.kind::i8, TMEM-accumulated"; \
        return dispatch_fused<K, T, N>(lc, rounds, secrets, ld, P, dim, keys, d_b_image, acc_in, out, flag, d_key_scratch); \
    }
    SDA_TC_SHAPES(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace sda
