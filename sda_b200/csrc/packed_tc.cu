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
// IMAD.WIDE stream is what bounds it (profiles/r01_k2.md).
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
#include <algorithm>
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"

namespace sda {

namespace {

using namespace tc;

constexpr int CTA = 128;             // threads = rows of one MMA tile = TMEM lanes

constexpr int gcd_c(int a, int b) { return b == 0 ? a : gcd_c(b, a % b); }

template <int K, int T, int N>
struct Shape {
    static_assert(T % 2 == 0 && T >= 2, "draws fill whole 16-byte chunks");
    static constexpr int G = 8 / gcd_c(T, 8);              // 128-batch tiles per pass of a CTA
    static constexpr int NB = T / gcd_c(T, 8);             // keystream blocks per thread per pass
    static constexpr int DC = T / 2;                       // 16-byte chunks of a row holding draws
    static constexpr int SC = (K + 1) / 2;                 // ... holding secrets
    static constexpr int C = DC + SC;
    static constexpr int NK = (C + 1) / 2;                 // MMAs per tile (32 bytes of K each)
    static constexpr uint32_t SBO_A = C * 128;             // an odd last chunk aliases the next group: B is 0 there
    static constexpr uint32_t A_TILE = 16 * SBO_A;
    static constexpr uint32_t A_BYTES = G * A_TILE + 128;  // + the aliased chunk past the last group
    static constexpr int NMMA = (8 * N + 15) / 16 * 16;
    static constexpr uint32_t SBO_B = 2 * NK * 128;
    static constexpr uint32_t B_BYTES = NMMA / 8 * SBO_B;
    static constexpr int TMEM_COLS = NMMA <= 32 ? 32 : NMMA <= 64 ? 64 : NMMA <= 128 ? 128 : 256;
    static constexpr uint32_t IDESC = idesc_u8(NMMA);
    static_assert(4 % DC == 0, "a keystream block covers whole rows");
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
#pragma unroll 2
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

// draw (hi word w0, lo word w1) -> v mod (p - 1) as (v & p) + 2 (v >> 61); `bad` when that is not gen_range's answer
__device__ __forceinline__ uint64_t reduce_draw(uint32_t w0, uint32_t w1, bool &bad) {
    const uint32_t h = w0 >> 29, hi = w0 & LOW29;
    bad = hi == LOW29 && w1 >= 0xffffffe0u;          // v mod 2^61 >= 2^61 - 32: rejected word or wrap-around
    return pack(w1, hi) + (uint64_t)(2u * h);
}

// one thread: the NK MMAs of a 128-row tile, completion signalled on `full_bar`
template <class S>
__device__ __forceinline__ void issue_tile(uint32_t taddr, uint32_t a_tile, uint32_t b_base, uint32_t full_bar) {
    const uint64_t da = umma_desc(a_tile, S::SBO_A), db = umma_desc(b_base, S::SBO_B);
#pragma unroll
    for (int kk = 0; kk < S::NK; kk++)
        umma_i8(taddr, da + ((2 * LBO * kk) >> 4), db + ((2 * LBO * kk) >> 4), S::IDESC, kk > 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(full_bar) : "memory");
}

template <int K, int T, int N, int ROUNDS>
__global__ void __launch_bounds__(CTA)
packed_share_tc_kernel(const int64_t *__restrict__ secrets, size_t ld, size_t dim, size_t B, size_t units_per_p,
                       size_t units_total, const ChaChaKey *__restrict__ keys, const uint4 *__restrict__ b_image,
                       int64_t *__restrict__ out, uint32_t two16, unsigned *flag) {
    typedef Shape<K, T, N> S;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sA = smem;                                   // G tiles of 128 rows
    uint8_t *sB = smem + ((S::A_BYTES + 127) & ~127u);    // the constant operand
    __shared__ __align__(8) uint64_t mbar[2];              // [0] full (MMA done), [1] drained (TMEM read out)
    __shared__ uint32_t tmem_base;

    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- one-time setup: TMEM, barrier, constant operand ------------------------------------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base)), "n"(S::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar[1])), "n"(CTA) : "memory");
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
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
    uint32_t parity = 0;

    size_t p = blockIdx.x / units_per_p, u = blockIdx.x % units_per_p;   // (participant, pass) of this CTA's unit
    for (size_t unit = blockIdx.x; unit < units_total; unit += gridDim.x, u += gridDim.x) {
        while (u >= units_per_p) {
            u -= units_per_p;
            p++;
        }
        const size_t b_base_batch = u * (size_t)(S::G * CTA);
        const int64_t *sec = secrets + p * ld;

        // ---- secrets of row `tid` of every tile: loads issued now, consumed after the keystream ----
        int64_t s[S::G][2 * S::SC];
#pragma unroll
        for (int q = 0; q < S::G; q++) {
            const size_t e0 = (b_base_batch + (size_t)q * CTA + tid) * K;
#pragma unroll
            for (int i = 0; i < 2 * S::SC; i++)
                s[q][i] = (i < K && e0 + i < dim) ? __ldg(sec + e0 + i) : 0;           // batched.rs:38-43 zero padding
        }
        // ---- randomness: NB keystream blocks per thread, reduced and scattered into the rows --
        {
            uint32_t k[8];
            const uint4 *src = reinterpret_cast<const uint4 *>(keys + p);
            const uint4 ka = __ldg(src), kb = __ldg(src + 1);
            k[0] = ka.x; k[1] = ka.y; k[2] = ka.z; k[3] = ka.w;
            k[4] = kb.x; k[5] = kb.y; k[6] = kb.z; k[7] = kb.w;
            bool any_bad = false;
#pragma unroll 1
            for (int nb = 0; nb < S::NB; nb++) {
                const uint32_t slot = nb * CTA + tid;                     // block of this pass, in stream order
                uint32_t w[16];
                chacha_block<ROUNDS>(k, u * (size_t)(CTA * S::NB) + slot, w);
#pragma unroll
                for (int cb = 0; cb < 4; cb++) {                          // 4 chunks of 2 draws
                    bool bad0, bad1;
                    const uint64_t xa = reduce_draw(w[4 * cb], w[4 * cb + 1], bad0);
                    const uint64_t xb = reduce_draw(w[4 * cb + 2], w[4 * cb + 3], bad1);
                    any_bad |= bad0 | bad1;
                    const uint32_t gc = slot * 4 + cb;                    // chunk index of the pass
                    const uint32_t batch = gc / S::DC, c = gc % S::DC;
                    const uint32_t q = batch / CTA, row = batch % CTA;
                    uint32_t xal, xah, xbl, xbh;
                    unpack(xa, xal, xah);
                    unpack(xb, xbl, xbh);
                    *reinterpret_cast<uint4 *>(sA + q * S::A_TILE + (row >> 3) * S::SBO_A + c * LBO + (row & 7) * 16) =
                        make_uint4(xal, xah, xbl, xbh);
                }
            }
            if (any_bad) atomicOr(flag, 1u);
        }
        // ---- secrets into the rows: their little-endian bytes are the limbs ----------------------
#pragma unroll
        for (int q = 0; q < S::G; q++) {
            uint32_t sign = 0;
#pragma unroll
            for (int i = 0; i < K; i++) sign |= (uint32_t)((uint64_t)s[q][i] >> 32);
            if ((int32_t)sign < 0) {
#pragma unroll
                for (int i = 0; i < K; i++)
                    if (s[q][i] < 0) s[q][i] = (int64_t)canon_negative(s[q][i]);
            }
#pragma unroll
            for (int c = 0; c < S::SC; c++) {
                uint32_t al, ah, bl, bh;
                unpack((uint64_t)s[q][2 * c], al, ah);
                unpack((uint64_t)s[q][2 * c + 1], bl, bh);
                *reinterpret_cast<uint4 *>(sA + q * S::A_TILE + (tid >> 3) * S::SBO_A + (S::DC + c) * LBO + (tid & 7) * 16) =
                    make_uint4(al, ah, bl, bh);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- per tile: D = A . B^T on the tensor core, then compose the shares ------------------
        // `full` completes when a tile's MMAs have written TMEM; `drained` when all 128 threads have
        // read their lane out of it, so thread 0 can launch the next tile's MMAs under everyone's
        // (and its own) compose arithmetic instead of after a CTA-wide barrier.
        if (tid == 0) issue_tile<S>(taddr, a_base, b_base, full_bar);
#pragma unroll 1
        for (int q = 0; q < S::G; q++) {
            mbar_wait(full_bar, parity);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t d[N][8];
#pragma unroll
            for (int j = 0; j < N; j++) tmem_ld8(my_taddr + 8 * j, d[j]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(drained_bar) : "memory");
            if (tid == 0 && q + 1 < S::G) {
                mbar_wait(drained_bar, parity);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                issue_tile<S>(taddr, a_base + (q + 1) * S::A_TILE, b_base, full_bar);
            }
            parity ^= 1;
            const size_t b = b_base_batch + (size_t)q * CTA + tid;
            int64_t *ob = out + p * (size_t)N * B + b;
#pragma unroll
            for (int j = 0; j < N; j++) {
                const uint64_t r = compose(d[j], two16);
                if (b < B) *ob = (int64_t)r;
                ob += B;
            }
        }
        // every thread is past its TMEM loads of the last tile, and every MMA of this pass has completed
        // (full was waited on): the staging barrier of the next pass orders the rest
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(S::TMEM_COLS) : "memory");
}

// the constant operand as it lies in shared memory
template <int K, int T, int N>
void build_b_image(const Matrix &m, uint8_t *img) {
    typedef Shape<K, T, N> S;
    typedef unsigned __int128 u128;
    memset(img, 0, S::B_BYTES);
    for (int j = 0; j < N; j++)
        for (int c = 0; c < S::C; c++)
            for (int v = 0; v < 2; v++) {
                int xi;                                    // index into x = [secrets ; randomness]
                if (c < S::DC) {
                    if (2 * c + v >= T) continue;
                    xi = K + 2 * c + v;
                } else {
                    if (2 * (c - S::DC) + v >= K) continue;
                    xi = 2 * (c - S::DC) + v;
                }
                for (int byte = 0; byte < 8; byte++) {
                    const uint64_t cst = (uint64_t)((u128)m.e[j * (K + T) + xi] * ((((u128)1) << (8 * byte)) % P61) % P61);
                    for (int s = 0; s < 8; s++) {
                        const int n = j * 8 + s;
                        img[(n / 8) * S::SBO_B + c * LBO + (n % 8) * 16 + v * 8 + byte] = (uint8_t)(cst >> (8 * s));
                    }
                }
            }
}

template <int K, int T, int N, int ROUNDS>
cudaError_t launch(const LaunchCtx &lc, const int64_t *secrets, size_t ld, size_t P, size_t dim, const ChaChaKey *keys,
                   const uint8_t *d_b_image, int64_t *out, unsigned *flag) {
    typedef Shape<K, T, N> S;
    const size_t B = (dim + K - 1) / K;
    const size_t units_per_p = (B + S::G * CTA - 1) / (S::G * CTA);
    const size_t units_total = units_per_p * P;
    const size_t smem = ((S::A_BYTES + 127) & ~127u) + S::B_BYTES;
    auto kern = packed_share_tc_kernel<K, T, N, ROUNDS>;
    static int per_sm = 0;      // resident CTAs per SM: every one of them must hold its TMEM columns
    if (per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncAttributes fa;
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, kern);
        if (e != cudaSuccess) return e;
        const int by_regs = 65536 / (((fa.numRegs + 7) & ~7) * CTA);
        const int by_smem = (int)((227u * 1024u) / (smem + fa.sharedSizeBytes + 1024));
        const int by_tmem = 512 / S::TMEM_COLS;
        per_sm = std::max(1, std::min(by_regs, std::min(by_smem, by_tmem)));
    }
    size_t grid = (size_t)lc.sm_count * per_sm;
    if (grid > units_total) grid = units_total;
    kern<<<(unsigned)grid, CTA, smem, lc.stream>>>(secrets, ld, dim, B, units_per_p, units_total, keys,
                                                   reinterpret_cast<const uint4 *>(d_b_image), out, 65536u, flag);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

template <int K, int T, int N>
cudaError_t dispatch(const LaunchCtx &lc, int rounds, const int64_t *secrets, size_t ld, size_t P, size_t dim,
                     const ChaChaKey *keys, const uint8_t *d_b_image, int64_t *out, unsigned *flag) {
    if (rounds == 8) return launch<K, T, N, 8>(lc, secrets, ld, P, dim, keys, d_b_image, out, flag);
    if (rounds == 12) return launch<K, T, N, 12>(lc, secrets, ld, P, dim, keys, d_b_image, out, flag);
    return launch<K, T, N, 20>(lc, secrets, ld, P, dim, keys, d_b_image, out, flag);
}

}  // namespace

#define SDA_TC_SHAPES(X) X(3, 2, 5) X(5, 4, 9) X(3, 4, 7) X(3, 4, 8)

size_t packed_share_tc_image_bytes(int k, int t, int n) {
#define X(K, T, N) if (k == K && t == T && n == N) return Shape<K, T, N>::B_BYTES;
    SDA_TC_SHAPES(X)
#undef X
    return 0;
}

void packed_share_tc_build_image(int k, int t, int n, const Matrix &mtx, uint8_t *img) {
#define X(K, T, N) if (k == K && t == T && n == N) return build_b_image<K, T, N>(mtx, img);
    SDA_TC_SHAPES(X)
#undef X
}

// d_b_image: device copy of the image built above (packed_share_tc_image_bytes bytes, 16-byte aligned)
cudaError_t launch_packed_share_tc(const LaunchCtx &lc, int rounds, int k, int t, int n, const int64_t *secrets, size_t ld,
                                   size_t P, size_t dim, const ChaChaKey *keys, const uint8_t *d_b_image,
                                   int64_t *shares_out, unsigned *flag) {
#define X(K, T, N)                                                                                        \
    if (k == K && t == T && n == N) {                                                                     \
        *lc.kernel_name = "packed_share<" #K "," #T "," #N ">/mersenne61 tcgen05.mma.kind::i8";           \
        return dispatch<K, T, N>(lc, rounds, secrets, ld, P, dim, keys, d_b_image, shares_out, flag);     \
    }
    SDA_TC_SHAPES(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace sda
