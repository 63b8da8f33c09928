// packed_tcg.cu -- packed-Shamir share generation on the tensor cores for ANY scheme shape the reference accepts
// (client/src/crypto/sharing/packed_shamir.rs:13-27, sharing/mod.rs:44-50): secret_count k, privacy_threshold t and
// share_count n are run-time values here (k + t <= 16, n <= 32, odd t included), any prime below 2^63.  It is the
// byte-limb GEMM of packed_tc.cu / packed_tc2.cu -- D = A . B^T on tcgen05.mma.kind::i8 with A = the raw bytes of
// [draws ; secrets] of a batch and B = the limbs of (M[j][i] 2^{8c} mod p) -- with everything the templated kernels
// fix at compile time read from a small parameter block, so that no scheme falls back to materialising its random
// draws in HBM one participant at a time.  The four shapes BASELINE names keep their templated kernels (this one
// trades their unrolling and their keystream / MMA overlap inside a CTA for generality; several CTAs per SM overlap
// instead).
//
// Row layout (K-major, no swizzle, tc_common.cuh): chunk c of a row is 16 bytes = two u64 values; chunks
// 0 .. DC-1 hold the t draws (DC = ceil(t/2)), chunks DC .. DC+SC-1 the k secrets (SC = ceil(k/2)); a half chunk
// without a value and the chunk that pads an odd chunk count to whole 32-byte K steps meet zero rows of B.
// A pass of a CTA = GS tiles of 128 consecutive batches of one participant (GS = 2 .. 8, chosen so that the pass's
// 16 GS t keystream blocks fill the 128 threads): the tiles are multiplied G at a time into G accumulators (G = 2, or 1
// for n > 16), and after each round every thread folds the limb sums of its rows, one share at a time (tcgen05.ld x8 per
// share, the next load issued under the current fold), and stores them; consecutive threads own consecutive batches, so
// stores coalesce.
#include <algorithm>
#include <cstring>
#include <vector>

#include "kernels.h"
#include "tc_common.cuh"

namespace sda {

namespace {

using namespace tc;

constexpr int CTAG = 128;

struct GShape {
    int k, t, n;
    int dc, sc, nch, nk;          // chunks of draws / secrets / per row, K steps
    int nmma, acc_cols, g, gs;    // MMA N, TMEM columns per accumulator, accumulators (tiles multiplied at once), tiles staged per pass
    int w5;                       // limb plan of the Mersenne path (packed_tc2.cu); 8 = plain bytes
    uint32_t sbo_a, a_tile, a_bytes, sbo_b, b_bytes;
    uint32_t idesc;
    uint32_t blocks_per_pass;     // keystream blocks a pass consumes: 16 gs t
};

inline int w5_for_g(int kt) {
    for (int w5 = 8; w5 >= 5; w5--)
        if ((long long)8 * kt * 255 * (255 + ((1ll << w5) - 1) * 256) < (1ll << 29)) return w5;
    return 0;
}

GShape make_shape(int k, int t, int n, bool m61) {
    GShape s{};
    s.k = k; s.t = t; s.n = n;
    s.dc = (t + 1) / 2;
    s.sc = (k + 1) / 2;
    s.nch = s.dc + s.sc;
    s.nk = (s.nch + 1) / 2;
    s.nmma = (8 * n + 15) / 16 * 16;
    s.acc_cols = s.nmma <= 32 ? 32 : s.nmma <= 64 ? 64 : s.nmma <= 128 ? 128 : 256;
    s.g = s.acc_cols <= 128 ? 2 : 1;
    s.w5 = m61 ? w5_for_g(k + t) : 8;
    s.sbo_a = (uint32_t)s.nch * 128;
    s.a_tile = 16 * s.sbo_a;
    // tiles staged per pass: a pass consumes 16 gs t keystream blocks, one per thread and round; the more of the 128
    // threads have a block in the last round the better (t = 2 needs gs = 4, t = 1 gs = 8), within 64 KB of operand tiles
    s.gs = s.g;
    {
        double best = 0;
        for (int gs = s.g; gs <= 8; gs *= 2) {
            if ((size_t)gs * s.a_tile > 64 * 1024 && gs > s.g) break;
            const int blocks = 16 * gs * t, rounds = (blocks + CTAG - 1) / CTAG;
            const double eff = (double)blocks / (rounds * CTAG);
            if (eff > best + 1e-9) {
                best = eff;
                s.gs = gs;
            }
        }
    }
    s.a_bytes = (uint32_t)s.gs * s.a_tile + 128;     // an odd chunk count reads one chunk past the last row group
    s.sbo_b = 2 * (uint32_t)s.nk * 128;
    s.b_bytes = (uint32_t)(s.nmma / 8) * s.sbo_b;
    s.idesc = idesc_u8(s.nmma);
    s.blocks_per_pass = 16u * (uint32_t)s.gs * (uint32_t)t;
    return s;
}

#define SDA_QRG(a, b, c, d)                                     \
    a += b; d ^= a; d = __funnelshift_l(d, d, 16);              \
    c += d; b ^= c; b = __funnelshift_l(b, b, 12);              \
    a += b; d ^= a; d = __funnelshift_l(d, d, 8);               \
    c += d; b ^= c; b = __funnelshift_l(b, b, 7);

template <int ROUNDS>
__device__ __forceinline__ void chacha_block_g(const uint32_t (&k)[8], uint64_t block, uint32_t (&o)[16]) {
    const uint32_t c0 = 0x61707865u, c1 = 0x3320646eu, c2 = 0x79622d32u, c3 = 0x6b206574u;
    const uint32_t b0 = (uint32_t)block, b1 = (uint32_t)(block >> 32);
    uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3;
    uint32_t x4 = k[0], x5 = k[1], x6 = k[2], x7 = k[3];
    uint32_t x8 = k[4], x9 = k[5], x10 = k[6], x11 = k[7];
    uint32_t x12 = b0, x13 = b1, x14 = 0, x15 = 0;
#pragma unroll 1
    for (int i = 0; i < ROUNDS / 2; i++) {
        SDA_QRG(x0, x4, x8, x12)
        SDA_QRG(x1, x5, x9, x13)
        SDA_QRG(x2, x6, x10, x14)
        SDA_QRG(x3, x7, x11, x15)
        SDA_QRG(x0, x5, x10, x15)
        SDA_QRG(x1, x6, x11, x12)
        SDA_QRG(x2, x7, x8, x13)
        SDA_QRG(x3, x4, x9, x14)
    }
    o[0] = x0 + c0;     o[1] = x1 + c1;     o[2] = x2 + c2;      o[3] = x3 + c3;
    o[4] = x4 + k[0];   o[5] = x5 + k[1];   o[6] = x6 + k[2];    o[7] = x7 + k[3];
    o[8] = x8 + k[4];   o[9] = x9 + k[5];   o[10] = x10 + k[6];  o[11] = x11 + k[7];
    o[12] = x12 + b0;   o[13] = x13 + b1;   o[14] = x14;         o[15] = x15;
}

// limb sums -> canonical residue.  Mersenne: packed_tc2.cu's fold with the limb plan's widths as run-time shifts;
// other primes: plain byte limbs composed into a 79-bit integer and reduced once (packed_tc.cu compose_generic).
template <bool M61>
__device__ __forceinline__ uint64_t compose_g(const uint32_t (&d)[8], uint32_t sh3, uint32_t nb, uint32_t mask3, const FieldParams &f) {
    const uint32_t e0 = d[0] + (d[1] << 8), e1 = d[2] + (d[3] << 8);
    const uint32_t e2 = d[4] + (d[5] << 8), e3 = d[6] + (d[7] << 8);
    if (M61) {
        const uint64_t x = pack(e0, e2) + ((uint64_t)e1 << 16);
        const uint32_t m3 = (e3 << sh3) & mask3;
        const uint32_t s3 = (e3 >> nb) + 1u;
        const uint64_t t = x + pack(s3, m3);
        uint32_t t_lo, t_hi;
        unpack(t, t_lo, t_hi);
        const uint32_t qm1 = (t_hi >> 29) - 1u;
        uint32_t r_lo, r_hi;
        unpack(t + pack(qm1, qm1), r_lo, r_hi);
        return pack(r_lo, r_hi & LOW29);
    }
    const uint64_t L = (uint64_t)e0 + ((uint64_t)e1 << 16), H = (uint64_t)e2 + ((uint64_t)e3 << 16);   // value = L + H 2^32 < 2^80
    const uint64_t lo = L + (H << 32);
    uint64_t hi = (H >> 32) + (lo < L);
    if (f.m <= (1ull << 17)) hi = reduce64_generic(f, hi);
    return reduce128_generic(f, hi, lo);
}

struct GParams {
    GShape s;
    FieldParams f;
    DrawParams dr;
};

template <int ROUNDS, bool M61>
__global__ void __launch_bounds__(CTAG)
packed_share_tcg_kernel(const int64_t *__restrict__ secrets, size_t ld, size_t dim, size_t B, uint32_t unit_begin,
                        uint32_t units_per_p, uint32_t units_total, const ChaChaKey *__restrict__ keys,
                        const uint4 *__restrict__ b_image, int64_t *__restrict__ out, unsigned *flag,
                        const __grid_constant__ GParams gp) {
    const GShape &S = gp.s;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sA = smem;                         // g tiles x 128 rows x nch chunks
    uint8_t *sB = smem + S.a_bytes;             // the constant operand
    __shared__ __align__(8) uint64_t mbar;      // full: the pass's MMAs are done
    __shared__ uint32_t tmem_base;

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t tmem_cols = (uint32_t)(S.g * S.acc_cols) < 32u ? 32u : (uint32_t)(S.g * S.acc_cols);

    if (warp == 0) {
        // the column count is a run-time value: one alloc instruction per power of two
        const uint32_t dst = smem_u32(&tmem_base);
        if (tmem_cols == 32) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" :: "r"(dst) : "memory");
        else if (tmem_cols == 64) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(dst) : "memory");
        else if (tmem_cols == 128) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(dst) : "memory");
        else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < S.b_bytes / 16; i += CTAG) reinterpret_cast<uint4 *>(sB)[i] = __ldg(b_image + i);
    // chunks nobody writes (the half of an odd last secret / draw, the pad behind the last row group) must still be
    // defined bytes: clear the operand tile once
    for (uint32_t i = tid; i < S.a_bytes / 16; i += CTAG) reinterpret_cast<uint4 *>(sA)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    const uint32_t my_taddr = taddr + ((uint32_t)(warp * 32) << 16);
    const uint32_t full_bar = smem_u32(&mbar), a_base = smem_u32(sA), b_base = smem_u32(sB);
    uint32_t parity = 0;

    const uint32_t K = (uint32_t)S.k, T = (uint32_t)S.t, N = (uint32_t)S.n, G = (uint32_t)S.g, GS = (uint32_t)S.gs;
    const uint32_t pass_batches = GS * CTAG;
    const uint32_t sh3 = 8u + (uint32_t)S.w5, nbits = 21u - (uint32_t)S.w5, mask3 = LOW29 & ~((1u << sh3) - 1u);
    const uint32_t step_p = gridDim.x / units_per_p, step_u = gridDim.x % units_per_p;
    const uint32_t unit_end = unit_begin + units_per_p;
    uint32_t p = blockIdx.x / units_per_p, u = unit_begin + blockIdx.x % units_per_p;

    for (uint32_t unit = blockIdx.x; unit < units_total; unit += gridDim.x) {
        const size_t b0 = (size_t)u * pass_batches;                 // first batch of the pass
        // ---- secrets of row `tid` of every tile: canonical, zero beyond the vector (batched.rs:38-43) ------------
        {
            const int64_t *sec = secrets + (size_t)p * ld;
            for (uint32_t q = 0; q < GS; q++) {
                const size_t e0 = (b0 + q * CTAG + tid) * K;
                uint8_t *row = sA + q * S.a_tile + (tid >> 3) * S.sbo_a + (tid & 7) * 16 + (uint32_t)S.dc * LBO;
                for (uint32_t i = 0; i < K; i++) {
                    int64_t v = e0 + i < dim ? __ldg(sec + e0 + i) : 0;
                    uint64_t x = (uint64_t)v;
                    if (v < 0) x = M61 ? canon_negative(v) : canon<false>(gp.f, v);
                    *reinterpret_cast<uint64_t *>(row + (i >> 1) * LBO + (i & 1) * 8) = x;
                }
            }
        }
        // ---- the pass's draws: blocks_per_pass keystream blocks, draw g of the pass belongs to batch g / T ----------
        {
            uint32_t k[8];
            const uint4 *src = reinterpret_cast<const uint4 *>(keys + p);
            const uint4 ka = __ldg(src), kb = __ldg(src + 1);
            k[0] = ka.x; k[1] = ka.y; k[2] = ka.z; k[3] = ka.w;
            k[4] = kb.x; k[5] = kb.y; k[6] = kb.z; k[7] = kb.w;
            const uint64_t blk0 = (uint64_t)u * S.blocks_per_pass;
            for (uint32_t blk = tid; blk < S.blocks_per_pass; blk += CTAG) {
                uint32_t w[16];
                chacha_block_g<ROUNDS>(k, blk0 + blk, w);
                uint32_t beta = (blk * 8u) / T, slot = (blk * 8u) % T;       // batch of the pass, draw index within it
                bool bad = false;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint32_t w0 = w[2 * i], w1 = w[2 * i + 1];
                    uint64_t x;
                    if (M61) {
                        // X = v + (v >> 61) == v mod (p - 1) modulo p unless v mod 2^61 >= 2^61 - 32 (packed_tc2.cu)
                        x = pack(w1, w0) + (uint64_t)(w0 >> 29);
                        bad |= (w0 & LOW29) == LOW29 && w1 >= 0xffffffe0u;
                    } else {
                        const uint64_t v = pack(w1, w0);
                        bad |= v >= gp.dr.zone;                       // rejected by gen_range: the stream shifts
                        x = reduce64_generic(gp.dr.f, v);
                    }
                    const uint32_t q = beta / CTAG, r = beta % CTAG;
                    *reinterpret_cast<uint64_t *>(sA + q * S.a_tile + (r >> 3) * S.sbo_a + (r & 7) * 16 + (slot >> 1) * LBO +
                                                  (slot & 1) * 8) = x;
                    if (++slot == T) {
                        slot = 0;
                        beta++;
                    }
                }
                if (bad) atomicOr(flag, 1u);
            }
        }
        // ---- rounds of G tiles: multiply into the G accumulators, then every thread folds and stores its rows -----------
        for (uint32_t q0 = 0; q0 < GS; q0 += G) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();          // operand rows complete (first round) / accumulators read out (later rounds)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (warp == 0) {
                uint32_t elected;
                asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
                if (elected) {
                    const uint64_t db = umma_desc(b_base, S.sbo_b);
                    for (uint32_t q = 0; q < G; q++) {
                        const uint64_t da = umma_desc(a_base + (q0 + q) * S.a_tile, S.sbo_a);
                        for (uint32_t kk = 0; kk < (uint32_t)S.nk; kk++)
                            umma_i8(taddr + q * (uint32_t)S.acc_cols, da + ((2 * LBO * kk) >> 4), db + ((2 * LBO * kk) >> 4), S.idesc, kk > 0);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(full_bar) : "memory");
                }
                __syncwarp();
            }
            mbar_wait(full_bar, parity);
            parity ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // share j of batch b at out[p][j][b]; the TMEM load of the next share runs under the fold of the current one
            for (uint32_t q = 0; q < G; q++) {
                const size_t b = b0 + (q0 + q) * CTAG + tid;
                int64_t *o = out + (size_t)p * N * B + b;
                const bool live = b < B;
                const uint32_t tq = my_taddr + q * (uint32_t)S.acc_cols;
                uint32_t d0[8], d1[8];
                tmem_ld8(tq, d0);
                for (uint32_t j = 0; j < N; j += 2) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (j + 1 < N) tmem_ld8(tq + 8 * (j + 1), d1);
                    const uint64_t r0 = compose_g<M61>(d0, sh3, nbits, mask3, gp.f);
                    if (live) o[(size_t)j * B] = (int64_t)r0;
                    if (j + 1 < N) {
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (j + 2 < N) tmem_ld8(tq + 8 * (j + 2), d0);
                        const uint64_t r1 = compose_g<M61>(d1, sh3, nbits, mask3, gp.f);
                        if (live) o[(size_t)(j + 1) * B] = (int64_t)r1;
                    }
                }
            }
        }
        // every thread has read its lanes before it reaches the next pass's barrier, after which TMEM is overwritten
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        uint32_t pn = p + step_p, un = u + step_u;
        if (un >= unit_end) {
            un -= units_per_p;
            pn++;
        }
        p = pn;
        u = un;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        if (tmem_cols == 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" :: "r"(taddr) : "memory");
        else if (tmem_cols == 64) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(taddr) : "memory");
        else if (tmem_cols == 128) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(taddr) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(taddr) : "memory");
    }
}

template <int ROUNDS, bool M61>
cudaError_t launch_g(const LaunchCtx &lc, const GParams &gp, const int64_t *secrets, size_t ld, size_t P, size_t dim,
                     size_t first_batch, size_t n_batches, const ChaChaKey *keys, const uint8_t *d_b_image, int64_t *out,
                     unsigned *flag) {
    const GShape &S = gp.s;
    const size_t pass = (size_t)S.gs * CTAG;
    const size_t B = (dim + S.k - 1) / S.k;
    if (first_batch % pass != 0 || first_batch > B) return cudaErrorInvalidValue;
    if (n_batches > B - first_batch) n_batches = B - first_batch;
    const size_t unit_begin = first_batch / pass, units_per_p = (n_batches + pass - 1) / pass, units_total = units_per_p * P;
    if (units_total == 0) return cudaSuccess;
    if ((unit_begin + units_per_p) >> 31 || units_total >> 31 || P >> 31) return cudaErrorInvalidValue;
    auto kern = packed_share_tcg_kernel<ROUNDS, M61>;
    const int tmem_cols = std::max(32, S.g * S.acc_cols);
    const size_t smem = smem_capping_residency((size_t)S.a_bytes + S.b_bytes, 512 / tmem_cols);
    // the attribute is per device and sized for the largest shape: set it to the maximum once
    static KernelSetup setup;
    int regs = 0;
    size_t static_smem = 0;
    const cudaError_t se = setup_kernel(setup, kern, 200 * 1024, &regs, &static_smem);
    if (se != cudaSuccess) return se;
    const int per_sm = resident_ctas(regs, CTAG, smem, static_smem, tmem_cols);
    size_t grid = (size_t)lc.sm_count * per_sm;
    if (grid > units_total) grid = units_total;
    kern<<<(unsigned)grid, CTAG, smem, lc.stream>>>(secrets, ld, dim, B, (uint32_t)unit_begin, (uint32_t)units_per_p,
                                                    (uint32_t)units_total, keys, reinterpret_cast<const uint4 *>(d_b_image),
                                                    out, flag, gp);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

}  // namespace

bool packed_share_tcg_supported(int k, int t, int n) { return k >= 1 && t >= 1 && k + t <= 16 && n >= 1 && n <= 32; }

size_t packed_share_tcg_image_bytes(int k, int t, int n) { return make_shape(k, t, n, true).b_bytes; }

size_t packed_share_tcg_slice_batches(int k, int t, int n) { return (size_t)make_shape(k, t, n, true).gs * CTAG; }

// the constant operand as it lies in shared memory (limb plan of the Mersenne path when p = 2^61 - 1, plain bytes otherwise)
void packed_share_tcg_build_image(int k, int t, int n, const Matrix &m, uint64_t p, uint8_t *img) {
    typedef unsigned __int128 u128;
    const bool m61 = p == P61;
    const GShape S = make_shape(k, t, n, m61);
    const int w[8] = {8, 8, 8, 8, 8, S.w5, 8, 13 - S.w5 + (m61 ? 0 : 3)};        // plain bytes: 8 x 8
    int pos[8];
    pos[0] = 0;
    for (int s = 1; s < 8; s++) pos[s] = pos[s - 1] + w[s - 1];
    memset(img, 0, S.b_bytes);
    for (int j = 0; j < n; j++)
        for (int c = 0; c < S.nch; c++)
            for (int v = 0; v < 2; v++) {
                int xi;                                        // index into x = [secrets ; randomness]
                if (c < S.dc) {
                    const int idx = 2 * c + v;
                    if (idx >= t) continue;
                    xi = k + idx;
                } else {
                    const int idx = 2 * (c - S.dc) + v;
                    if (idx >= k) continue;
                    xi = idx;
                }
                for (int byte = 0; byte < 8; byte++) {
                    const uint64_t cst = (uint64_t)((u128)m.e[j * (k + t) + xi] * ((((u128)1) << (8 * byte)) % p) % p);
                    for (int s = 0; s < 8; s++) {
                        const int col = j * 8 + s;
                        img[(col / 8) * S.sbo_b + c * LBO + (col % 8) * 16 + v * 8 + byte] =
                            (uint8_t)((cst >> pos[s]) & ((1u << w[s]) - 1u));
                    }
                }
            }
}

cudaError_t launch_packed_share_tcg(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, int rounds, int k, int t,
                                    int n, const int64_t *secrets, size_t ld, size_t P, size_t dim, size_t first_batch,
                                    size_t n_batches, const ChaChaKey *keys, const uint8_t *d_b_image, int64_t *shares_out,
                                    unsigned *flag) {
    if (!packed_share_tcg_supported(k, t, n)) return cudaErrorInvalidValue;
    const bool m61 = f.kind == FIELD_MERSENNE61;
    GParams gp{make_shape(k, t, n, m61), f, dr};
    *lc.kernel_name = m61 ? "packed_share<run-time shape>/mersenne61 tcgen05.mma.kind::i8"
                          : "packed_share<run-time shape>/any prime tcgen05.mma.kind::i8";
#define SDA_LG(R, M) return launch_g<R, M>(lc, gp, secrets, ld, P, dim, first_batch, n_batches, keys, d_b_image, shares_out, flag)
    if (m61) {
        if (rounds == 8) SDA_LG(8, true);
        if (rounds == 12) SDA_LG(12, true);
        SDA_LG(20, true);
    }
    if (rounds == 8) SDA_LG(8, false);
    if (rounds == 12) SDA_LG(12, false);
    SDA_LG(20, false);
#undef SDA_LG
}

}  // namespace sda
