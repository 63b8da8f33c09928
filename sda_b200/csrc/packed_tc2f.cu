// packed_tc2f.cu -- share generation fused with the clerk sums (participate.rs:75-76 + clerk.rs:71-86 on one box) on the
// paired-tile machinery of packed_tc2.cuh: the second generation of packed_tc.cu's packed_share_combine_tc_kernel.
//
// A CTA owns one pass of batches (256 of them: one E/O tile pair) and walks the participants: for each it
// stages that participant's secrets and draws of the pass as operand tiles exactly as the share-generation kernel does and
// multiplies them into the SAME TMEM accumulators (tcgen05.mma with accumulate).  No share is ever folded or stored per
// participant: the limb sums keep growing (at most 8 (k + t) 255^2 per participant and limb, so 256 participants fit the
// s32 accumulators), are drained every 256 participants and at the end -- composed with 64-bit arithmetic and added to the
// running sums of the pass, which live in the output between drains.  Per participant a thread's work is its share of the
// keystream (one block for t = 4), the staging of 2k secrets and one barrier: the kernel is the keystream plus ~12 %
// (profiles/r02_fused.md).
//
// A pass is ONE tile pair (256 batches: two 64-column accumulators, four CTAs per SM by TMEM for n <= 8); where that leaves a
// participant's pass fewer keystream blocks than the CTA has threads (t = 2: 64), GP = 2 participants are staged and
// multiplied per step, half of the warps drawing for each.
//
// p = 2^61 - 1, the shapes BASELINE names, any round count.  Same rejection contract as the other kernels: a gen_range
// rejection raises `flag` and the caller redoes the call on the materialising path.
#include "packed_tc2.cuh"

namespace sda {

namespace {

constexpr int MAX_ACCUM2 = 256;       // participants per TMEM accumulation (packed_tc.cu: 256 * 96 * 255^2 < 2^31)

// limb sums d[s] < 2^31 at the limb plan's positions -> canonical value mod p (runs once per 256 participants)
template <int W5>
__device__ __forceinline__ uint64_t compose_wide2(const uint32_t (&d)[8]) {
    // positions 0, 8, 16, 24 | 32 + (0, 8, 8 + W5, 16 + W5)
    const uint64_t lo = (uint64_t)d[0] + ((uint64_t)d[1] << 8) + ((uint64_t)d[2] << 16) + ((uint64_t)d[3] << 24);   // < 2^56
    const uint64_t hi = (uint64_t)d[4] + ((uint64_t)d[5] << 8) + ((uint64_t)d[6] << (8 + W5)) + ((uint64_t)d[7] << (16 + W5));   // < 2^56
    // hi 2^32 = (hi mod 2^29) 2^32 + (hi >> 29) 2^61 == (hi mod 2^29) 2^32 + (hi >> 29)
    uint64_t v = lo + ((hi & LOW29) << 32) + (hi >> 29);       // < 2^56 + 2^61 + 2^27
    v = (v & P61) + (v >> 61);
    return v >= P61 ? v - P61 : v;
}

// the NK MMAs of one tile; `fresh`: the first of them overwrites the accumulator
template <class S>
__device__ __forceinline__ void issue_tile_acc(uint32_t taddr, uint32_t d_tile, uint32_t s_tile, uint32_t b_img, bool fresh) {
    const uint64_t dd = umma_desc(d_tile, S::SBO_D), ds = umma_desc(s_tile, S::SBO_S), db = umma_desc(b_img, S::SBO_B);
#pragma unroll
    for (int kk = 0; kk < S::NKD; kk++)
        umma_i8(taddr, dd + ((2 * LBO * kk) >> 4), db + ((2 * LBO * kk) >> 4), S::IDESC, (kk > 0 || !fresh) ? 1u : 0u);
#pragma unroll
    for (int kk = 0; kk < S::NKS; kk++)
        umma_i8(taddr, ds + ((2 * LBO * kk) >> 4), db + ((2 * LBO * (S::NKD + kk)) >> 4), S::IDESC, 1);
}

template <int K, int T, int N>
struct FusedShape2 {
    typedef Shape2<K, T, N, 1> S;                                       // one tile pair per pass
    static constexpr int GP = S::NBLK < CTA2 ? CTA2 / S::NBLK : 1;       // participants per step
    static_assert(GP * S::NBLK >= CTA2 || S::NBLK % CTA2 == 0 || GP == 1, "keystream blocks per step");
    static constexpr int TCOLS = 2 * S::ACC_COLS;                        // accumulators E and O
    static_assert((long long)MAX_ACCUM2 * 8 * (K + T) * 255 * 255 < (1ll << 31), "limb sums of a full accumulation fit s32");
    static constexpr uint32_t D_BYTES = GP * S::D_BYTES, S_BYTES = GP * S::S_BYTES, IN_BYTES = GP * S::IN_BYTES;
    static constexpr uint32_t SMEM = 2 * D_BYTES + S_BYTES + 2 * S::B_IMG + IN_BYTES;
    static constexpr int BY_SMEM = SMEM + 1152 > 227 * 1024 / 2 ? 1 : SMEM + 1152 > 227 * 1024 / 3 ? 2 : SMEM + 1152 > 227 * 1024 / 4 ? 3 : 4;
    static constexpr int RESIDENT = 512 / TCOLS < BY_SMEM ? 512 / TCOLS : BY_SMEM;
};

template <int K, int T, int N, int ROUNDS>
__global__ void __launch_bounds__(CTA2, FusedShape2<K, T, N>::RESIDENT)
packed_share_combine_tc2_kernel(const int64_t *__restrict__ secrets, size_t ld, size_t dim, size_t B, uint32_t P,
                                uint32_t units, uint32_t full_in_units,
                                const ChaChaKey *__restrict__ keys, const ChaChaPre *__restrict__ pres,
                                const uint4 *__restrict__ b_image, const int64_t *acc_in, int64_t *out,   // acc_in may equal out
                                unsigned *flag, int bulk_ok) {
    typedef FusedShape2<K, T, N> F;
    typedef typename F::S S;
    constexpr int GP = F::GP;
    static_assert(S::PAIRS == 1, "one tile pair per pass");
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sD = smem;                                    // 2 x GP x ({E, O} tiles x 128 rows x draws)
    uint8_t *sS = smem + 2 * F::D_BYTES;                   // GP x ({E, O} tiles x 128 rows x secrets)
    uint8_t *sB = sS + F::S_BYTES;                         // the constant operand, E image then O image
    int64_t *sIn = reinterpret_cast<int64_t *>(sB + 2 * S::B_IMG);   // the coming GP participants' raw secrets of the pass
    __shared__ __align__(8) uint64_t mbar[2];              // [0] full (MMAs done), [1] secrets landed
    __shared__ uint32_t tmem_base;

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    // keystream duty of this thread inside a step: participant `my_g` of the step, thread `my_t` of its CTA2 / GP drawers
    const int my_g = tid / (CTA2 / GP), my_t = tid % (CTA2 / GP);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base)), "n"(F::TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < 2 * S::B_IMG / 16; i += CTA2) reinterpret_cast<uint4 *>(sB)[i] = __ldg(b_image + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    const uint32_t my_taddr = taddr + ((uint32_t)(warp * 32) << 16);
    const uint32_t full_bar = smem_u32(&mbar[0]), landed_bar = smem_u32(&mbar[1]);
    const uint32_t d_base = smem_u32(sD), s_base = smem_u32(sS), b_base = smem_u32(sB), sin_addr = smem_u32(sIn);
    uint32_t parity = 0, landed_parity = 0, buf = 0;

    // the raw secrets of participants p0 .. p0 + np - 1 of pass u into sIn (one bulk copy each, or the threads' own loads)
    auto load_group = [&](uint32_t p0, uint32_t np, uint32_t u, bool by_bulk) {
        if (by_bulk) {
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(landed_bar), "r"(np * S::IN_BYTES) : "memory");
                for (uint32_t g = 0; g < np; g++) {
                    const int64_t *src = secrets + (size_t)(p0 + g) * ld + (size_t)u * (S::PASS * K);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(sin_addr + g * S::IN_BYTES), "l"(src), "r"(S::IN_BYTES), "r"(landed_bar) : "memory");
                }
            }
        } else {
            for (uint32_t g = 0; g < np; g++)
                fill_secrets2<S, K>(secrets, ld, dim, p0 + g, u, tid, sIn + g * (S::IN_BYTES / 8));
        }
    };
    // this thread's part of the keystream of participants p0 .. of pass u into draw buffer `b`
    // (the key is requested a step ahead: its load is the only global-memory latency on a step's critical path)
    auto draw_group = [&](const KeyRegs &kr, uint32_t np, uint32_t u, uint32_t b) {
        if ((uint32_t)my_g < np) stage_draws2<S, ROUNDS>(kr, u, my_t, sD + b * F::D_BYTES + my_g * S::D_BYTES, flag);
    };
    auto key_of = [&](uint32_t p) { return load_keys2(keys, pres, p < P ? p : P - 1); };

    for (uint32_t u = blockIdx.x; u < units; u += gridDim.x) {
        const bool by_bulk = bulk_ok != 0 && u < full_in_units;     // the pass lies inside the vectors, sources 16-byte aligned
        // this thread's batches: 2 tid and 2 tid + 1 of the pair.  The running sums live in `out` between drains (a drain every
        // 256 participants: their traffic is nothing next to the keystream), the first drain starts from acc_in
        const size_t pass_first = (size_t)u * S::PASS;
        bool drained_before = false;
        const uint32_t np0 = P < (uint32_t)GP ? P : (uint32_t)GP;
        load_group(0, np0, u, by_bulk);
        draw_group(key_of((uint32_t)my_g), np0, u, buf);
        uint32_t in_tmem = 0;
        for (uint32_t p = 0; p < P; p += GP) {
            const uint32_t np = P - p < (uint32_t)GP ? P - p : (uint32_t)GP;           // participants of this step
            const uint32_t pn = p + GP;
            const bool more = pn < P;
            const uint32_t npn = more ? (P - pn < (uint32_t)GP ? P - pn : (uint32_t)GP) : 0;
            KeyRegs knext;
            if (more) knext = key_of(pn + (uint32_t)my_g);
            if (by_bulk) {
                mbar_wait(landed_bar, landed_parity);
                landed_parity ^= 1;
            }
#pragma unroll
            for (int g = 0; g < GP; g++) {
                if ((uint32_t)g < np) {
                    uint4 v[K];
                    const uint4 *row = reinterpret_cast<const uint4 *>(sIn + g * (S::IN_BYTES / 8) + (2 * tid) * K);
#pragma unroll
                    for (int i = 0; i < K; i++) v[i] = row[i];
                    uint32_t sign = 0;
#pragma unroll
                    for (int i = 0; i < K; i++) sign |= v[i].y | v[i].w;
                    if ((int32_t)sign < 0) {
#pragma unroll
                        for (int i = 0; i < K; i++) {
                            canon_pair2(v[i].x, v[i].y);
                            canon_pair2(v[i].z, v[i].w);
                        }
                    }
                    uint8_t *te = sS + g * S::S_BYTES + (tid >> 3) * S::SBO_S + (tid & 7) * 16;
#pragma unroll
                    for (int c = 0; c < S::SC; c++) {
                        *reinterpret_cast<uint4 *>(te + c * LBO) = v[c];                          // E: words 0 .. SC-1
                        *reinterpret_cast<uint4 *>(te + S::S_TILE + c * LBO) = v[K - S::SC + c];  // O: words K-SC .. K-1
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (warp == 0) {
                if (elect_one()) {
                    for (uint32_t g = 0; g < np; g++) {
                        const uint32_t dg = d_base + buf * F::D_BYTES + g * S::D_BYTES, sg = s_base + g * S::S_BYTES;
                        issue_tile_acc<S>(taddr, dg, sg, b_base, in_tmem + g == 0);
                        issue_tile_acc<S>(taddr + S::ACC_COLS, dg + S::D_TILE, sg + S::S_TILE, b_base + S::B_IMG, in_tmem + g == 0);
                    }
                    commit2(full_bar);
                }
                __syncwarp();
            }
            in_tmem += np;
            // the next step's secrets (everyone is past the barrier, i.e. has read this step's) and keystream under the MMAs
            if (more) {
                load_group(pn, npn, u, by_bulk);
                draw_group(knext, npn, u, buf ^ 1);
            }
            mbar_wait(full_bar, parity);               // tiles consumed: the next step may overwrite the staged secrets
            parity ^= 1;
            buf ^= 1;
            if (in_tmem + GP > MAX_ACCUM2 || !more) {  // drain before the accumulators fill up, and at the end
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int64_t *src = drained_before ? out : acc_in;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const size_t b = pass_first + (size_t)(2 * tid + h);
                    uint32_t d[N][8];
                    tmem_ld_shares<N>(my_taddr + h * S::ACC_COLS, d);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (b < B) {
#pragma unroll
                        for (int j = 0; j < N; j++) {
                            int64_t a = src != nullptr ? src[(size_t)j * B + b] : 0;
                            if (a < 0) a = (int64_t)canon_negative(a);
                            const uint64_t a0 = (uint64_t)a >= P61 ? (((uint64_t)a & P61) + ((uint64_t)a >> 61)) % P61 : (uint64_t)a;
                            const uint64_t x = a0 + compose_wide2<S::W5>(d[j]);
                            out[(size_t)j * B + b] = (int64_t)(x >= P61 ? x - P61 : x);
                        }
                    }
                }
                drained_before = true;
                in_tmem = 0;
                // the next MMA (fresh accumulators) is issued after the next staging barrier, which every thread reaches
                // only after these loads
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
        }
        // sIn, sS and the draw buffers are reused by the next pass: every thread is past its reads (the last full wait),
        // and the first staging barrier of the next pass orders the rest
        __syncthreads();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(F::TCOLS) : "memory");
}

template <int K, int T, int N, int ROUNDS>
cudaError_t launch_fused2(const LaunchCtx &lc, const int64_t *secrets, size_t ld, size_t P, size_t dim, const ChaChaKey *keys,
                          uint32_t *d_pre, const uint8_t *d_b_image, const int64_t *acc_in, int64_t *out, unsigned *flag) {
    typedef FusedShape2<K, T, N> F;
    typedef typename F::S S;
    constexpr int TCOLS = F::TCOLS;
    const size_t B = (dim + K - 1) / K;
    const size_t units = (B + S::PASS - 1) / S::PASS;
    if (units == 0 || P == 0) return cudaSuccess;
    if (units >> 31 || P >> 31 || (B * (size_t)T + 7) / 8 >> 32) return cudaErrorInvalidValue;
    auto kern = packed_share_combine_tc2_kernel<K, T, N, ROUNDS>;
    const size_t smem = smem_capping_residency(F::SMEM, 512 / TCOLS);
    static KernelSetup setup;
    int regs = 0;
    size_t static_smem = 0;
    const cudaError_t se = setup_kernel(setup, kern, smem, &regs, &static_smem);
    if (se != cudaSuccess) return se;
    const int per_sm = resident_ctas(regs, CTA2, smem, static_smem, TCOLS);
    const size_t grid = std::min<size_t>(units, (size_t)lc.sm_count * per_sm);
    const int bulk_ok = reinterpret_cast<uintptr_t>(secrets) % 16 == 0 && (ld % 2 == 0 || P == 1);
    const size_t full_in = dim / ((size_t)S::PASS * K);
    ChaChaPre *pres = reinterpret_cast<ChaChaPre *>(d_pre);
    chacha_prepare_kernel<<<(unsigned)((P + 127) / 128), 128, 0, lc.stream>>>(keys, P, pres);
    ++*lc.nlaunch;
    kern<<<(unsigned)grid, CTA2, smem, lc.stream>>>(secrets, ld, dim, B, (uint32_t)P, (uint32_t)units,
                                                    (uint32_t)std::min<size_t>(full_in, 0xffffffffu),
                                                    keys, pres,
                                                    reinterpret_cast<const uint4 *>(d_b_image), acc_in, out, flag, bulk_ok);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

}  // namespace

#define SDA_TC2F_SHAPES(X) X(3, 2, 5) X(5, 4, 9) X(3, 4, 7) X(3, 4, 8)

bool packed_share_combine_tc2_supported(int k, int t, int n, size_t dim) {
    const size_t B = (dim + (size_t)k - 1) / (size_t)k;
    if ((B * (size_t)t + 7) / 8 >> 32) return false;
#define X(K, T, N) if (k == K && t == T && n == N) return true;
    SDA_TC2F_SHAPES(X)
#undef X
    return false;
}

// fused share generation + clerk accumulation over the participants: out[n][B] = acc_in + sum_p shares(p).
// d_key_scratch: packed_share_tc2_key_scratch_bytes(P); d_b_image: packed_share_tc2_build_image's two images.
cudaError_t launch_packed_share_combine_tc2(const LaunchCtx &lc, int rounds, int k, int t, int n, const int64_t *secrets,
                                            size_t ld, size_t P, size_t dim, const ChaChaKey *keys, uint32_t *d_key_scratch,
                                            const uint8_t *d_b_image, const int64_t *acc_in, int64_t *out, unsigned *flag) {
#define X(K, T, N)                                                                                                        \
    if (k == K && t == T && n == N) {                                                                                     \
        *lc.kernel_name = "packed_share_combine<" #K "," #T "," #N ">/mersenne61 tcgen05.mma.kind::i8, paired tiles, TMEM-accumulated"; \
        if (rounds == 8) return launch_fused2<K, T, N, 8>(lc, secrets, ld, P, dim, keys, d_key_scratch, d_b_image, acc_in, out, flag); \
        if (rounds == 12) return launch_fused2<K, T, N, 12>(lc, secrets, ld, P, dim, keys, d_key_scratch, d_b_image, acc_in, out, flag); \
        return launch_fused2<K, T, N, 20>(lc, secrets, ld, P, dim, keys, d_key_scratch, d_b_image, acc_in, out, flag);   \
    }
    SDA_TC2F_SHAPES(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace sda
