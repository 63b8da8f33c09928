// reveal_tc.cu -- recipient-side packed-Shamir reconstruction over p = 2^61 - 1 on the tensor cores.
//
//   client/src/crypto/sharing/batched.rs:68-97 + packed_shamir.rs:73-77 -> tss 0.2 `reconstruct`:
//   per batch, Newton interpolation through (1, 0) and the present clerks' points, evaluated at the
//   secret points.  The map is linear and depends only on the index set, so the host builds
//   R (k x m') once (api.cu) and the kernel evaluates  secrets_b = R . shares_b  for every batch.
//
// Same construction as packed_tc.cu: the product is linear in the bytes of the shares, so with
//     A[b][(s,c)]      = byte c of clerk s's combined share of batch b      (u8, straight from memory)
//     B[(e,l)][(s,c)]  = byte l of (R[e][s] 2^{8c} mod p)                   (u8, constant)
// one tcgen05.mma.kind::i8 per 32 bytes of row leaves the eight limb sums of every secret of 128
// batches in TMEM, and the thread that owns a batch composes its k secrets and writes them out
// (truncated to `dimension`, batched.rs:94).  Shapes are runtime values (any k <= 16, m' <= 16):
// one kernel per TMEM allocation size.  Shares may be any i64 (negative ones are canonicalised while
// staging); outputs are canonical.
#include <algorithm>
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"

namespace sda {

namespace {

using namespace tc;

constexpr int CTA = 128;

struct RevealShape {
    int k, m;            // secrets per batch, present clerks
    int chunks;          // 16-byte chunks per row = ceil(m / 2)
    int nk;              // MMAs per tile
    int n_mma;           // 8 k rounded up to 16
    uint32_t sbo_a, a_bytes, sbo_b, b_bytes;
    int tmem_cols;
};

RevealShape make_shape(int k, int m) {
    RevealShape s;
    s.k = k;
    s.m = m;
    s.chunks = (m + 1) / 2;
    s.nk = (s.chunks + 1) / 2;
    s.n_mma = (8 * k + 15) / 16 * 16;
    s.sbo_a = (uint32_t)s.chunks * 128;
    s.a_bytes = 16 * s.sbo_a + 128;        // + the aliased chunk when `chunks` is odd
    s.sbo_b = 2u * s.nk * 128;
    s.b_bytes = (uint32_t)(s.n_mma / 8) * s.sbo_b;
    s.tmem_cols = s.n_mma <= 32 ? 32 : s.n_mma <= 64 ? 64 : 128;
    return s;
}

template <int TMEM_COLS>
__global__ void __launch_bounds__(CTA)
reveal_tc_kernel(const int64_t *__restrict__ shares, size_t ld, size_t nbatches, size_t dimension, int k, int m, int chunks,
                 int nk, uint32_t sbo_a, uint32_t a_bytes, uint32_t sbo_b, uint32_t b_bytes, uint32_t idesc,
                 const uint4 *__restrict__ b_image, int64_t *__restrict__ out, uint32_t two16, int bulk_ok) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sA = smem;
    uint8_t *sB = smem + ((a_bytes + 127) & ~127u);
    // a tile's 128 k secrets are contiguous in the output: composed into shared memory, stored by one bulk copy
    uint64_t *sOut = reinterpret_cast<uint64_t *>(sB + ((b_bytes + 127) & ~127u));
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < b_bytes / 16; i += CTA) reinterpret_cast<uint4 *>(sB)[i] = __ldg(b_image + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    const uint32_t my_taddr = taddr + ((uint32_t)(warp * 32) << 16);
    const uint32_t bar = smem_u32(&mbar), a_base = smem_u32(sA), b_base = smem_u32(sB);
    uint8_t *my_row = sA + (tid >> 3) * sbo_a + (tid & 7) * 16;
    uint32_t parity = 0;

    const size_t tiles = (nbatches + CTA - 1) / CTA;
    // this thread's row of the coming tile -- the m shares of its batch, two per 16-byte chunk (batched.rs:83-85).
    // Where TMEM leaves room for at most 8 CTAs per SM the row is loaded a tile ahead, in flight under the MMA and the
    // compose of the current tile (9 x [2M] shares: 90 -> 76 us); with 32 columns per CTA the 16 resident CTAs hide
    // the loads better than the 9 that the prefetch registers would leave (7 x [3.33M]: 81 us against 103 us).
    constexpr bool PREFETCH = TMEM_COLS >= 64;
    constexpr int MAX_CHUNKS = 8;                      // m' <= 16
    int64_t v[2 * MAX_CHUNKS];
    auto load_row = [&](size_t tile) {
        const size_t b = tile * CTA + tid;
#pragma unroll
        for (int c = 0; c < MAX_CHUNKS; c++) {
            v[2 * c] = v[2 * c + 1] = 0;
            if (c < chunks && b < nbatches) {
                v[2 * c] = __ldg(shares + (size_t)(2 * c) * ld + b);
                if (2 * c + 1 < m) v[2 * c + 1] = __ldg(shares + (size_t)(2 * c + 1) * ld + b);
            }
        }
    };
    if (PREFETCH && blockIdx.x < tiles) load_row(blockIdx.x);
    for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const size_t b = tile * CTA + tid;
        auto stage_chunk = [&](int c, int64_t v0, int64_t v1) {
            if (v0 < 0) v0 = (int64_t)canon_negative(v0);
            if (v1 < 0) v1 = (int64_t)canon_negative(v1);
            uint32_t al, ah, bl, bh;
            unpack((uint64_t)v0, al, ah);
            unpack((uint64_t)v1, bl, bh);
            *reinterpret_cast<uint4 *>(my_row + c * LBO) = make_uint4(al, ah, bl, bh);
        };
        if constexpr (PREFETCH) {
#pragma unroll
            for (int c = 0; c < MAX_CHUNKS; c++)
                if (c < chunks) stage_chunk(c, v[2 * c], v[2 * c + 1]);
        } else {
            for (int c = 0; c < chunks; c++) {
                int64_t v0 = 0, v1 = 0;
                if (b < nbatches) {
                    v0 = __ldg(shares + (size_t)(2 * c) * ld + b);
                    if (2 * c + 1 < m) v1 = __ldg(shares + (size_t)(2 * c + 1) * ld + b);
                }
                stage_chunk(c, v0, v1);
            }
        }
        // the previous tile's bulk store has read its staging buffer before anyone passes the barrier below
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t da = umma_desc(a_base, sbo_a), db = umma_desc(b_base, sbo_b);
            for (int kk = 0; kk < nk; kk++)
                umma_i8(taddr, da + ((2 * LBO * kk) >> 4), db + ((2 * LBO * kk) >> 4), idesc, kk > 0);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
        }
        if (PREFETCH && tile + gridDim.x < tiles) load_row(tile + gridDim.x);
        mbar_wait(bar, parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const size_t o0 = b * (size_t)k;
        const size_t tile_first = tile * (size_t)(CTA * k);
        const bool by_bulk = bulk_ok != 0 && tile_first + (size_t)(CTA * k) <= dimension;     // a whole tile inside the vector
        uint32_t d[8], dn[8];
        tmem_ld8(my_taddr, d);
        for (int e = 0; e < k; e++) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (e + 1 < k) tmem_ld8(my_taddr + 8 * (e + 1), dn);      // the next secret's limbs, in flight under this compose
            const uint64_t r = compose(d, two16);
            if (by_bulk) sOut[tid * k + e] = r;
            else if (b < nbatches && o0 + e < dimension) out[o0 + e] = (int64_t)r;      // batched.rs:94 truncate
#pragma unroll
            for (int i = 0; i < 8; i++) d[i] = dn[i];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();          // TMEM and the rows are free for the next tile; the staged secrets are complete
        if (by_bulk && tid == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(out + tile_first), "r"(smem_u32(sOut)), "r"((uint32_t)(CTA * k * 8)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // shared memory outlives the last store
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(TMEM_COLS) : "memory");
}

template <int TMEM_COLS>
cudaError_t launch(const LaunchCtx &lc, const RevealShape &s, const int64_t *shares, size_t ld, size_t nbatches,
                   size_t dimension, const uint8_t *d_b_image, int64_t *out) {
    auto kern = reveal_tc_kernel<TMEM_COLS>;
    // never more CTAs on an SM than can hold their TMEM columns (tc_common.cuh): this kernel needs few registers
    // and little shared memory, so without the floor 12 CTAs land on an SM that has columns for 512 / TMEM_COLS
    const size_t smem = smem_capping_residency(((s.a_bytes + 127) & ~127u) + ((s.b_bytes + 127) & ~127u) + (size_t)CTA * s.k * 8,
                                               512 / TMEM_COLS);
    const int bulk_ok = reinterpret_cast<uintptr_t>(out) % 16 == 0;      // bulk stores need 16-byte aligned destinations
    static KernelSetup setup;                              // one per TMEM size (this function is a template)
    int regs = 0;
    size_t static_smem = 0;
    const cudaError_t se = setup_kernel(setup, kern, 100 * 1024, &regs, &static_smem);    // any shape's request fits
    if (se != cudaSuccess) return se;
    const int per_sm = resident_ctas(regs, CTA, smem, static_smem, TMEM_COLS);
    const size_t tiles = (nbatches + CTA - 1) / CTA;
    const size_t grid = std::min<size_t>(tiles, (size_t)lc.sm_count * per_sm);
    kern<<<(unsigned)grid, CTA, smem, lc.stream>>>(shares, ld, nbatches, dimension, s.k, s.m, s.chunks, s.nk, s.sbo_a,
                                                   s.a_bytes, s.sbo_b, s.b_bytes, idesc_u8(s.n_mma),
                                                   reinterpret_cast<const uint4 *>(d_b_image), out, 65536u, bulk_ok);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

}  // namespace

bool reveal_tc_supported(int k, int m) { return k >= 1 && k <= 16 && m >= 1 && m <= 16; }

size_t reveal_tc_image_bytes(int k, int m) { return reveal_tc_supported(k, m) ? make_shape(k, m).b_bytes : 0; }

// B[(e,l)][(s,c)] = byte l of (R[e][s] 2^{8c} mod p), in the shared-memory operand layout
void reveal_tc_build_image(int k, int m, const Matrix &R, uint8_t *img) {
    typedef unsigned __int128 u128;
    const RevealShape sh = make_shape(k, m);
    memset(img, 0, sh.b_bytes);
    for (int e = 0; e < k; e++)
        for (int s = 0; s < m; s++)
            for (int byte = 0; byte < 8; byte++) {
                const uint64_t cst = (uint64_t)((u128)R.e[e * m + s] * ((((u128)1) << (8 * byte)) % P61) % P61);
                for (int l = 0; l < 8; l++) {
                    const int n = e * 8 + l;
                    img[(n / 8) * sh.sbo_b + (s / 2) * LBO + (n % 8) * 16 + (s % 2) * 8 + byte] = (uint8_t)(cst >> (8 * l));
                }
            }
}

cudaError_t launch_reveal_tc(const LaunchCtx &lc, int k, int m, const int64_t *shares, size_t ld, size_t dimension,
                             const uint8_t *d_b_image, int64_t *secrets_out) {
    if (dimension == 0) return cudaSuccess;
    const RevealShape s = make_shape(k, m);
    const size_t nbatches = (dimension + k - 1) / k;
    if (s.tmem_cols == 32) return launch<32>(lc, s, shares, ld, nbatches, dimension, d_b_image, secrets_out);
    if (s.tmem_cols == 64) return launch<64>(lc, s, shares, ld, nbatches, dimension, d_b_image, secrets_out);
    return launch<128>(lc, s, shares, ld, nbatches, dimension, d_b_image, secrets_out);
}

}  // namespace sda
