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
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"

namespace sda {

namespace {

using namespace tc;

constexpr int CTA = 128;

// CH = 16-byte chunks per row = ceil(m' / 2) is a template parameter: the loads, the staging and the MMA issue of a tile
// are straight-line code (the run-time-shaped first version spent 511 instructions per warp and tile, most of them
// predicates and branches of its shape loops; profiles/r02_reveal.md).  k, the parity of m' and the TMEM allocation stay
// run-time values.
template <int CH>
struct RevealShape {
    static constexpr int NK = (CH + 1) / 2;                       // MMAs per tile (32 bytes of K each)
    static constexpr uint32_t SBO_A = CH * 128;
    static constexpr uint32_t A_BYTES = 16 * SBO_A + 128;         // + the aliased chunk when CH is odd
    static constexpr uint32_t SBO_B = 2 * NK * 128;
    static constexpr int W5 = w5_for(2 * CH);                     // limb plan of up to 2 CH shares per row (tc_common.cuh)
};
inline int n_mma_for(int k) { return (8 * k + 15) / 16 * 16; }
inline int tmem_cols_for(int k) { return n_mma_for(k) <= 32 ? 32 : n_mma_for(k) <= 64 ? 64 : 128; }

// PREFETCH: this thread's row of the coming tile is loaded a tile ahead, in flight under the MMA and the compose of the
// current one.  Used where TMEM leaves room for at most 8 CTAs per SM (k >= 5); with 32 columns per CTA the 16 resident
// CTAs hide the loads themselves and the prefetch registers would only lower their number.
template <int CH, bool PREFETCH>
// (measured: prefetching with 32 columns changes nothing, holding the registers to 32 for 16 CTAs costs 12 %)
__global__ void __launch_bounds__(CTA, PREFETCH || CH > 4 ? 8 : 12)
reveal_tc_kernel(const int64_t *__restrict__ shares, size_t ld, size_t nbatches, size_t dimension, int k, int m_odd,
                 uint32_t tmem_cols, uint32_t b_bytes, uint32_t idesc, const uint4 *__restrict__ b_image,
                 int64_t *__restrict__ out, uint32_t two16, int bulk_ok) {
    typedef RevealShape<CH> S;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sA = smem;
    uint8_t *sB = smem + ((S::A_BYTES + 127) & ~127u);
    // a tile's 128 k secrets are contiguous in the output: composed into shared memory, stored by one bulk copy
    uint64_t *sOut = reinterpret_cast<uint64_t *>(sB + ((b_bytes + 127) & ~127u));
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base)), "r"(tmem_cols) : "memory");
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
    const uint32_t bar = smem_u32(&mbar), a_base = smem_u32(sA), b_base = smem_u32(sB), out_base = smem_u32(sOut);
    uint8_t *my_row = sA + (tid >> 3) * S::SBO_A + (tid & 7) * 16;
    uint64_t *my_out = sOut + tid * k;
    uint32_t parity = 0;

    const uint32_t tiles = (uint32_t)((nbatches + CTA - 1) / CTA);       // the launcher keeps it below 2^32
    const uint32_t full_in = (uint32_t)(nbatches / CTA);                 // tiles whose 128 batches all exist
    const uint32_t full_out = (uint32_t)min((unsigned long long)(dimension / ((size_t)CTA * k)), 0xffffffffull);   // ... whose 128 k secrets do
    // this thread's row of a tile -- the m' shares of its batch, two per 16-byte chunk (batched.rs:83-85)
    int64_t v[2 * CH];
    auto load_row = [&](uint32_t tile) {
        const int64_t *src = shares + ((size_t)tile * CTA + tid);
        const bool live = tile < full_in || (size_t)tile * CTA + tid < nbatches;
#pragma unroll
        for (int i = 0; i < 2 * CH; i++) {
            const bool present = live && (i < 2 * CH - 1 || !m_odd);
            v[i] = present ? __ldg(src) : 0;
            src += ld;
        }
    };
    if (PREFETCH && blockIdx.x < tiles) load_row(blockIdx.x);
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        if (!PREFETCH) load_row(tile);
        // negative representatives are rare: one test per row
        uint32_t sign = 0;
#pragma unroll
        for (int i = 0; i < 2 * CH; i++) sign |= (uint32_t)((uint64_t)v[i] >> 32);
        if ((int32_t)sign < 0) {
#pragma unroll
            for (int i = 0; i < 2 * CH; i++)
                if (v[i] < 0) v[i] = (int64_t)canon_negative(v[i]);
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
            uint32_t al, ah, bl, bh;
            unpack((uint64_t)v[2 * c], al, ah);
            unpack((uint64_t)v[2 * c + 1], bl, bh);
            *reinterpret_cast<uint4 *>(my_row + c * LBO) = make_uint4(al, ah, bl, bh);
        }
        // the previous tile's bulk store has read its staging buffer before anyone passes the barrier below
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint64_t da = umma_desc(a_base, S::SBO_A), db = umma_desc(b_base, S::SBO_B);
#pragma unroll
                for (int kk = 0; kk < S::NK; kk++)
                    umma_i8(taddr, da + ((2 * LBO * kk) >> 4), db + ((2 * LBO * kk) >> 4), idesc, kk > 0);
                commit2(bar);
            }
            __syncwarp();
        }
        if (PREFETCH && tile + gridDim.x < tiles) load_row(tile + gridDim.x);
        mbar_wait(bar, parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool by_bulk = bulk_ok != 0 && tile < full_out;            // a whole tile inside the vector
        const size_t o0 = ((size_t)tile * CTA + tid) * (size_t)k;
        uint32_t d[8], dn[8];
        tmem_ld8(my_taddr, d);
        for (int e = 0; e < k; e++) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (e + 1 < k) tmem_ld8(my_taddr + 8 * (e + 1), dn);      // the next secret's limbs, in flight under this compose
            const uint64_t r = compose2<S::W5>(d, two16);
            if (by_bulk) my_out[e] = r;
            else if (o0 + e < dimension) out[o0 + e] = (int64_t)r;      // batched.rs:94 truncate
#pragma unroll
            for (int i = 0; i < 8; i++) d[i] = dn[i];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();          // TMEM and the rows are free for the next tile; the staged secrets are complete
        if (by_bulk && tid == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(out + (size_t)tile * CTA * k), "r"(out_base), "r"((uint32_t)(CTA * k * 8)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // shared memory outlives the last store
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(tmem_cols) : "memory");
}

template <int CH, bool PREFETCH>
cudaError_t launch(const LaunchCtx &lc, int k, int m, const int64_t *shares, size_t ld, size_t nbatches, size_t dimension,
                   const uint8_t *d_b_image, int64_t *out) {
    typedef RevealShape<CH> S;
    auto kern = reveal_tc_kernel<CH, PREFETCH>;
    const int tmem_cols = tmem_cols_for(k);
    const uint32_t b_bytes = (uint32_t)(n_mma_for(k) / 8) * S::SBO_B;
    // never more CTAs on an SM than can hold their TMEM columns (tc_common.cuh): this kernel needs few registers
    // and little shared memory, so without the floor 12 CTAs land on an SM that has columns for 512 / TMEM_COLS
    const size_t smem = smem_capping_residency(((S::A_BYTES + 127) & ~127u) + ((b_bytes + 127) & ~127u) + (size_t)CTA * k * 8,
                                               512 / tmem_cols);
    const int bulk_ok = reinterpret_cast<uintptr_t>(out) % 16 == 0;      // bulk stores need 16-byte aligned destinations
    static KernelSetup setup;                              // one per instantiation (this function is a template)
    int regs = 0;
    size_t static_smem = 0;
    const cudaError_t se = setup_kernel(setup, kern, 100 * 1024, &regs, &static_smem);    // any shape's request fits
    if (se != cudaSuccess) return se;
    const int per_sm = resident_ctas(regs, CTA, smem, static_smem, tmem_cols);
    const size_t tiles = (nbatches + CTA - 1) / CTA;
    if (tiles >> 32) return cudaErrorInvalidValue;
    const size_t grid = std::min<size_t>(tiles, (size_t)lc.sm_count * per_sm);
    kern<<<(unsigned)grid, CTA, smem, lc.stream>>>(shares, ld, nbatches, dimension, k, m & 1, (uint32_t)tmem_cols, b_bytes,
                                                   idesc_u8(n_mma_for(k)), reinterpret_cast<const uint4 *>(d_b_image), out,
                                                   65536u, bulk_ok);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

// ---- paired tiles: thread r owns the adjacent batches 2r and 2r + 1 of a 256-batch tile (tile E = the even batches, tile O
// the odd ones, as in packed_tc2.cuh).  A clerk's two shares of the pair are one 16-byte load, the 2k secrets of the pair are
// contiguous in the output, and barriers, waits and the bulk store are paid once per 256 batches instead of once per 128.
template <int CH, bool PREFETCH>
__global__ void __launch_bounds__(CTA, PREFETCH || CH > 4 ? 4 : 8)
reveal_tc2_kernel(const int64_t *__restrict__ shares, size_t ld, size_t nbatches, size_t dimension, int k, int m_odd,
                  uint32_t acc_cols, uint32_t b_bytes, uint32_t idesc, const uint4 *__restrict__ b_image,
                  int64_t *__restrict__ out, uint32_t two16, int bulk_ok, int vec_ok) {
    typedef RevealShape<CH> S;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sA = smem;                                                  // tile E, then tile O
    uint8_t *sB = smem + 2 * ((S::A_BYTES + 127) & ~127u);
    uint64_t *sOut = reinterpret_cast<uint64_t *>(sB + ((b_bytes + 127) & ~127u));     // the tile's 256 k secrets
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    constexpr uint32_t A_TILE = (S::A_BYTES + 127) & ~127u;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base)), "r"(2 * acc_cols) : "memory");
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
    const uint32_t ta = taddr + ((uint32_t)(warp * 32) << 16), tb = ta + acc_cols;
    const uint32_t bar = smem_u32(&mbar), a_base = smem_u32(sA), b_base = smem_u32(sB), out_base = smem_u32(sOut);
    uint8_t *my_row = sA + (tid >> 3) * S::SBO_A + (tid & 7) * 16;
    uint64_t *my_out = sOut + (size_t)(2 * tid) * k;
    uint32_t parity = 0;

    const uint32_t tiles = (uint32_t)((nbatches + 2 * CTA - 1) / (2 * CTA));
    const uint32_t full_in = (uint32_t)(nbatches / (2 * CTA));
    const uint32_t full_out = (uint32_t)min((unsigned long long)(dimension / ((size_t)2 * CTA * k)), 0xffffffffull);
    // this thread's pair of a tile: share i of batch 2r in (x, y), of batch 2r + 1 in (z, w)
    uint4 v[2 * CH];
    auto load_pair = [&](uint32_t tile) {
        const size_t b0 = (size_t)tile * (2 * CTA) + 2 * tid;
        const int64_t *src = shares + b0;
        const bool both = tile < full_in || b0 + 1 < nbatches, first = both || b0 < nbatches;
#pragma unroll
        for (int i = 0; i < 2 * CH; i++) {
            const bool present = i < 2 * CH - 1 || !m_odd;
            uint4 x = make_uint4(0, 0, 0, 0);
            if (present && both && vec_ok) {
                x = __ldg(reinterpret_cast<const uint4 *>(src));
            } else if (present && first) {
                const uint2 lo = __ldg(reinterpret_cast<const uint2 *>(src));
                x.x = lo.x; x.y = lo.y;
                if (both) {
                    const uint2 hi = __ldg(reinterpret_cast<const uint2 *>(src + 1));
                    x.z = hi.x; x.w = hi.y;
                }
            }
            v[i] = x;
            src += ld;
        }
    };
    if (PREFETCH && blockIdx.x < tiles) load_pair(blockIdx.x);
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        if (!PREFETCH) load_pair(tile);
        uint32_t sign = 0;
#pragma unroll
        for (int i = 0; i < 2 * CH; i++) sign |= v[i].y | v[i].w;
        if ((int32_t)sign < 0) {                       // negative representatives are rare: one test per pair
#pragma unroll
            for (int i = 0; i < 2 * CH; i++) {
                if ((int32_t)v[i].y < 0) unpack(canon_negative((int64_t)pack(v[i].x, v[i].y)), v[i].x, v[i].y);
                if ((int32_t)v[i].w < 0) unpack(canon_negative((int64_t)pack(v[i].z, v[i].w)), v[i].z, v[i].w);
            }
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
            *reinterpret_cast<uint4 *>(my_row + c * LBO) = make_uint4(v[2 * c].x, v[2 * c].y, v[2 * c + 1].x, v[2 * c + 1].y);
            *reinterpret_cast<uint4 *>(my_row + A_TILE + c * LBO) = make_uint4(v[2 * c].z, v[2 * c].w, v[2 * c + 1].z, v[2 * c + 1].w);
        }
        // the previous tile's bulk store has read its staging buffer before anyone passes the barrier below
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint64_t db = umma_desc(b_base, S::SBO_B);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint64_t da = umma_desc(a_base + h * A_TILE, S::SBO_A);
#pragma unroll
                    for (int kk = 0; kk < S::NK; kk++)
                        umma_i8(taddr + h * acc_cols, da + ((2 * LBO * kk) >> 4), db + ((2 * LBO * kk) >> 4), idesc, kk > 0);
                }
                commit2(bar);
            }
            __syncwarp();
        }
        if (PREFETCH && tile + gridDim.x < tiles) load_pair(tile + gridDim.x);
        mbar_wait(bar, parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool by_bulk = bulk_ok != 0 && tile < full_out;            // a whole tile inside the vector
        const size_t o0 = ((size_t)tile * (2 * CTA) + 2 * tid) * (size_t)k;    // first secret of batch 2r; batch 2r + 1 follows
        uint32_t de[8], dd[8], ne[8], nd[8];
        tmem_ld8(ta, de);
        tmem_ld8(tb, dd);
        for (int e = 0; e < k; e++) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (e + 1 < k) {                                   // the next secret's limbs, in flight under these composes
                tmem_ld8(ta + 8 * (e + 1), ne);
                tmem_ld8(tb + 8 * (e + 1), nd);
            }
            const uint64_t re = compose2<S::W5>(de, two16), ro = compose2<S::W5>(dd, two16);
            if (by_bulk) {
                my_out[e] = re;
                my_out[k + e] = ro;
            } else {
                if (o0 + e < dimension) out[o0 + e] = (int64_t)re;              // batched.rs:94 truncate
                if (o0 + k + e < dimension) out[o0 + k + e] = (int64_t)ro;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                de[i] = ne[i];
                dd[i] = nd[i];
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();          // TMEM and the rows are free for the next tile; the staged secrets are complete
        if (by_bulk && tid == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(out + (size_t)tile * (2 * CTA) * k), "r"(out_base), "r"((uint32_t)(2 * CTA * k * 8)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // shared memory outlives the last store
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(2 * acc_cols) : "memory");
}

template <int CH, bool PREFETCH>
cudaError_t launch2t(const LaunchCtx &lc, int k, int m, const int64_t *shares, size_t ld, size_t nbatches, size_t dimension,
                     const uint8_t *d_b_image, int64_t *out) {
    typedef RevealShape<CH> S;
    auto kern = reveal_tc2_kernel<CH, PREFETCH>;
    const int acc_cols = tmem_cols_for(k);
    const uint32_t b_bytes = (uint32_t)(n_mma_for(k) / 8) * S::SBO_B;
    const size_t smem = smem_capping_residency(2 * ((S::A_BYTES + 127) & ~127u) + ((b_bytes + 127) & ~127u) + (size_t)2 * CTA * k * 8,
                                               512 / (2 * acc_cols));
    const int bulk_ok = reinterpret_cast<uintptr_t>(out) % 16 == 0;
    const int vec_ok = reinterpret_cast<uintptr_t>(shares) % 16 == 0 && ld % 2 == 0;     // 16-byte loads of a pair's shares
    static KernelSetup setup;
    int regs = 0;
    size_t static_smem = 0;
    const cudaError_t se = setup_kernel(setup, kern, 120 * 1024, &regs, &static_smem);    // any shape's request fits
    if (se != cudaSuccess) return se;
    const int per_sm = resident_ctas(regs, CTA, smem, static_smem, 2 * acc_cols);
    const size_t tiles = (nbatches + 2 * CTA - 1) / (2 * CTA);
    if (tiles >> 32) return cudaErrorInvalidValue;
    const size_t grid = std::min<size_t>(tiles, (size_t)lc.sm_count * per_sm);
    kern<<<(unsigned)grid, CTA, smem, lc.stream>>>(shares, ld, nbatches, dimension, k, m & 1, (uint32_t)acc_cols, b_bytes,
                                                   idesc_u8(n_mma_for(k)), reinterpret_cast<const uint4 *>(d_b_image), out,
                                                   65536u, bulk_ok, vec_ok);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

template <int CH>
cudaError_t launch_ch(const LaunchCtx &lc, int k, int m, const int64_t *shares, size_t ld, size_t nbatches, size_t dimension,
                      const uint8_t *d_b_image, int64_t *out) {
    static const char *which = getenv("SDA_B200_REVEAL_KERNEL");     // "single": the one-batch-per-thread kernel (side-by-side runs)
    if (!(which && which[0] == 's') && tmem_cols_for(k) <= 128)
        return tmem_cols_for(k) >= 64 ? launch2t<CH, true>(lc, k, m, shares, ld, nbatches, dimension, d_b_image, out)
                                      : launch2t<CH, false>(lc, k, m, shares, ld, nbatches, dimension, d_b_image, out);
    return tmem_cols_for(k) >= 64 ? launch<CH, true>(lc, k, m, shares, ld, nbatches, dimension, d_b_image, out)
                                  : launch<CH, false>(lc, k, m, shares, ld, nbatches, dimension, d_b_image, out);
}

}  // namespace

bool reveal_tc_supported(int k, int m) { return k >= 1 && k <= 16 && m >= 1 && m <= 16; }

static uint32_t sbo_b_for(int m) { return 2u * (uint32_t)(((m + 1) / 2 + 1) / 2) * 128u; }

size_t reveal_tc_image_bytes(int k, int m) { return reveal_tc_supported(k, m) ? (size_t)(n_mma_for(k) / 8) * sbo_b_for(m) : 0; }

// B[(e,l)][(s,c)] = limb l of (R[e][s] 2^{8c} mod p) under the limb plan of ceil(m / 2) chunks, in the shared-memory
// operand layout
void reveal_tc_build_image(int k, int m, const Matrix &R, uint8_t *img) {
    typedef unsigned __int128 u128;
    const uint32_t sbo_b = sbo_b_for(m);
    const LimbPlan lp = limb_plan(2 * ((m + 1) / 2));
    memset(img, 0, reveal_tc_image_bytes(k, m));
    for (int e = 0; e < k; e++)
        for (int s = 0; s < m; s++)
            for (int byte = 0; byte < 8; byte++) {
                const uint64_t cst = (uint64_t)((u128)R.e[e * m + s] * ((((u128)1) << (8 * byte)) % P61) % P61);
                for (int l = 0; l < 8; l++) {
                    const int n = e * 8 + l;
                    img[(n / 8) * sbo_b + (s / 2) * LBO + (n % 8) * 16 + (s % 2) * 8 + byte] =
                        (uint8_t)((cst >> lp.pos[l]) & ((1u << lp.w[l]) - 1u));
                }
            }
}

cudaError_t launch_reveal_tc(const LaunchCtx &lc, int k, int m, const int64_t *shares, size_t ld, size_t dimension,
                             const uint8_t *d_b_image, int64_t *secrets_out) {
    if (dimension == 0) return cudaSuccess;
    if (!reveal_tc_supported(k, m)) return cudaErrorInvalidValue;
    const size_t nbatches = (dimension + k - 1) / k;
    switch ((m + 1) / 2) {
#define X(CH) case CH: return launch_ch<CH>(lc, k, m, shares, ld, nbatches, dimension, d_b_image, secrets_out);
    X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#undef X
    }
    return cudaErrorInvalidValue;
}

}  // namespace sda
