// tc_common.cuh -- device helpers shared by the tcgen05 kernels (packed_tc.cu, reveal_tc.cu): UMMA shared-memory
// descriptors for the no-swizzle K-major operand layout, the kind::i8 MMA, mbarrier waits, TMEM loads, and
// the composition of eight byte-limb sums into a canonical residue mod 2^61 - 1.
//
// Operand layout (both operands, u8, K-major, SWIZZLE_NONE): a core matrix is 8 rows x 16 bytes stored
// contiguously (128 bytes); LBO = 128 bytes steps to the next 16-byte chunk along K, SBO steps to the next
// 8 rows.  Row r, K byte kb of a tile therefore lives at (r / 8) * SBO + (kb / 16) * 128 + (r % 8) * 16 + kb % 16.
// One MMA consumes 32 bytes of K = two chunks.  tools/tc_probe.cu checks this encoding against a host GEMM.
#pragma once
#include <cstdint>
#include <mutex>

#include "field.cuh"

namespace sda {
namespace tc {

constexpr uint32_t LBO = 128;
constexpr uint32_t LOW29 = 0x1fffffffu;

// instruction descriptor: s32 accumulators, u8 x u8, both operands K-major, M = 128, N = n_mma (multiple of 16)
__host__ __device__ constexpr uint32_t idesc_u8(int n_mma) {
    return (2u << 4) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// The block scheduler places CTAs by registers, threads and shared memory only.  A CTA that lands on an SM whose 512
// TMEM columns are taken sits in tcgen05.alloc until a neighbour EXITS -- with persistent CTAs that serialises
// whole shares of the grid and leaves other SMs empty (profiles/r01_k2_variants.md).  Kernels whose residency is
// bounded by TMEM therefore ask for at least this much dynamic shared memory: more than 1 / (max_ctas + 1) of an
// SM's 228 KB, so that max_ctas + 1 CTAs can never be resident whatever the per-CTA overheads are.
inline size_t smem_capping_residency(size_t dyn_smem, int max_ctas) {
    const size_t floor_bytes = (228u * 1024u) / (size_t)(max_ctas + 1) + 1;
    return dyn_smem > floor_bytes ? dyn_smem : floor_bytes;
}

// Function attributes are per device and the host entry points may be called from several threads (one context
// each): raise a kernel's dynamic shared-memory limit once per device, under a lock, and remember what the
// occupancy arithmetic needs.
struct KernelSetup {
    std::mutex mu;
    bool done[64] = {};
    int regs[64] = {};
    size_t static_smem[64] = {};
};
template <class Kern>
inline cudaError_t setup_kernel(KernelSetup &ks, Kern kern, size_t max_dyn_smem, int *regs, size_t *static_smem) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(ks.mu);
    if (!ks.done[dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncAttributes fa;
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, kern);
        if (e != cudaSuccess) return e;
        ks.regs[dev] = fa.numRegs;
        ks.static_smem[dev] = fa.sharedSizeBytes;
        ks.done[dev] = true;
    }
    *regs = ks.regs[dev];
    *static_smem = ks.static_smem[dev];
    return cudaSuccess;
}
// CTAs of `threads` threads an SM can hold, given registers, shared memory and TMEM columns per CTA
inline int resident_ctas(int regs, int threads, size_t dyn_smem, size_t static_smem, int tmem_cols) {
    const int by_regs = 65536 / (((regs + 7) & ~7) * threads);
    const int by_smem = (int)((227u * 1024u) / (dyn_smem + static_smem + 1024));
    const int by_tmem = 512 / tmem_cols;
    const int n = by_regs < by_smem ? by_regs : by_smem;
    return n < by_tmem ? (n > 1 ? n : 1) : by_tmem;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t taddr, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        :: "r"(taddr), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// the retry loop lives inside the asm block: written as a C loop around try_wait, ptxas recomputes the
// barrier's shared-window address (S2R SR_CgaCtaId, MOV, LEA) on every iteration of the spin
// try_wait suspends the thread for up to the time hint (ns) before it reports failure: with a generous hint a waiting warp
// sleeps instead of re-issuing the test (profiles/r02_k2.md: the retries were 1.2 % of the share-gen kernel's instructions)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "SDA_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra SDA_MBAR_DONE;\n\t"
        "bra SDA_MBAR_WAIT;\n"
        "SDA_MBAR_DONE:\n\t}"
        :: "r"(bar), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
// the first 8 n columns of this thread's lane (n shares x 8 limb sums) in as few loads as powers of two allow
template <int N>
__device__ __forceinline__ void tmem_ld_shares(uint32_t taddr, uint32_t (&d)[N][8]) {
    uint32_t *v = &d[0][0];
    int c = 0;
#pragma unroll
    for (; c + 32 <= 8 * N; c += 32) tmem_ld32(taddr + c, v + c);
    if (8 * N - c >= 16) {
        tmem_ld16(taddr + c, v + c);
        c += 16;
    }
    if (8 * N - c >= 8) tmem_ld8(taddr + c, *reinterpret_cast<uint32_t (*)[8]>(v + c));
}
__device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
__device__ __forceinline__ void unpack(uint64_t v, uint32_t &lo, uint32_t &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t mac_u(uint32_t a, uint32_t b, uint64_t c) {     // see packed_m61.cu
    uint64_t d;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %1, %2;\n\tadd.u64 %0, %3, t;\n\t}" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}

// canonical  sum_s d[s] 2^{8s}  mod p  for limb sums d[s] < 2^23
__device__ __forceinline__ uint64_t compose(const uint32_t (&d)[8], uint32_t two16) {
    const uint32_t e0 = d[0] + (d[1] << 8), e1 = d[2] + (d[3] << 8);            // < 2^31.6
    const uint32_t e2 = d[4] + (d[5] << 8), e3 = d[6] + (d[7] << 8);
    const uint64_t L = mac_u(e1, two16, e0), H = mac_u(e3, two16, e2);           // < 2^48: value = L + H 2^32
    uint32_t l_lo, l_hi, h_lo, h_hi;
    unpack(L, l_lo, l_hi);
    unpack(H, h_lo, h_hi);
    // H 2^32 = (h_lo & 2^29-1) 2^32 + (h_lo >> 29) 2^61 + h_hi 2^64 == (..) 2^32 + (h_lo >> 29) + 8 h_hi
    const uint32_t small = (h_lo >> 29) + (h_hi * 8u + 1u);                     // + 1: t == value + 1
    const uint64_t t = pack(l_lo, l_hi + (h_lo & LOW29)) + small;               // in [1, 2^62)
    uint32_t t_lo, t_hi;
    unpack(t, t_lo, t_hi);
    const int64_t qm1 = (int64_t)(int32_t)((t_hi >> 29) - 1u);                  // floor((t - 1) / p) - 1 in {-1, 0}
    uint32_t r_lo, r_hi;
    unpack(t + (uint64_t)qm1, r_lo, r_hi);
    return pack(r_lo, r_hi & LOW29);
}

static __device__ __noinline__ uint64_t canon_negative(int64_t v) {                 // v < 0 -> [0, p)
    const uint64_t a = 0ull - (uint64_t)v;
    uint64_t r = (a & P61) + (a >> 61);
    r = r >= P61 ? r - P61 : r;
    return r ? P61 - r : 0;
}


// ---- the uneven limb plan of the paired-tile share generation and of the reveal kernel --------------------------
// width of limb 5 of the constant operand (see compose2): the widest for which e2 = d4 + 256 d5 stays below 2^29
// when a limb sum has 8 (k + t) terms of at most 255 * (2^w - 1)
constexpr int w5_for(int kt) {
    for (int w5 = 8; w5 >= 5; w5--)
        if ((long long)8 * kt * 255 * (255 + ((1ll << w5) - 1) * 256) < (1ll << 29)) return w5;
    return 0;
}

struct LimbPlan {
    int w[8], pos[8];
};
inline LimbPlan limb_plan(int kt) {
    const int w5 = w5_for(kt);
    LimbPlan lp{{8, 8, 8, 8, 8, w5, 8, 13 - w5}, {0, 8, 16, 24, 32, 40, 40 + w5, 48 + w5}};
    return lp;
}


// canonical  sum_s d[s] 2^{pos[s]}  mod p  for the limb sums of one share (limb plan of k + t: positions
// 0,8,16,24,32,40,40+W5,48+W5).  e0..e3 are 32-bit; X = e0 + e1 2^16 + e2 2^32 < 2^61 + 2^48 is one wide multiply
// whose addend is the register pair (e0, e2); e3 2^{40+W5} == (e3 mod 2^{21-W5}) 2^{40+W5} + (e3 >> (21-W5)) because
// 2^61 == 1; t = value + 1 lies in [1, 2^62) and the last four instructions are packed_tc.cu's.
#ifndef SDA_TC2_XWIDE
#define SDA_TC2_XWIDE 0      // 1: X by IMAD.WIDE with a run-time multiplier (one FMA-pipe instruction) instead of LEA + LEA.HI.X
#endif
#ifndef SDA_TC2_FINAL_X
#define SDA_TC2_FINAL_X 1    // 1: the final 64-bit add takes its high addend from a register (IMAD.X) instead of a sign extension
#endif
template <int W5>
__device__ __forceinline__ uint64_t compose2(const uint32_t (&d)[8], uint32_t two16) {
    const uint32_t e0 = d[0] + (d[1] << 8), e1 = d[2] + (d[3] << 8);
    const uint32_t e2 = d[4] + (d[5] << 8), e3 = d[6] + (d[7] << 8);
    uint64_t x;
#if SDA_TC2_XWIDE
    // two16 == 65536 is a kernel parameter so that ptxas keeps the multiply (it turns a literal into two shifts-and-adds)
    asm("{\n\t.reg .u64 a;\n\tmov.b64 a, {%1, %2};\n\tmad.wide.u32 %0, %3, %4, a;\n\t}" : "=l"(x) : "r"(e0), "r"(e2), "r"(e1), "r"(two16));
#else
    asm("{\n\t.reg .u64 a;\n\tmov.b64 a, {%1, %2};\n\tmad.wide.u32 %0, %3, 65536, a;\n\t}" : "=l"(x) : "r"(e0), "r"(e2), "r"(e1));
#endif
    constexpr uint32_t SH3 = 8 + W5;                                  // position of e3 inside the high word
    const uint32_t m3 = (e3 << SH3) & (LOW29 & ~((1u << SH3) - 1u));
    const uint32_t s3 = (e3 >> (21 - W5)) + 1u;                       // + 1: t == value + 1
    const uint64_t t = x + pack(s3, m3);
    uint32_t t_lo, t_hi;
    unpack(t, t_lo, t_hi);
    const uint32_t qm1 = (t_hi >> 29) - 1u;                           // floor((t - 1) / p) - 1 in {-1, 0} as two's complement
    uint32_t r_lo, r_hi;
#if SDA_TC2_FINAL_X
    unpack(t + pack(qm1, qm1), r_lo, r_hi);                           // the 64-bit value -1 or 0: both words equal qm1
#else
    unpack(t + (uint64_t)(int64_t)(int32_t)qm1, r_lo, r_hi);
#endif
    return pack(r_lo, r_hi & LOW29);
}


__device__ __forceinline__ void commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}


}  // namespace tc
}  // namespace sda
