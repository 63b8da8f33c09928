// packed_tc2.cuh -- K2 for p = 2^61 - 1, second generation of the byte-limb GEMM of packed_tc.cu
// (client/src/crypto/sharing/packed_shamir.rs:40-43 -> tss 0.2 `share`, + batched.rs:18-53).
//
// The arithmetic is the one packed_tc.cu explains: shares_b = M . [secrets_b ; draws_b] is linear in the BYTES of its
// inputs, so D = A . B^T with A = the raw bytes of the rows and B = the limbs of (M[j][i] 2^{8c} mod p) is a dense
// u8 GEMM on tcgen05.mma.kind::i8, and a thread that owns a batch only folds eight limb sums per share.
// profiles/r01_k2_tc.md showed that kernel bound by instruction issue (0.69 IPC with every pipe below its own
// limit), so this one is the same dataflow with fewer instructions per batch:
//
//   * tiles come in PAIRS: MMA tile E holds the even batches of 256 consecutive ones, tile O the odd batches,
//     so thread r owns the adjacent batches 2r and 2r + 1 -- its two results of a share row are ONE 16-byte
//     store, its 2K secrets are K aligned 16-byte words of the raw vector which go into the operand tiles as they
//     are (for odd K the middle word serves both rows; the constant operand has a zero where the neighbour's
//     secret sits, hence two operand images, E and O);
//   * the limbs of the constant operand are not 8 x 8 bits: widths (8,8,8,8,8,W5,8,13-W5) with W5 chosen per
//     k + t so that  e0 + e1 2^16 + e2 2^32  stays below 2^61: a shift-and-add pair on the register PAIR (e0, e2), and
//     the top term folds with one mask and one shift -- 15 instructions per share and no wide multiply instead of
//     16 and two (tests/test_k2_model.py restates it with Python integers and checks every bound);
//   * store addresses are a running 64-bit pointer per thread (one add per share row);
//   * pass bookkeeping ((participant, pass) of the next unit, "does this pass lie inside the vector") is 32-bit and
//     warp-uniform, and the MMAs are issued from warp-uniform code (elect.sync) with uniform descriptors;
//   * the keystream block is fully unrolled and starts from per-participant first-round constants (chacha_pre.cuh).
//
// The same kernel serves, by template flags: RTN -- the share count at run time for every (k, t) with k + t <= 16
// (packed_tc2n.cu); MASKED -- the participant's mask drawn and added while the secrets are staged (packed_tc2m.cu).  The
// share-gen -> clerk-sum kernel of packed_tc2f.cu is built from the same pieces.
//
// Shared memory per CTA (k=3, t=2, n=5): draws 2 x 8 KB, secrets 16 KB, two operand images 6 KB, raw secrets of
// the coming pass 12 KB = 51 KB, four CTAs per SM (TMEM: two 64-column accumulators each).
#pragma once
#include <algorithm>
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"
#include "chacha_pre.cuh"

namespace sda {

namespace {

using namespace tc;

constexpr int CTA2 = 128;            // threads = rows of one MMA tile = TMEM lanes
constexpr int SDA_TC2_MAX_GROUPS = 4;  // run-time share count: up to 4 groups of 8 shares

constexpr int gcd_k(int a, int b) { return b == 0 ? a : gcd_k(b, a % b); }

// FORCE_PAIRS: 0 = the choice below; the fused share-gen -> clerk-sum kernel forces one pair per pass (packed_tc2f.cu)
template <int K, int T, int N, int FORCE_PAIRS = 0>
struct Shape2 {
    static_assert(T >= 1, "at least one draw per batch");
    static constexpr int TT = T;
    static constexpr int KT = K + T;
    static constexpr int W5 = w5_for(KT);
    static_assert(W5 >= 5, "k + t too large for the limb plan");
    static constexpr int DC = (T + 1) / 2;                   // 16-byte chunks of a row holding draws (odd t: half of the last unused)
    static constexpr int SC = (K + 1) / 2;                   // ... holding secrets
    static constexpr int NKD = (DC + 1) / 2, NKS = (SC + 1) / 2;
    static constexpr int NK = NKD + NKS;                     // MMAs per tile (32 bytes of K each)
    static constexpr int NMMA = (8 * N + 15) / 16 * 16;
    // shared memory with `pairs` tile pairs (256 batches each) per pass: draws twice, secrets, two images, raw secrets
    static constexpr uint32_t smem_for(int pairs) {
        return 2 * (pairs * 2 * 16 * DC * 128 + 128) + (pairs * 2 * 16 * SC * 128 + 128) + 2 * (NMMA / 8 * 2 * NK * 128) +
               pairs * 256 * K * 8;
    }
    // two pairs per pass give every thread whole keystream blocks when t is not a multiple of 4 -- where four CTAs
    // still fit an SM with them (56960 bytes each after the per-CTA reserve); otherwise one pair, and the warps of a CTA
    // share the pass's 32 t blocks unevenly
    static constexpr int PAIRS = FORCE_PAIRS ? FORCE_PAIRS : (T % 4 != 0 && smem_for(2) <= 56960 ? 2 : 1);
    static constexpr int PASS = PAIRS * 256;                 // batches per pass
    static constexpr int NBLK = PASS * T / 8;                // keystream blocks per pass (8 draws each), a multiple of 32
    static constexpr int NB = (NBLK + CTA2 - 1) / CTA2;      // ... per thread; odd t leaves the last round to warps 0 and 1
    // an odd chunk count lets the second chunk of the last K step alias the next 8-row group: B is 0 there
    static constexpr uint32_t SBO_D = DC * 128, SBO_S = SC * 128;
    static constexpr uint32_t D_TILE = 16 * SBO_D, S_TILE = 16 * SBO_S;
    static constexpr uint32_t D_BYTES = PAIRS * 2 * D_TILE + 128, S_BYTES = PAIRS * 2 * S_TILE + 128;
    static constexpr uint32_t SBO_B = 2 * NK * 128;
    static constexpr uint32_t B_IMG = NMMA / 8 * SBO_B;      // one operand image; the kernel holds two (E, O)
    static constexpr uint32_t IN_BYTES = PASS * K * 8;       // the raw secrets of one pass
    static constexpr uint32_t SMEM = 2 * D_BYTES + S_BYTES + 2 * B_IMG + IN_BYTES;
    static constexpr int ACC_COLS = NMMA <= 32 ? 32 : NMMA <= 64 ? 64 : 128;
    static constexpr int ACC_BUFS = 2 * ACC_COLS <= 128 ? 2 : 1;   // E and O side by side when four CTAs still fit
    static constexpr int TMEM_COLS = ACC_BUFS * ACC_COLS;
    static constexpr uint32_t IDESC = idesc_u8(NMMA);
    static_assert(SMEM == smem_for(PAIRS), "smem_for restates SMEM");
    // CTAs per SM that shared memory and TMEM allow: what the register allocation of the RTN instantiations is held to
    static constexpr int RESIDENT = SMEM + 1152 > 227 * 1024 / 2 ? 1 : SMEM + 1152 > 227 * 1024 / 3 ? 2 : SMEM + 1152 > 227 * 1024 / 4 ? 3 : 4;
    static_assert(D_BYTES % 128 == 0 && S_BYTES % 128 == 0 && B_IMG % 16 == 0 && IN_BYTES % 16 == 0, "alignment");
    // the four BASELINE shapes fit four CTAs per SM (SMEM <= 56 KB); larger k of the run-time-share-count grid fit fewer
};

// draw v = (hi word w0, lo word w1) -> a u64 congruent to v mod (p - 1) modulo p:  v mod (p - 1) = (v & p) + 2 (v >> 61)
// unless that reaches p - 1, and v itself == (v & p) + (v >> 61) (mod p), so X = v + (v >> 61) serves as the operand row:
// any u64 bit pattern is a valid row (the GEMM is linear in its bytes).  gen_range differs from this only when
// v mod 2^61 >= 2^61 - 32 (a rejected word, a wrap-around of the reduction, or the overflow of this very sum).  The
// necessary condition "bits 32..60 all ones" (2^-29 per draw) is accumulated as the running maximum of the masked
// high words (one three-input VIMNMX per two draws), and the caller settles `suspect >= 2^29 - 1` exactly.
__device__ __forceinline__ uint64_t reduce_draw2(uint32_t w0, uint32_t w1) {
    uint32_t h;
    asm("shr.u32 %0, %1, 29;" : "=r"(h) : "r"(w0));
    return pack(w1, w0) + (uint64_t)h;
}
// the running maximum over a block's 8 draws.  SDA_TC2_SUSPECT_PAIRS: two draws share one masked word -- "the OR of
// their high words has bits 0..28 all ones" is still necessary for either of them, costs one LOP3 per pair instead of
// one per draw, and is true by chance for 2^-12 of the pairs (the exact test that follows then finds nothing).
#ifndef SDA_TC2_SUSPECT_PAIRS
#define SDA_TC2_SUSPECT_PAIRS 1
#endif
__device__ __forceinline__ uint32_t suspect_of_block(const uint32_t (&w)[16]) {
    uint32_t suspect = 0;
#if SDA_TC2_SUSPECT_PAIRS
#pragma unroll
    for (int d = 0; d < 8; d += 2) suspect = max(suspect, (w[2 * d] | w[2 * d + 2]) & LOW29);
#else
#pragma unroll
    for (int d = 0; d < 8; d++) suspect = max(suspect, w[2 * d] & LOW29);
#endif
    return suspect;
}

// one elected lane of a converged warp: the NK MMAs of a 128-row tile into the accumulator at `taddr`
template <class S>
__device__ __forceinline__ void issue_tile2(uint32_t taddr, uint32_t d_tile, uint32_t s_tile, uint32_t b_img) {
    const uint64_t dd = umma_desc(d_tile, S::SBO_D), ds = umma_desc(s_tile, S::SBO_S), db = umma_desc(b_img, S::SBO_B);
#pragma unroll
    for (int kk = 0; kk < S::NKD; kk++)
        umma_i8(taddr, dd + ((2 * LBO * kk) >> 4), db + ((2 * LBO * kk) >> 4), S::IDESC, kk > 0);
#pragma unroll
    for (int kk = 0; kk < S::NKS; kk++)
        umma_i8(taddr, ds + ((2 * LBO * kk) >> 4), db + ((2 * LBO * (S::NKD + kk)) >> 4), S::IDESC, 1);
}
// the keystream of one pass: NB blocks per thread, every draw reduced and scattered into the row it belongs to.
// Chunk gc of the pass (16 bytes = two draws, stream order) belongs to batch gc / DC of the pass; batch beta sits in
// tile pair beta / 256, tile E or O by its parity, row (beta % 256) / 2.
// a participant's key and first-round constants, loaded well ahead of the keystream that needs them (the loads are
// the only global-memory latency on a pass's critical path)
struct KeyRegs {
    uint4 ka, kb, pa, pb, pc;
};
__device__ __forceinline__ KeyRegs load_keys2(const ChaChaKey *__restrict__ keys, const ChaChaPre *__restrict__ pres, uint32_t p) {
    const uint4 *src = reinterpret_cast<const uint4 *>(keys + p);
    const uint4 *ps = reinterpret_cast<const uint4 *>(pres + p);
    KeyRegs r;
    r.ka = __ldg(src);
    r.kb = __ldg(src + 1);
    r.pa = __ldg(ps);
    r.pb = __ldg(ps + 1);
    r.pc = __ldg(ps + 2);
    return r;
}

template <class S, int ROUNDS>
__device__ __forceinline__ void stage_draws2(const KeyRegs &kr, uint32_t u, int tid, uint8_t *sD, unsigned *flag) {
    uint32_t k[8], pre[12];
    k[0] = kr.ka.x; k[1] = kr.ka.y; k[2] = kr.ka.z; k[3] = kr.ka.w;
    k[4] = kr.kb.x; k[5] = kr.kb.y; k[6] = kr.kb.z; k[7] = kr.kb.w;
    pre[0] = kr.pa.x; pre[1] = kr.pa.y; pre[2] = kr.pa.z; pre[3] = kr.pa.w;
    pre[4] = kr.pb.x; pre[5] = kr.pb.y; pre[6] = kr.pb.z; pre[7] = kr.pb.w;
    pre[8] = kr.pc.x; pre[9] = kr.pc.y; pre[10] = kr.pc.z; pre[11] = kr.pc.w;
    const uint32_t blk0 = u * (uint32_t)S::NBLK;                  // the launcher keeps a participant below 2^32 blocks
    // where 16-byte chunk c of batch beta (of the pass) lies
    auto chunk_at = [&](uint32_t beta, uint32_t c) {
        const uint32_t tile = (beta >> 8) * 2 + (beta & 1), row = (beta & 255) >> 1;
        return sD + tile * S::D_TILE + (row >> 3) * S::SBO_D + c * LBO + (row & 7) * 16;
    };
#pragma unroll 1
    for (int nb = 0; nb < S::NB; nb++) {
        const uint32_t slot = nb * CTA2 + tid;                    // block of this pass, in stream order
        if constexpr (S::NBLK % CTA2 != 0) {
            if (slot >= (uint32_t)S::NBLK) break;                 // warp-uniform: NBLK is a multiple of 32
        }
        uint32_t w[16];
        chacha_block2<ROUNDS>(k, pre, blk0 + slot, w);
        const uint32_t suspect = suspect_of_block(w);
        if constexpr (S::TT % 2 != 0) {
            // odd t: a block's 8 draws cross batch boundaries at odd positions -- one 8-byte store per draw
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint64_t x = reduce_draw2(w[2 * i], w[2 * i + 1]);
                const uint32_t g = slot * 8 + i, beta = g / S::TT, c = g % S::TT;
                uint32_t xl, xh;
                unpack(x, xl, xh);
                *reinterpret_cast<uint2 *>(chunk_at(beta, c >> 1) + (c & 1) * 8) = make_uint2(xl, xh);
            }
        } else if constexpr (!(S::DC == 1 || S::DC == 2 || S::DC % 4 == 0)) {
            // even t whose chunks per batch do not divide a block's four: every chunk placed on its own
#pragma unroll
            for (int cb = 0; cb < 4; cb++) {
                const uint64_t xa = reduce_draw2(w[4 * cb], w[4 * cb + 1]);
                const uint64_t xb = reduce_draw2(w[4 * cb + 2], w[4 * cb + 3]);
                const uint32_t gc = slot * 4 + cb;
                uint32_t xal, xah, xbl, xbh;
                unpack(xa, xal, xah);
                unpack(xb, xbl, xbh);
                *reinterpret_cast<uint4 *>(chunk_at(gc / S::DC, gc % S::DC)) = make_uint4(xal, xah, xbl, xbh);
            }
        } else {
        // the block's four chunks: chunk 0 goes to `dst`, the others to compile-time offsets from it
        const uint32_t gc0 = slot * 4;
        const uint32_t beta0 = gc0 / S::DC, c0 = gc0 % S::DC;
        const uint32_t tile0 = (beta0 >> 8) * 2 + (beta0 & 1), row0 = (beta0 & 255) >> 1;
        uint8_t *dst = sD + tile0 * S::D_TILE + (row0 >> 3) * S::SBO_D + c0 * LBO + (row0 & 7) * 16;
#pragma unroll
        for (int cb = 0; cb < 4; cb++) {                          // 4 chunks of 2 draws
            const uint64_t xa = reduce_draw2(w[4 * cb], w[4 * cb + 1]);
            const uint64_t xb = reduce_draw2(w[4 * cb + 2], w[4 * cb + 3]);
            // DC = 1: batches beta0 .. beta0 + 3 (beta0 a multiple of 4): tiles E, O, E, O, rows row0, row0, row0 + 1, row0 + 1
            // DC = 2: batches beta0, beta0 + 1 (beta0 even): tiles E, E, O, O, chunks 0, 1, 0, 1 of row row0
            // DC % 4 == 0: one batch, chunks c0 .. c0 + 3
            const uint32_t delta = S::DC == 1 ? (cb & 1) * S::D_TILE + (cb >> 1) * 16
                                 : S::DC == 2 ? (cb >> 1) * S::D_TILE + (cb & 1) * LBO
                                              : cb * LBO;
            uint32_t xal, xah, xbl, xbh;
            unpack(xa, xal, xah);
            unpack(xb, xbl, xbh);
            *reinterpret_cast<uint4 *>(dst + delta) = make_uint4(xal, xah, xbl, xbh);
        }
        }
        if (suspect >= LOW29) {
            bool bad = false;
#pragma unroll
            for (int d = 0; d < 8; d++) bad |= (w[2 * d] & LOW29) == LOW29 && w[2 * d + 1] >= 0xffffffe0u;
            if (bad) atomicOr(flag, 1u);
        }
    }
}

// The raw secrets of pass (p, u) into the staging buffer by the threads themselves, zero beyond the vector
// (batched.rs:38-43): the path of a pass that is not wholly inside the vector or whose source is not 16-byte aligned.
// Every thread writes, and later reads, only its own 2K words per pair.
template <class S, int K>
__device__ __forceinline__ void fill_secrets2(const int64_t *__restrict__ secrets, size_t ld, size_t dim, uint32_t p, uint32_t u,
                                              int tid, int64_t *sIn) {
    const int64_t *sec = secrets + (size_t)p * ld;
    const size_t e_first = ((size_t)u * S::PASS + 2 * tid) * K;
#pragma unroll
    for (int q = 0; q < S::PAIRS; q++) {
        const size_t e0 = e_first + (size_t)q * (256 * K);
#pragma unroll
        for (int i = 0; i < 2 * K; i++) sIn[(q * 256 + 2 * tid) * K + i] = e0 + i < dim ? __ldg(sec + e0 + i) : 0;
    }
}

__device__ __forceinline__ void canon_pair2(uint32_t &lo, uint32_t &hi);

// MASKED kernels (participate.rs:53-54 then :75-76 in one pass): the mask of every secret of pass (p, u) is drawn here and
// added to the raw secrets where they lie in the staging buffer, so the masked secrets exist only as operand rows.
// Element e of a participant takes draw e of its mask stream (full.rs:24-27, chacha.rs:38-41: one gen_range(0, m) per
// secret): keystream block e / 8, a pass's PASS K / 8 blocks spread over the threads.  Over 2^61 - 1 a draw is
// sd = (v & p) + (v >> 61) -- gen_range's value unless the low 61 bits of v are within 8 of 2^61 (a rejected word or a
// wrap of the sum), which raises `flag` for the caller to redo the call on the exact path -- and a secret enters as any
// u64 congruent to it: a negative one is canonicalised first, x + sd cannot overflow, and no reduction is needed (the
// GEMM is linear in the bytes of its rows).  Elements at or beyond `dim` are the zero padding of the last batch
// (batched.rs:38-43) and stay unmasked.  mask_row != nullptr: the Full scheme's mask vector of this participant.
template <class S, int K, int ROUNDS>
__device__ __forceinline__ void add_masks2(const KeyRegs &kr, uint32_t u, int tid, int64_t *sIn, size_t dim,
                                           int64_t *__restrict__ mask_row, unsigned *flag) {
    constexpr int NMB = S::PASS * K / 8;                          // mask blocks of a pass
    uint32_t k[8], pre[12];
    k[0] = kr.ka.x; k[1] = kr.ka.y; k[2] = kr.ka.z; k[3] = kr.ka.w;
    k[4] = kr.kb.x; k[5] = kr.kb.y; k[6] = kr.kb.z; k[7] = kr.kb.w;
    pre[0] = kr.pa.x; pre[1] = kr.pa.y; pre[2] = kr.pa.z; pre[3] = kr.pa.w;
    pre[4] = kr.pb.x; pre[5] = kr.pb.y; pre[6] = kr.pb.z; pre[7] = kr.pb.w;
    pre[8] = kr.pc.x; pre[9] = kr.pc.y; pre[10] = kr.pc.z; pre[11] = kr.pc.w;
    const size_t e_pass = (size_t)u * (S::PASS * K);              // first element of the pass
#pragma unroll 1
    for (int nb = 0; nb < (NMB + CTA2 - 1) / CTA2; nb++) {
        const uint32_t slot = nb * CTA2 + tid;
        if constexpr (NMB % CTA2 != 0) {
            if (slot >= (uint32_t)NMB) break;
        }
        uint32_t w[16];
        chacha_block2<ROUNDS>(k, pre, u * (uint32_t)NMB + slot, w);
        if (suspect_of_block(w) >= LOW29) {
            bool bad = false;
#pragma unroll
            for (int d = 0; d < 8; d++) bad |= (w[2 * d] & LOW29) == LOW29 && w[2 * d + 1] >= 0xfffffff8u;
            if (bad) atomicOr(flag, 1u);
        }
        int64_t *row = sIn + slot * 8;
        const size_t e0 = e_pass + (size_t)slot * 8;
#pragma unroll
        for (int h = 0; h < 4; h++) {                             // two elements per 16-byte word
            uint4 x = *reinterpret_cast<const uint4 *>(row + 2 * h);
            uint64_t sd[2];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const uint32_t w0 = w[4 * h + 2 * i], w1 = w[4 * h + 2 * i + 1];
                sd[i] = pack(w1, w0 & LOW29) + (uint64_t)(w0 >> 29);
                if (e0 + 2 * h + i >= dim) sd[i] = 0;             // padding of the last batch: no mask, no mask output
            }
            if ((int32_t)(x.y | x.w) < 0) {
                canon_pair2(x.x, x.y);
                canon_pair2(x.z, x.w);
            }
            uint32_t al, ah, bl, bh;
            unpack(pack(x.x, x.y) + sd[0], al, ah);
            unpack(pack(x.z, x.w) + sd[1], bl, bh);
            *reinterpret_cast<uint4 *>(row + 2 * h) = make_uint4(al, ah, bl, bh);
            if (mask_row != nullptr) {
                if (e0 + 2 * h < dim) mask_row[e0 + 2 * h] = (int64_t)sd[0];
                if (e0 + 2 * h + 1 < dim) mask_row[e0 + 2 * h + 1] = (int64_t)sd[1];
            }
        }
    }
}

// one thread: the whole pass -- PASS batches x K secrets, contiguous in the participant's vector -- into the staging
// buffer with one bulk copy; completion (by byte count) on `bar`
template <class S, int K>
__device__ __forceinline__ void bulk_load_secrets2(const int64_t *__restrict__ secrets, size_t ld, uint32_t p, uint32_t u,
                                                   uint32_t sin_addr, uint32_t bar) {
    const int64_t *src = secrets + (size_t)p * ld + (size_t)u * (S::PASS * K);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(S::IN_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(sin_addr), "l"(src), "r"(S::IN_BYTES), "r"(bar) : "memory");
}

__device__ __forceinline__ void canon_pair2(uint32_t &lo, uint32_t &hi) {
    if ((int32_t)hi < 0) unpack(canon_negative((int64_t)pack(lo, hi)), lo, hi);
}

__device__ __forceinline__ void st_global_v2(char *ptr, uint64_t a, uint64_t b) {
    asm volatile("st.global.v2.u64 [%0], {%1, %2};" :: "l"(ptr), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void st_global_1(char *ptr, uint64_t a) {
    asm volatile("st.global.u64 [%0], %1;" :: "l"(ptr), "l"(a) : "memory");
}

// the odd batch's shares composed and both batches' shares stored: share row j of the pair at `off + j row_bytes`
template <class S, int N>
__device__ __forceinline__ void store_pair(const uint32_t (&d)[N][8], const uint64_t (&re)[N], char *ptr, size_t row_bytes,
                                           bool fast_store, uint32_t live, uint32_t two16) {
    if (fast_store) {
#pragma unroll
        for (int j = 0; j < N; j++) {
            st_global_v2(ptr, re[j], compose2<S::W5>(d[j], two16));
            ptr += row_bytes;
        }
    } else {
#pragma unroll
        for (int j = 0; j < N; j++) {
            const uint64_t ro = compose2<S::W5>(d[j], two16);
            if (live >= 1) st_global_1(ptr, re[j]);
            if (live == 2) st_global_1(ptr + 8, ro);
            ptr += row_bytes;
        }
    }
}

// RTN: the share count is a run-time value; the operand images and accumulators are sized for groups of N shares, and a
// group is folded and stored four shares per 32-column tcgen05.ld, then singly, instead of through N-wide unrolled register
// arrays.  packed_tc2n.cu instantiates it.  MASKED: see add_masks2.
template <int K, int T, int N, int ROUNDS, bool RTN = false, bool MASKED = false>
#ifndef SDA_TC2_RTN_CHUNK
#define SDA_TC2_RTN_CHUNK 1          // the run-time share count folds four shares per TMEM load and wait (0: one at a time, pipelined: 1-10 % slower)
#endif
#ifndef SDA_TC2_TEMPLATED_MINB
#define SDA_TC2_TEMPLATED_MINB 1
#endif
__global__ void __launch_bounds__(CTA2, RTN || MASKED ? Shape2<K, T, N>::RESIDENT : SDA_TC2_TEMPLATED_MINB)
packed_share_tc2_kernel(const int64_t *__restrict__ secrets, size_t ld, size_t dim, size_t B, uint32_t unit_begin,
                        uint32_t units_per_p, uint32_t units_total, uint32_t full_in_units, uint32_t full_out_units,
                        const ChaChaKey *__restrict__ keys, const ChaChaPre *__restrict__ pres, const uint4 *__restrict__ b_image,
                        int64_t *__restrict__ out, unsigned *flag, int bulk_ok, int vec_ok, uint32_t two16, uint32_t n_rt,
                        const ChaChaKey *__restrict__ mkeys = nullptr, const ChaChaPre *__restrict__ mpres = nullptr,
                        int64_t *__restrict__ mask_out = nullptr) {
    typedef Shape2<K, T, N> S;
    static_assert(!RTN || S::ACC_BUFS == 2, "the run-time share count needs both accumulators of a pair at once");
    const uint32_t nsh = RTN ? n_rt : (uint32_t)N;         // shares per batch
    // RTN: the shares go through the accumulators in groups of N (one pair of operand images per group)
    const uint32_t ngroups = RTN ? (n_rt + N - 1) / N : 1u;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sD = smem;                                    // 2 x (PAIRS x {E, O} tiles x 128 rows x draws)
    uint8_t *sS = smem + 2 * S::D_BYTES;                   // PAIRS x {E, O} tiles x 128 rows x secrets
    uint8_t *sB = sS + S::S_BYTES;                         // the constant operand, E image then O image
    int64_t *sIn = reinterpret_cast<int64_t *>(sB + ngroups * (2 * S::B_IMG));   // the coming pass's raw secrets
    __shared__ __align__(8) uint64_t mbar[3];              // [0] full (MMAs done), [1] drained (TMEM read out), [2] secrets landed
    __shared__ uint32_t tmem_base;

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction, and known to be

    // ---- one-time setup: TMEM, barriers, constant operand ------------------------------------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base)), "n"(S::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar[1])), "n"(CTA2) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar[2])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < ngroups * (2 * S::B_IMG / 16); i += CTA2) reinterpret_cast<uint4 *>(sB)[i] = __ldg(b_image + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    const uint32_t my_taddr = taddr + ((uint32_t)(warp * 32) << 16);
    const uint32_t full_bar = smem_u32(&mbar[0]), drained_bar = smem_u32(&mbar[1]), landed_bar = smem_u32(&mbar[2]);
    const uint32_t d_base = smem_u32(sD), s_base = smem_u32(sS), b_base = smem_u32(sB), sin_addr = smem_u32(sIn);
    uint32_t parity = 0, buf = 0, landed_parity = 0, dparity = 0;

    // (participant, pass) of this CTA's units: unit x is pass unit_begin + x % units_per_p of participant x / units_per_p.
    // All of it is 32-bit and the same in every thread.
    const uint32_t step_p = gridDim.x / units_per_p, step_u = gridDim.x % units_per_p;
    const uint32_t unit_end = unit_begin + units_per_p;
    uint32_t p = blockIdx.x / units_per_p, u = unit_begin + blockIdx.x % units_per_p;
    // a pass that lies wholly inside its vector arrives by bulk copy when the source is 16-byte aligned (bulk_ok)
    auto by_bulk = [&](uint32_t uu) { return bulk_ok != 0 && uu < full_in_units; };

    // output address of (p, u): share row 0, this thread's even batch of the pass's first pair.  A step adds step_p
    // participants and step_u passes; when the pass index wraps, one participant more and units_per_p passes fewer.
    char *opass = reinterpret_cast<char *>(out + (size_t)p * nsh * B + (size_t)u * S::PASS + 2 * tid);
    const int64_t step_bytes = 8 * ((int64_t)step_p * nsh * (int64_t)B + (int64_t)step_u * S::PASS);
    const int64_t wrap_bytes = 8 * ((int64_t)nsh * (int64_t)B - (int64_t)units_per_p * S::PASS);
    if (blockIdx.x < units_total) {
        if (by_bulk(u)) {
            if (tid == 0) bulk_load_secrets2<S, K>(secrets, ld, p, u, sin_addr, landed_bar);
        } else {
            fill_secrets2<S, K>(secrets, ld, dim, p, u, tid, sIn);
        }
        stage_draws2<S, ROUNDS>(load_keys2(keys, pres, p), u, tid, sD, flag);
        if constexpr (MASKED) {
            // the first pass's secrets are in place (landed, or written by their threads): mask them there
            if (by_bulk(u)) {
                mbar_wait(landed_bar, landed_parity);
                landed_parity ^= 1;
            } else {
                __syncthreads();
            }
            add_masks2<S, K, ROUNDS>(load_keys2(mkeys, mpres, p), u, tid, sIn, dim,
                                     mask_out != nullptr ? mask_out + (size_t)p * dim : nullptr, flag);
        }
    }

    for (uint32_t unit = blockIdx.x; unit < units_total; unit += gridDim.x) {
        // this CTA's next unit; its key is requested now and used after the staging barrier
        uint32_t pn = p + step_p, un = u + step_u;
        if (un >= unit_end) {
            un -= units_per_p;
            pn++;
        }
        const bool more = unit + gridDim.x < units_total;
        KeyRegs knext;
        if (more) knext = load_keys2(keys, pres, pn);
        // ---- this thread's 2K secrets of every pair (batches 2 tid, 2 tid + 1 of the pair): K aligned 16-byte words
        //      of the raw vector, which are the operand chunks as they are -------------------------------------------
        if constexpr (MASKED) {
            __syncthreads();                     // the masked secrets were written by other threads (add_masks2)
        } else if (by_bulk(u)) {
            mbar_wait(landed_bar, landed_parity);
            landed_parity ^= 1;
        }
        uint4 v[S::PAIRS][K];                    // all loads first: their latency overlaps instead of adding up per pair
#pragma unroll
        for (int q = 0; q < S::PAIRS; q++) {
            const uint4 *row = reinterpret_cast<const uint4 *>(sIn + (q * 256 + 2 * tid) * K);
#pragma unroll
            for (int i = 0; i < K; i++) v[q][i] = row[i];
        }
#pragma unroll
        for (int q = 0; q < S::PAIRS; q++) {
            if constexpr (!MASKED) {             // (masked rows are non-negative sums already, up to 64 bits wide)
                uint32_t sign = 0;
#pragma unroll
                for (int i = 0; i < K; i++) sign |= v[q][i].y | v[q][i].w;
                if ((int32_t)sign < 0) {
#pragma unroll
                    for (int i = 0; i < K; i++) {
                        canon_pair2(v[q][i].x, v[q][i].y);
                        canon_pair2(v[q][i].z, v[q][i].w);
                    }
                }
            }
            uint8_t *te = sS + (2 * q) * S::S_TILE + (tid >> 3) * S::SBO_S + (tid & 7) * 16;
#pragma unroll
            for (int c = 0; c < S::SC; c++) {
                *reinterpret_cast<uint4 *>(te + c * LBO) = v[q][c];                          // E: words 0 .. SC-1
                *reinterpret_cast<uint4 *>(te + S::S_TILE + c * LBO) = v[q][K - S::SC + c];  // O: words K-SC .. K-1
            }
        }
        // rows complete: the secrets just written and the draws written during the previous pass
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_cur = d_base + buf * S::D_BYTES;
        if (warp == 0) {
            if (elect_one()) {
                issue_tile2<S>(taddr, d_cur, s_base, b_base);
                if constexpr (S::ACC_BUFS == 2) issue_tile2<S>(taddr + S::ACC_COLS, d_cur + S::D_TILE, s_base + S::S_TILE, b_base + S::B_IMG);
                commit2(full_bar);
                // everyone is past the barrier, i.e. has read this pass's raw secrets: the next pass's may land
                if (more && by_bulk(un)) bulk_load_secrets2<S, K>(secrets, ld, pn, un, sin_addr, landed_bar);
            }
            __syncwarp();
        }

        // ---- the next pass's keystream, under this pass's first MMAs -----------------------------------
        if (more) {
            stage_draws2<S, ROUNDS>(knext, un, tid, sD + (buf ^ 1) * S::D_BYTES, flag);
            if (!by_bulk(un)) fill_secrets2<S, K>(secrets, ld, dim, pn, un, tid, sIn);
            if constexpr (MASKED) {
                if (by_bulk(un)) {
                    mbar_wait(landed_bar, landed_parity);
                    landed_parity ^= 1;
                } else {
                    __syncthreads();
                }
                add_masks2<S, K, ROUNDS>(load_keys2(mkeys, mpres, pn), un, tid, sIn, dim,
                                         mask_out != nullptr ? mask_out + (size_t)pn * dim : nullptr, flag);
            }
        }

        // ---- per pair: D = A . B^T on the tensor core, then compose the shares of batches 2 tid and 2 tid + 1 ------
        // share row 0 of this thread's first pair of the pass
        char *optr = opass;
        const bool all_live = u < full_out_units;                             // every batch of the pass is below B
        const size_t row_bytes = B * 8u;
        const size_t pass_first = (size_t)u * S::PASS;
#pragma unroll
        for (int q = 0; q < S::PAIRS; q++) {
            uint64_t re[N];
            // batches of this thread that exist: 2 (both), 1 (only the even one) or 0
            uint32_t live = 2;
            if (!all_live) {
                const size_t b_even = pass_first + (size_t)(q * 256 + 2 * tid);
                live = b_even + 1 < B ? 2 : (b_even < B ? 1 : 0);
            }
            const bool fast_store = all_live && vec_ok != 0;                 // uniform: one 16-byte store per share row
            if constexpr (RTN) {
                const uint32_t ta = my_taddr, tb = my_taddr + S::ACC_COLS;
                char *ptr = optr;
                auto put = [&](const uint32_t (&de)[8], const uint32_t (&dd)[8]) {
                    const uint64_t ra = compose2<S::W5>(de, two16), rb = compose2<S::W5>(dd, two16);
                    if (fast_store) st_global_v2(ptr, ra, rb);
                    else {
                        if (live >= 1) st_global_1(ptr, ra);
                        if (live == 2) st_global_1(ptr + 8, rb);
                    }
                    ptr += row_bytes;
                };
#pragma unroll 1
                for (uint32_t g = 0; g < ngroups; g++) {
                    const uint32_t nsg = min((uint32_t)N, nsh - g * N);          // shares of this group
                    mbar_wait(full_bar, parity);
                    parity ^= 1;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#if SDA_TC2_RTN_CHUNK
                    // four shares at a time (one 32-column load per accumulator, one wait), then the rest singly
                    uint32_t j4 = 0;
#pragma unroll 1
                    for (; j4 + 4 <= nsg; j4 += 4) {
                        uint32_t de[4][8], dd[4][8];
                        tmem_ld32(ta + 8 * j4, &de[0][0]);
                        tmem_ld32(tb + 8 * j4, &dd[0][0]);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < 4; i++) put(de[i], dd[i]);
                    }
#pragma unroll 1
                    for (; j4 < nsg; j4++) {
                        uint32_t e0[8], o0[8];
                        tmem_ld8(ta + 8 * j4, e0);
                        tmem_ld8(tb + 8 * j4, o0);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        put(e0, o0);
                    }
#else
                    uint32_t e0[8], o0[8], e1[8], o1[8];
                    tmem_ld8(ta, e0);
                    tmem_ld8(tb, o0);
#pragma unroll 1
                    for (uint32_t j = 0; j < nsg; j += 2) {
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (j + 1 < nsg) {
                            tmem_ld8(ta + 8 * (j + 1), e1);
                            tmem_ld8(tb + 8 * (j + 1), o1);
                        }
                        put(e0, o0);
                        if (j + 1 < nsg) {
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                            if (j + 2 < nsg) {
                                tmem_ld8(ta + 8 * (j + 2), e0);
                                tmem_ld8(tb + 8 * (j + 2), o0);
                            }
                            put(e1, o1);
                        }
                    }
#endif
                    // the accumulators are read out: the next group of this pair, or the first group of the next pair
                    const bool next_group = g + 1 < ngroups;
                    if (next_group || q + 1 < S::PAIRS) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(drained_bar) : "memory");
                        if (warp == 0) {
                            mbar_wait(drained_bar, dparity);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            if (elect_one()) {
                                const uint32_t qn = next_group ? 2 * q : 2 * q + 2;
                                const uint32_t bn = b_base + (next_group ? g + 1 : 0u) * (2 * S::B_IMG);
                                issue_tile2<S>(taddr, d_cur + qn * S::D_TILE, s_base + qn * S::S_TILE, bn);
                                issue_tile2<S>(taddr + S::ACC_COLS, d_cur + (qn + 1) * S::D_TILE, s_base + (qn + 1) * S::S_TILE,
                                               bn + S::B_IMG);
                                commit2(full_bar);
                            }
                            __syncwarp();
                        }
                        dparity ^= 1;
                    }
                }
                optr += 256 * 8;
            } else if constexpr (S::ACC_BUFS == 2) {
                mbar_wait(full_bar, parity);
                parity ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                {
                    uint32_t d[N][8];
                    tmem_ld_shares<N>(my_taddr, d);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < N; j++) re[j] = compose2<S::W5>(d[j], two16);
                }
                uint32_t d[N][8];
                tmem_ld_shares<N>(my_taddr + S::ACC_COLS, d);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (q + 1 < S::PAIRS) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(drained_bar) : "memory");
                    if (warp == 0) {
                        mbar_wait(drained_bar, dparity);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (elect_one()) {
                            issue_tile2<S>(taddr, d_cur + (2 * q + 2) * S::D_TILE, s_base + (2 * q + 2) * S::S_TILE, b_base);
                            issue_tile2<S>(taddr + S::ACC_COLS, d_cur + (2 * q + 3) * S::D_TILE, s_base + (2 * q + 3) * S::S_TILE,
                                           b_base + S::B_IMG);
                            commit2(full_bar);
                        }
                        __syncwarp();
                    }
                    dparity ^= 1;
                }
                store_pair<S, N>(d, re, optr, row_bytes, fast_store, live, two16);
                optr += 256 * 8;
            } else {
                // one accumulator: E, then O into the same columns while E is being composed
                mbar_wait(full_bar, parity);
                parity ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                {
                    uint32_t d[N][8];
                    tmem_ld_shares<N>(my_taddr, d);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(drained_bar) : "memory");
                    if (warp == 0) {
                        mbar_wait(drained_bar, dparity);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (elect_one()) {
                            issue_tile2<S>(taddr, d_cur + (2 * q + 1) * S::D_TILE, s_base + (2 * q + 1) * S::S_TILE, b_base + S::B_IMG);
                            commit2(full_bar);
                        }
                        __syncwarp();
                    }
                    dparity ^= 1;
#pragma unroll
                    for (int j = 0; j < N; j++) re[j] = compose2<S::W5>(d[j], two16);
                }
                mbar_wait(full_bar, parity);
                parity ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t d[N][8];
                tmem_ld_shares<N>(my_taddr, d);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (q + 1 < S::PAIRS) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(drained_bar) : "memory");
                    if (warp == 0) {
                        mbar_wait(drained_bar, dparity);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (elect_one()) {
                            issue_tile2<S>(taddr, d_cur + (2 * q + 2) * S::D_TILE, s_base + (2 * q + 2) * S::S_TILE, b_base);
                            commit2(full_bar);
                        }
                        __syncwarp();
                    }
                    dparity ^= 1;
                }
                store_pair<S, N>(d, re, optr, row_bytes, fast_store, live, two16);
                optr += 256 * 8;
            }
        }
        // every thread is past its TMEM loads of the last pair and every MMA of this pass has completed
        // (`full` was waited on), so the next pass may overwrite the secrets and reuse TMEM
        // the same address for the next unit, by the difference (64-bit multiplies stay out of the loop)
        opass += (un < u ? wrap_bytes : 0) + step_bytes;
        p = pn;
        u = un;
        buf ^= 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(S::TMEM_COLS) : "memory");
}

// the constant operand as it lies in shared memory: image E (rows = even batches of a pair) then image O
template <int K, int T, int N>
void build_b_image2(const Matrix &m, uint64_t p, uint8_t *img, int n_rt = N) {
    typedef Shape2<K, T, N> S;
    typedef unsigned __int128 u128;
    const LimbPlan lp = limb_plan(K + T);
    memset(img, 0, 2 * S::B_IMG);
    for (int par = 0; par < 2; par++)
        for (int j = 0; j < n_rt; j++)
            for (int part = 0; part < 2; part++)                // 0: draws (K steps 0..NKD), 1: secrets (the rest)
                for (int c = 0; c < (part ? S::SC : S::DC); c++)
                    for (int v = 0; v < 2; v++) {
                        // draw / secret index within the batch: an odd row of an odd-K pair starts half a chunk late
                        const int idx = 2 * c + v - ((part && par && (K & 1)) ? 1 : 0);
                        if (idx < 0 || idx >= (part ? K : T)) continue;
                        const int xi = part ? idx : K + idx;        // index into x = [secrets ; randomness]
                        const int cg = part ? 2 * S::NKD + c : c;   // chunk along K of the whole row
                        for (int byte = 0; byte < 8; byte++) {
                            const uint64_t cst = (uint64_t)((u128)m.e[j * (K + T) + xi] * ((((u128)1) << (8 * byte)) % p) % p);
                            for (int s = 0; s < 8; s++) {
                                const int n = j * 8 + s;
                                img[par * S::B_IMG + (n / 8) * S::SBO_B + cg * LBO + (n % 8) * 16 + v * 8 + byte] =
                                    (uint8_t)((cst >> lp.pos[s]) & ((1u << lp.w[s]) - 1u));
                            }
                        }
                    }
}

// MASKED: mkeys = the participants' mask keys, d_pre holds 2 P entries (sharing keys, then mask keys), mask_out = the
// Full scheme's mask vectors [P][dim] or nullptr; first_batch must be 0 and n_batches cover the vector (a mask stream is
// consumed from its first draw)
template <int K, int T, int N, int ROUNDS, bool RTN = false, bool MASKED = false>
cudaError_t launch2(const LaunchCtx &lc, const int64_t *secrets, size_t ld, size_t P, size_t dim, size_t first_batch,
                    size_t n_batches, const ChaChaKey *keys, uint32_t *d_pre, const uint8_t *d_b_image, int64_t *out,
                    unsigned *flag, int n_rt = N, const ChaChaKey *mkeys = nullptr, int64_t *mask_out = nullptr) {
    typedef Shape2<K, T, N> S;
    const size_t B = (dim + K - 1) / K;
    if (first_batch % S::PASS != 0 || first_batch > B) return cudaErrorInvalidValue;
    if (n_batches > B - first_batch) n_batches = B - first_batch;
    const size_t unit_begin = first_batch / S::PASS;
    const size_t units_per_p = (n_batches + S::PASS - 1) / S::PASS;
    const size_t units_total = units_per_p * P;
    if (units_total == 0) return cudaSuccess;
    if ((unit_begin + units_per_p) >> 31 || units_total >> 31 || P >> 31) return cudaErrorInvalidValue;
    auto kern = packed_share_tc2_kernel<K, T, N, ROUNDS, RTN, MASKED>;
    if (MASKED && (mkeys == nullptr || first_batch != 0 || ((dim + 7) / 8) >> 32)) return cudaErrorInvalidValue;
    // never more CTAs on an SM than can hold their TMEM columns (tc_common.cuh)
    // RTN: one more pair of operand images per further group of N shares
    const size_t n_groups = RTN ? ((size_t)n_rt + N - 1) / N : 1;
    const size_t smem = smem_capping_residency(S::SMEM + (n_groups - 1) * (2 * S::B_IMG), 512 / S::TMEM_COLS);
    static KernelSetup setup;
    int regs = 0;
    size_t static_smem = 0;
    // the attribute is set once per device: for RTN, to what the largest share count (SDA_TC2_MAX_GROUPS groups) needs
    const size_t smem_limit = RTN ? smem_capping_residency(S::SMEM + (SDA_TC2_MAX_GROUPS - 1) * (2 * S::B_IMG), 512 / S::TMEM_COLS) : smem;
    if (n_groups > SDA_TC2_MAX_GROUPS || smem_limit > 227u * 1024u) return cudaErrorInvalidValue;
    const cudaError_t se = setup_kernel(setup, kern, smem_limit, &regs, &static_smem);
    if (se != cudaSuccess) return se;
    const int per_sm = resident_ctas(regs, CTA2, smem, static_smem, S::TMEM_COLS);
    size_t grid = (size_t)lc.sm_count * per_sm;
    if (grid > units_total) grid = units_total;
    // bulk copies need 16-byte aligned sources: every pass of every participant starts at an even element
    const int bulk_ok = reinterpret_cast<uintptr_t>(secrets) % 16 == 0 && (ld % 2 == 0 || P == 1);
    // 16-byte stores need every share row to start at an even element
    const int vec_ok = reinterpret_cast<uintptr_t>(out) % 16 == 0 && B % 2 == 0;
    const size_t full_in = dim / ((size_t)S::PASS * K), full_out = B / S::PASS;
    // a participant's keystream stays below 2^32 blocks (B T / 8 of them): the counter's high word is 0
    if ((B * (size_t)T + 7) / 8 >> 32) return cudaErrorInvalidValue;
    ChaChaPre *pres = reinterpret_cast<ChaChaPre *>(d_pre);
    chacha_prepare_kernel<<<(unsigned)((P + 127) / 128), 128, 0, lc.stream>>>(keys, P, pres);
    ++*lc.nlaunch;
    if (MASKED) {
        chacha_prepare_kernel<<<(unsigned)((P + 127) / 128), 128, 0, lc.stream>>>(mkeys, P, pres + P);
        ++*lc.nlaunch;
    }
    kern<<<(unsigned)grid, CTA2, smem, lc.stream>>>(secrets, ld, dim, B, (uint32_t)unit_begin, (uint32_t)units_per_p,
                                                    (uint32_t)units_total, (uint32_t)std::min<size_t>(full_in, 0xffffffffu),
                                                    (uint32_t)std::min<size_t>(full_out, 0xffffffffu), keys, pres,
                                                    reinterpret_cast<const uint4 *>(d_b_image), out, flag, bulk_ok, vec_ok, 65536u, (uint32_t)n_rt,
                                                    mkeys, MASKED ? pres + P : nullptr, mask_out);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

}  // namespace

}  // namespace sda
