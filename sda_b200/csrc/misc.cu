// misc.cu -- recipient-side packed-Shamir reconstruction, the exact gen_range stream
// (stream compaction), and the synthetic-input generator.
//
//   reconstruct   client/src/crypto/sharing/batched.rs:68-97 + packed_shamir.rs:73-77 -> tss 0.2
//                 `reconstruct` (Newton interpolation through (1,0) and the present clerks'
//                 points, evaluated at the secret points).  That map is linear and depends only
//                 on the index set, so the host builds R (k x m') once and the kernel does
//                 secrets_b = R . shares_b for every batch b.
//   draw_exact    rand 0.3 `Range::ind_sample` loop: the e-th sample is the
//                 e-th ACCEPTED u64 of the stream -> count / scan / scatter over the keystream.
#include <algorithm>
#include <cmath>

#include "kernels.h"
#include "vecio.cuh"

namespace sda {

namespace {

constexpr int CTA = 128;

// ---- reveal ---------------------------------------------------------------------------------
template <bool M61>
__global__ void __launch_bounds__(CTA)
packed_reconstruct_kernel(const int64_t *__restrict__ shares, size_t ld, size_t nbatches, size_t dimension, int k,
                          int m, Matrix R, int64_t *__restrict__ out, FieldParams f, int lazy) {
    __shared__ uint64_t r_s[MAX_K * MAX_N];
    for (int i = threadIdx.x; i < k * m; i += CTA) r_s[i] = R.e[i];
    __syncthreads();
    const size_t b = (size_t)blockIdx.x * CTA + threadIdx.x;
    if (b >= nbatches) return;
    uint64_t y[MAX_N];
#pragma unroll 4
    for (int s = 0; s < m; s++) y[s] = canon<M61>(f, __ldg(shares + (size_t)s * ld + b));   // batched.rs:83-85
    for (int e = 0; e < k; e++) {
        const size_t o = b * (size_t)k + e;
        if (o >= dimension) break;                                                         // batched.rs:94
        uint64_t lo = 0, hi = 0;
        int cnt = 0;
        for (int s = 0; s < m; s++) {
            const uint64_t a = r_s[e * m + s];
            const uint64_t pl = a * y[s], ph = __umul64hi(a, y[s]);
            lo += pl;
            hi += ph + (lo < pl);
            if (++cnt == lazy) {
                lo = M61 ? reduce128_m61(hi, lo) : reduce128_generic(f, hi, lo);
                hi = 0;
                cnt = 0;
            }
        }
        out[o] = (int64_t)(M61 ? reduce128_m61(hi, lo) : reduce128_generic(f, hi, lo));
    }
}

// ---- fixed-point codec of real-valued vectors (not in the reference) ---------
// encode: q = rint(x 2^frac_bits) (ties to even, exact in double) -> residue in [0, m)
template <bool M61>
__global__ void __launch_bounds__(CTA)
fixed_encode_kernel(const float *__restrict__ x, size_t n, double scale, int64_t *__restrict__ out, FieldParams f) {
    size_t i = (size_t)blockIdx.x * CTA + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * CTA;
    for (; i < n; i += stride) out[i] = (int64_t)canon<M61>(f, __double2ll_rn((double)__ldg(x + i) * scale));
}
// decode: centred lift of the residue, / 2^frac_bits / divisor in double, rounded to float
template <bool M61>
__global__ void __launch_bounds__(CTA)
fixed_decode_kernel(const int64_t *__restrict__ in, size_t n, double scale, double divisor, float *__restrict__ out,
                    FieldParams f) {
    size_t i = (size_t)blockIdx.x * CTA + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * CTA;
    for (; i < n; i += stride) {
        const uint64_t r = canon<M61>(f, __ldg(in + i));
        const int64_t c = r > f.m / 2 ? (int64_t)r - (int64_t)f.m : (int64_t)r;
        out[i] = (float)__ddiv_rn(__ddiv_rn((double)c, scale), divisor);
    }
}

// ---- synthetic inputs -----------------------------------------------------------------------
__global__ void __launch_bounds__(CTA)
synth_fill_kernel(ChaChaKey key, uint64_t start, size_t count, int64_t *__restrict__ out, FieldParams f, bool m61) {
    const uint64_t blk0 = start / 8;
    const size_t nblk = (size_t)((start + count + 7) / 8 - blk0);
    size_t u = (size_t)blockIdx.x * CTA + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * CTA;
    for (; u < nblk; u += stride) {
        uint64_t v[8];
        chacha_draws8<20>(key, blk0 + u, v);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint64_t e = (blk0 + u) * 8 + i;
            if (e >= start && e < start + count)
                out[e - start] = (int64_t)(m61 ? reduce64_m61(v[i]) : reduce64_generic(f, v[i]));
        }
    }
}

// ---- exact gen_range stream -------------------------------------------------------------------
constexpr int XCTA = 256;               // threads per CTA, 8 stream positions each
constexpr size_t XCHUNK = XCTA * 8;     // stream positions per CTA

template <int ROUNDS>
__device__ __forceinline__ unsigned accept_mask(const ChaChaKey &key, const DrawParams &dr, size_t u, size_t window,
                                                uint64_t (&v)[8]) {
    chacha_draws8<ROUNDS>(key, u, v);
    unsigned mask = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (u * 8 + i < window && v[i] < dr.zone) mask |= 1u << i;
    return mask;
}

template <int ROUNDS>
__global__ void __launch_bounds__(XCTA)
draw_count_kernel(ChaChaKey key, DrawParams dr, size_t window, uint64_t *__restrict__ counts) {
    __shared__ unsigned warp_sum[XCTA / 32];
    const size_t u = (size_t)blockIdx.x * XCTA + threadIdx.x;
    uint64_t v[8];
    unsigned c = __popc(accept_mask<ROUNDS>(key, dr, u, window, v));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < XCTA / 32; w++) t += warp_sum[w];
        counts[blockIdx.x] = t;
    }
}

// single-CTA exclusive scan of counts[0..n) in place; total -> counts[n]
__global__ void __launch_bounds__(1024) scan_kernel(uint64_t *counts, size_t n, size_t need, unsigned *status) {
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t carry_s, total_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (size_t base = 0; base < n; base += 1024) {
        const size_t i = base + threadIdx.x;
        const uint64_t x = i < n ? counts[i] : 0;
        uint64_t incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint64_t t = wsum[lane];
            uint64_t ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t y = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += y;
            }
            wsum[lane] = ti - t;
            if (lane == 31) total_s = ti;
        }
        __syncthreads();
        if (i < n) counts[i] = carry_s + wsum[warp] + incl - x;
        __syncthreads();
        if (threadIdx.x == 0) carry_s += total_s;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        counts[n] = carry_s;
        if (carry_s < need) *status = 1u;
    }
}

template <int ROUNDS>
__global__ void __launch_bounds__(XCTA)
draw_write_kernel(ChaChaKey key, DrawParams dr, size_t window, const uint64_t *__restrict__ offsets, size_t count,
                  uint64_t *__restrict__ out) {
    __shared__ unsigned warp_off[XCTA / 32];
    const size_t u = (size_t)blockIdx.x * XCTA + threadIdx.x;
    uint64_t v[8];
    const unsigned mask = accept_mask<ROUNDS>(key, dr, u, window, v);
    const unsigned c = __popc(mask);
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += y;
    }
    if ((threadIdx.x & 31) == 31) warp_off[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned wbase = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) wbase += warp_off[w];
    size_t pos = offsets[blockIdx.x] + wbase + (incl - c);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (mask & (1u << i)) {
            if (pos < count) {
                bool r;
                out[pos] = dr.kind == DRAW_M61          ? draw_reduce<DRAW_M61>(dr, v[i], r)
                           : dr.kind == DRAW_M61_MINUS1 ? draw_reduce<DRAW_M61_MINUS1>(dr, v[i], r)
                                                        : draw_reduce<DRAW_GENERIC>(dr, v[i], r);
            }
            pos++;
        }
    }
}

int lazy_terms(uint64_t m) {
    unsigned __int128 cap = ((unsigned __int128)m << 64) - m;
    unsigned __int128 sq = (unsigned __int128)(m - 1) * (m - 1);
    if (sq == 0) return 1 << 20;
    unsigned __int128 q = cap / sq;
    if (q > (1u << 20)) q = 1u << 20;
    return q ? (int)q : 1;
}

}  // namespace

cudaError_t launch_packed_reconstruct(const LaunchCtx &lc, const FieldParams &f, int k, int m, const Matrix &R,
                                      const int64_t *shares, size_t ld, size_t dimension, int64_t *secrets_out) {
    if (dimension == 0) return cudaSuccess;
    const size_t nb = (dimension + k - 1) / k;
    const unsigned grid = (unsigned)((nb + CTA - 1) / CTA);
    if (f.kind == FIELD_MERSENNE61)
        packed_reconstruct_kernel<true><<<grid, CTA, 0, lc.stream>>>(shares, ld, nb, dimension, k, m, R, secrets_out,
                                                                     f, 8);
    else
        packed_reconstruct_kernel<false><<<grid, CTA, 0, lc.stream>>>(shares, ld, nb, dimension, k, m, R, secrets_out,
                                                                      f, lazy_terms(f.m));
    ++*lc.nlaunch;
    return cudaGetLastError();
}

cudaError_t launch_fixed_encode(const LaunchCtx &lc, const FieldParams &f, int frac_bits, const float *x, size_t n,
                                int64_t *out) {
    if (n == 0) return cudaSuccess;
    size_t ctas = std::min<size_t>((n + CTA - 1) / CTA, (size_t)lc.sm_count * 32);
    const double scale = ldexp(1.0, frac_bits);
    if (f.kind == FIELD_MERSENNE61) fixed_encode_kernel<true><<<(unsigned)ctas, CTA, 0, lc.stream>>>(x, n, scale, out, f);
    else fixed_encode_kernel<false><<<(unsigned)ctas, CTA, 0, lc.stream>>>(x, n, scale, out, f);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

cudaError_t launch_fixed_decode(const LaunchCtx &lc, const FieldParams &f, int frac_bits, uint64_t divisor,
                                const int64_t *in, size_t n, float *out) {
    if (n == 0) return cudaSuccess;
    size_t ctas = std::min<size_t>((n + CTA - 1) / CTA, (size_t)lc.sm_count * 32);
    const double scale = ldexp(1.0, frac_bits);
    if (f.kind == FIELD_MERSENNE61)
        fixed_decode_kernel<true><<<(unsigned)ctas, CTA, 0, lc.stream>>>(in, n, scale, (double)divisor, out, f);
    else
        fixed_decode_kernel<false><<<(unsigned)ctas, CTA, 0, lc.stream>>>(in, n, scale, (double)divisor, out, f);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

cudaError_t launch_synth_fill(const LaunchCtx &lc, const FieldParams &f, uint32_t stream_id, uint64_t start,
                              size_t count, int64_t *out) {
    if (count == 0) return cudaSuccess;
    static const char tag[] = "sda-b200-synthetic-v1";
    uint8_t kb[32] = {0};
    for (size_t i = 0; i < sizeof(tag) - 1; i++) kb[i] = (uint8_t)tag[i];
    ChaChaKey key = key_from_seed_bytes(kb);
    key.w[7] = stream_id;
    const size_t nblk = (size_t)((start + count + 7) / 8 - start / 8);
    size_t ctas = (nblk + CTA - 1) / CTA;
    const size_t cap = (size_t)lc.sm_count * 64;
    if (ctas > cap) ctas = cap;
    synth_fill_kernel<<<(unsigned)ctas, CTA, 0, lc.stream>>>(key, start, count, out, f, f.kind == FIELD_MERSENNE61);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

size_t draw_exact_scratch_elems(size_t window) { return (window + XCHUNK - 1) / XCHUNK + 1; }

cudaError_t launch_draw_exact(const LaunchCtx &lc, const DrawParams &dr, int rounds, const ChaChaKey &key,
                              size_t count, size_t window, uint64_t *out, uint64_t *scratch, unsigned *status) {
    if (count == 0) return cudaSuccess;
    const size_t nchunks = (window + XCHUNK - 1) / XCHUNK;
    if (nchunks > 0x7fffffffu) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)nchunks;
    if (rounds == 8) draw_count_kernel<8><<<grid, XCTA, 0, lc.stream>>>(key, dr, window, scratch);
    else if (rounds == 12) draw_count_kernel<12><<<grid, XCTA, 0, lc.stream>>>(key, dr, window, scratch);
    else draw_count_kernel<20><<<grid, XCTA, 0, lc.stream>>>(key, dr, window, scratch);
    ++*lc.nlaunch;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    scan_kernel<<<1, 1024, 0, lc.stream>>>(scratch, nchunks, count, status);
    ++*lc.nlaunch;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (rounds == 8) draw_write_kernel<8><<<grid, XCTA, 0, lc.stream>>>(key, dr, window, scratch, count, out);
    else if (rounds == 12) draw_write_kernel<12><<<grid, XCTA, 0, lc.stream>>>(key, dr, window, scratch, count, out);
    else draw_write_kernel<20><<<grid, XCTA, 0, lc.stream>>>(key, dr, window, scratch, count, out);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

}  // namespace sda
