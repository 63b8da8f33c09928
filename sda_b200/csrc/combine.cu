// combine.cu -- K3: column-wise modular sums over the participant axis.
//
// Replaces ShareCombiner::combine (client/src/crypto/sharing/combiner.rs:15-29), the additive
// SecretReconstructor (additive.rs:55-73) and the Full MaskCombiner (masking/full.rs:37-52):
//     out[i] = fold over rows p of (out[i] + row_p[i]) % m
// The reference reduces after every add (one idiv per element); here each thread carries a
// signed 128-bit sum per column and reduces once, so the loop is 5 integer ops per 8 bytes
// and the kernel is bound by HBM: 8 bytes read per input share element.
//
// Layout: rows[P][ld] participant-major as the clerk receives them (server snapshot
// transpose, server/src/snapshot.rs:11-27).  A thread owns 4 adjacent columns (one 256-bit
// load per row), a CTA 1024 adjacent columns (8 KB contiguous per row), grid.y splits the
// participant axis into slices so that the grid is many waves deep; slice partials are
// canonical and are summed by a second, tiny launch of the same kernel.
#include "kernels.h"

namespace sda {

namespace {

constexpr int CTA = 256;

// streaming loads: read once, keep out of L1, first in line for L2 eviction
template <int VEC>
struct Pack;
template <>
struct Pack<1> {
    int64_t v[1];
    __device__ __forceinline__ void load(const int64_t *p) {
        asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(v[0]) : "l"(p));
    }
    __device__ __forceinline__ void store(int64_t *p) const { p[0] = v[0]; }
};
template <>
struct Pack<2> {
    int64_t v[2];
    __device__ __forceinline__ void load(const int64_t *p) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(v[0]), "=l"(v[1]) : "l"(p));
    }
    __device__ __forceinline__ void store(int64_t *p) const {
        asm volatile("st.global.v2.s64 [%0], {%1,%2};" ::"l"(p), "l"(v[0]), "l"(v[1]) : "memory");
    }
};
template <>
struct Pack<4> {   // sm_100 256-bit global access
    int64_t v[4];
    __device__ __forceinline__ void load(const int64_t *p) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3])
                     : "l"(p));
    }
    __device__ __forceinline__ void store(int64_t *p) const {
        asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v[0]), "l"(v[1]), "l"(v[2]), "l"(v[3])
                     : "memory");
    }
};

// bias for `rows` summed rows: m * ceil(rows * 2^63 / m), as (hi, lo)
struct Bias {
    uint64_t hi, lo;
};

template <bool M61, int VEC>
__global__ void __launch_bounds__(CTA)
combine_kernel(const int64_t *__restrict__ rows, size_t ld, size_t P, size_t L, size_t rows_per_slice,
               const int64_t *acc_in, int64_t *out, size_t out_ld, FieldParams f, Bias bias) {
    constexpr int UNROLL = VEC == 4 ? 4 : 8;
    const size_t p0 = (size_t)blockIdx.y * rows_per_slice;
    const size_t p1 = min(P, p0 + rows_per_slice);
    int64_t *o = out + (size_t)blockIdx.y * out_ld;
    const size_t c = ((size_t)blockIdx.x * CTA + threadIdx.x) * VEC;
    if (c >= L) return;   // L % VEC == 0 on the vector paths
    Acc128 a[VEC];
#pragma unroll
    for (int j = 0; j < VEC; j++) a[j].init();
    const int64_t *col = rows + c;
    size_t p = p0;
    for (; p + UNROLL <= p1; p += UNROLL) {
        Pack<VEC> v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) v[u].load(col + (p + u) * ld);
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int j = 0; j < VEC; j++) a[j].add(v[u].v[j]);
    }
    for (; p < p1; p++) {
        Pack<VEC> v;
        v.load(col + p * ld);
#pragma unroll
        for (int j = 0; j < VEC; j++) a[j].add(v.v[j]);
    }
    if (acc_in != nullptr && blockIdx.y == 0) {
#pragma unroll
        for (int j = 0; j < VEC; j++) a[j].add(acc_in[c + j]);
    }
    Pack<VEC> r;
#pragma unroll
    for (int j = 0; j < VEC; j++) r.v[j] = (int64_t)reduce_acc<M61>(f, a[j], bias.hi, bias.lo);
    r.store(o + c);
}

template <bool M61, bool UNSIGNED>
__global__ void __launch_bounds__(CTA)
mod_reduce_kernel(const int64_t *__restrict__ in, size_t n, int64_t *__restrict__ out, FieldParams f) {
    size_t i = (size_t)blockIdx.x * CTA + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * CTA;
    for (; i < n; i += stride)
        out[i] = (int64_t)(UNSIGNED ? reduce64<M61>(f, (uint64_t)in[i]) : canon<M61>(f, in[i]));
}

template <bool M61>
__global__ void __launch_bounds__(CTA)
submod_kernel(const int64_t *__restrict__ a, const int64_t *__restrict__ b, size_t n, int64_t *__restrict__ out,
              FieldParams f) {
    size_t i = (size_t)blockIdx.x * CTA + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * CTA;
    for (; i < n; i += stride) out[i] = (int64_t)submod(canon<M61>(f, a[i]), canon<M61>(f, b[i]), f.m);
}

Bias make_bias(uint64_t m, size_t rows) {
    // smallest multiple of m that is >= rows * 2^63
    unsigned __int128 need = (unsigned __int128)rows << 63;
    unsigned __int128 q = (need + m - 1) / m;
    unsigned __int128 b = q * m;
    return Bias{(uint64_t)(b >> 64), (uint64_t)b};
}

// how the participant axis is cut: enough CTAs for ~16 waves, at least 32 rows per slice
size_t pick_slices(int sm_count, size_t P, size_t col_ctas) {
    const size_t want = (size_t)sm_count * 8 * 16;
    if (col_ctas >= want || P < 64) return 1;
    size_t s = (want + col_ctas - 1) / col_ctas;
    size_t max_s = P / 32;
    if (s > max_s) s = max_s;
    if (s > 65535) s = 65535;
    return s ? s : 1;
}

int pick_vec(const void *p, size_t ld, size_t L) {
    const uintptr_t a = (uintptr_t)p;
    if (a % 32 == 0 && ld % 4 == 0 && L % 4 == 0) return 4;
    if (a % 16 == 0 && ld % 2 == 0 && L % 2 == 0) return 2;
    return 1;
}

template <bool M61>
cudaError_t combine_pass(const LaunchCtx &lc, const FieldParams &f, const int64_t *rows, size_t ld, size_t P,
                         size_t L, size_t slices, const int64_t *acc_in, int64_t *out, size_t out_ld) {
    const size_t rps = (P + slices - 1) / slices;
    const Bias bias = make_bias(f.m, rps + 1);
    int vec = pick_vec(rows, ld, L);
    const int vo = pick_vec(out, out_ld, L);
    if (vo < vec) vec = vo;
    const size_t cols_per_cta = (size_t)CTA * vec;
    dim3 grid((unsigned)((L + cols_per_cta - 1) / cols_per_cta), (unsigned)slices);
    if (vec == 4)
        combine_kernel<M61, 4><<<grid, CTA, 0, lc.stream>>>(rows, ld, P, L, rps, acc_in, out, out_ld, f, bias);
    else if (vec == 2)
        combine_kernel<M61, 2><<<grid, CTA, 0, lc.stream>>>(rows, ld, P, L, rps, acc_in, out, out_ld, f, bias);
    else
        combine_kernel<M61, 1><<<grid, CTA, 0, lc.stream>>>(rows, ld, P, L, rps, acc_in, out, out_ld, f, bias);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

}  // namespace

size_t combine_scratch_elems(int sm_count, size_t P, size_t L) {
    const size_t col_ctas = (L + 4 * CTA - 1) / (4 * CTA);
    const size_t s = pick_slices(sm_count, P, col_ctas ? col_ctas : 1);
    const size_t Lp = (L + 3) & ~(size_t)3;
    return s > 1 ? s * Lp : 0;
}

cudaError_t launch_combine(const LaunchCtx &lc, const FieldParams &f, const int64_t *rows, size_t ld, size_t P,
                           size_t L, const int64_t *acc_in, int64_t *out, int64_t *scratch, size_t scratch_elems) {
    if (L == 0) return cudaSuccess;
    const bool m61 = f.kind == FIELD_MERSENNE61;
    const size_t col_ctas = (L + 4 * CTA - 1) / (4 * CTA);
    size_t slices = pick_slices(lc.sm_count, P, col_ctas);
    const size_t Lp = (L + 3) & ~(size_t)3;
    if (slices > 1 && (scratch == nullptr || scratch_elems < slices * Lp)) slices = 1;
    cudaError_t e;
    if (slices == 1) {
        return m61 ? combine_pass<true>(lc, f, rows, ld, P, L, 1, acc_in, out, Lp)
                   : combine_pass<false>(lc, f, rows, ld, P, L, 1, acc_in, out, Lp);
    }
    e = m61 ? combine_pass<true>(lc, f, rows, ld, P, L, slices, nullptr, scratch, Lp)
            : combine_pass<false>(lc, f, rows, ld, P, L, slices, nullptr, scratch, Lp);
    if (e != cudaSuccess) return e;
    return m61 ? combine_pass<true>(lc, f, scratch, Lp, slices, L, 1, acc_in, out, Lp)
               : combine_pass<false>(lc, f, scratch, Lp, slices, L, 1, acc_in, out, Lp);
}

cudaError_t launch_mod_reduce(const LaunchCtx &lc, const FieldParams &f, const int64_t *in, size_t n, int64_t *out,
                              bool input_unsigned) {
    if (n == 0) return cudaSuccess;
    size_t ctas = (n + CTA - 1) / CTA;
    const size_t cap = (size_t)lc.sm_count * 32;
    if (ctas > cap) ctas = cap;
    const bool m61 = f.kind == FIELD_MERSENNE61;
    if (m61 && input_unsigned) mod_reduce_kernel<true, true><<<(unsigned)ctas, CTA, 0, lc.stream>>>(in, n, out, f);
    else if (m61) mod_reduce_kernel<true, false><<<(unsigned)ctas, CTA, 0, lc.stream>>>(in, n, out, f);
    else if (input_unsigned) mod_reduce_kernel<false, true><<<(unsigned)ctas, CTA, 0, lc.stream>>>(in, n, out, f);
    else mod_reduce_kernel<false, false><<<(unsigned)ctas, CTA, 0, lc.stream>>>(in, n, out, f);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

cudaError_t launch_submod(const LaunchCtx &lc, const FieldParams &f, const int64_t *a, const int64_t *b, size_t n,
                          int64_t *out) {
    if (n == 0) return cudaSuccess;
    size_t ctas = (n + CTA - 1) / CTA;
    const size_t cap = (size_t)lc.sm_count * 32;
    if (ctas > cap) ctas = cap;
    if (f.kind == FIELD_MERSENNE61)
        submod_kernel<true><<<(unsigned)ctas, CTA, 0, lc.stream>>>(a, b, n, out, f);
    else
        submod_kernel<false><<<(unsigned)ctas, CTA, 0, lc.stream>>>(a, b, n, out, f);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

}  // namespace sda
