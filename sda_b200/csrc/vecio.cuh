// vecio.cuh -- contiguous per-thread runs of i64 moved with the widest global access the
// alignment allows: 256-bit (sm_100 LDG/STG.256), 128-bit, or scalar.  `lanes` (4, 2 or 1
// elements) is kernel-uniform and chosen on the host from pointer/stride alignment.
#pragma once
#include <cstdint>

namespace sda {

__device__ __forceinline__ void ld256_stream(const int64_t *p, int64_t &a, int64_t &b, int64_t &c, int64_t &d) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                 : "l"(p));
}
__device__ __forceinline__ void ld128(const int64_t *p, int64_t &a, int64_t &b) {
    asm volatile("ld.global.nc.v2.s64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}
__device__ __forceinline__ void st256(int64_t *p, int64_t a, int64_t b, int64_t c, int64_t d) {
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ void st128(int64_t *p, int64_t a, int64_t b) {
    asm volatile("st.global.v2.s64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}

// v[0..CNT) = p[0..CNT) for i < nvalid, else 0
template <int CNT>
__device__ __forceinline__ void load_run(const int64_t *p, int64_t (&v)[CNT], int nvalid, int lanes) {
    if (nvalid >= CNT) {
        if (CNT % 4 == 0 && lanes == 4) {
#pragma unroll
            for (int i = 0; i < CNT; i += 4) ld256_stream(p + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
            return;
        }
        if (CNT % 2 == 0 && lanes >= 2) {
#pragma unroll
            for (int i = 0; i < CNT; i += 2) ld128(p + i, v[i], v[i + 1]);
            return;
        }
#pragma unroll
        for (int i = 0; i < CNT; i++) v[i] = __ldg(p + i);
        return;
    }
#pragma unroll
    for (int i = 0; i < CNT; i++) v[i] = i < nvalid ? __ldg(p + i) : 0;
}

template <int CNT>
__device__ __forceinline__ void store_run(int64_t *p, const int64_t (&v)[CNT], int nvalid, int lanes) {
    if (nvalid >= CNT) {
        if (CNT % 4 == 0 && lanes == 4) {
#pragma unroll
            for (int i = 0; i < CNT; i += 4) st256(p + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
            return;
        }
        if (CNT % 2 == 0 && lanes >= 2) {
#pragma unroll
            for (int i = 0; i < CNT; i += 2) st128(p + i, v[i], v[i + 1]);
            return;
        }
#pragma unroll
        for (int i = 0; i < CNT; i++) p[i] = v[i];
        return;
    }
#pragma unroll
    for (int i = 0; i < CNT; i++)
        if (i < nvalid) p[i] = v[i];
}

// widest lane count usable for runs of `cnt` elements starting at base + j*stride (any j)
inline int pick_lanes(const void *base, size_t stride_elems, int cnt) {
    const uintptr_t a = (uintptr_t)base;
    if (cnt % 4 == 0 && a % 32 == 0 && stride_elems % 4 == 0) return 4;
    if (cnt % 2 == 0 && a % 16 == 0 && stride_elems % 2 == 0) return 2;
    return 1;
}

}  // namespace sda
