// snapshot.cu -- the server's snapshot transpose (server/src/snapshot.rs:11-27 -> stores.rs:86-101
// `iter_snapshot_clerk_jobs_data`): every participation carries one encrypted share vector per clerk; a snapshot
// regroups them so that clerk c's job holds blob c of every participation, in participation order.  On one box,
// with the participations' blobs concatenated in device memory, that is a segmented copy:
//     in : blob (p, c) = bytes [in_off[p n + c], in_off[p n + c + 1])      participation-major
//     out: blob (p, c) at bytes [out_off[c P + p], out_off[c P + p + 1])    clerk-major
// One CTA moves one blob (grid-stride over the P n blobs) with 16-byte accesses where source and destination share
// their alignment, bytes at the edges.  Pure data movement: 2 bytes of HBM traffic per payload byte.
#include "kernels.h"

namespace sda {

namespace {

constexpr int TCTA = 256;

__global__ void __launch_bounds__(TCTA)
snapshot_transpose_kernel(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, size_t P, size_t n,
                          uint8_t *__restrict__ out, const uint64_t *__restrict__ out_off) {
    const size_t blobs = P * n;
    for (size_t b = blockIdx.x; b < blobs; b += gridDim.x) {
        const size_t p = b / n, c = b % n;
        const uint8_t *src = in + in_off[b];
        const size_t len = (size_t)(in_off[b + 1] - in_off[b]);
        uint8_t *dst = out + out_off[c * P + p];
        if (((reinterpret_cast<uintptr_t>(src) ^ reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
            // same phase: bytes up to the first 16-byte boundary, vectors, bytes after the last one
            size_t head = (16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15;
            if (head > len) head = len;
            for (size_t i = threadIdx.x; i < head; i += TCTA) dst[i] = src[i];
            const size_t nv = (len - head) / 16;
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
            uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
            for (size_t i = threadIdx.x; i < nv; i += TCTA) d4[i] = __ldg(s4 + i);
            for (size_t i = head + nv * 16 + threadIdx.x; i < len; i += TCTA) dst[i] = src[i];
        } else {
            for (size_t i = threadIdx.x; i < len; i += TCTA) dst[i] = src[i];
        }
    }
}

}  // namespace

cudaError_t launch_snapshot_transpose(const LaunchCtx &lc, const uint8_t *in, const uint64_t *d_in_off, size_t P, size_t n,
                                      uint8_t *out, const uint64_t *d_out_off) {
    const size_t blobs = P * n;
    if (blobs == 0) return cudaSuccess;
    const size_t grid = blobs < (size_t)lc.sm_count * 32 ? blobs : (size_t)lc.sm_count * 32;
    snapshot_transpose_kernel<<<(unsigned)grid, TCTA, 0, lc.stream>>>(in, d_in_off, P, n, out, d_out_off);
    ++*lc.nlaunch;
    return cudaGetLastError();
}

}  // namespace sda
