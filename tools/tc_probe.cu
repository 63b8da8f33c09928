// tc_probe.cu -- developer probe (not part of the product): one tcgen05.mma kind::i8 tile
// D[128 x 48] (s32, TMEM) = A[128 x 64] (u8, smem, K-major, no swizzle) . B[48 x 64]^T, checked on the host.
// Pins down the shared-memory descriptor / instruction descriptor / TMEM addressing used by packed_tc.cu.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int M = 128, N = 48, K = 64;
constexpr uint32_t LBO = 128, SBO = 512;    // bytes: next 16-byte K chunk, next 8-row group

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t taddr, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        :: "r"(taddr), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}

__global__ void __launch_bounds__(128) probe(const uint8_t *A, const uint8_t *B, int32_t *D) {
    __shared__ __align__(128) uint8_t sA[(M / 8) * SBO];
    __shared__ __align__(128) uint8_t sB[(N / 8) * SBO];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int c = 0; c < 4; c++)
        *reinterpret_cast<uint4 *>(sA + (tid / 8) * SBO + c * LBO + (tid % 8) * 16) =
            *reinterpret_cast<const uint4 *>(A + tid * K + c * 16);
    if (tid < N)
        for (int c = 0; c < 4; c++)
            *reinterpret_cast<uint4 *>(sB + (tid / 8) * SBO + c * LBO + (tid % 8) * 16) =
                *reinterpret_cast<const uint4 *>(B + tid * K + c * 16);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // s32 acc, u8 x u8, K-major both
        const uint64_t da = make_desc(smem_u32(sA)), db = make_desc(smem_u32(sB));
        mma_i8(taddr, da, db, idesc, 0);
        mma_i8(taddr, da + ((2 * LBO) >> 4), db + ((2 * LBO) >> 4), idesc, 1);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[48];
    const uint32_t ta = taddr + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                   "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                   "=r"(v[30]), "=r"(v[31]) : "r"(ta));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]),
                   "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]) : "r"(ta + 32));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < N; n++) D[tid * N + n] = (int32_t)v[n];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(taddr) : "memory");
}

int main() {
    uint8_t hA[M * K], hB[N * K];
    srand(1);
    for (auto &x : hA) x = rand() & 255;
    for (auto &x : hB) x = rand() & 255;
    uint8_t *dA, *dB; int32_t *dD; static int32_t hD[M * N];
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, sizeof hD);
    probe<<<1, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            int32_t ref = 0;
            for (int k = 0; k < K; k++) ref += (int32_t)hA[m * K + k] * hB[n * K + k];
            if (ref != hD[m * N + n] && bad++ < 10) printf("mismatch m=%d n=%d got %d want %d\n", m, n, hD[m * N + n], ref);
        }
    printf("%s: %d mismatches of %d\n", bad ? "FAIL" : "PASS", bad, M * N);
    return bad != 0;
}
