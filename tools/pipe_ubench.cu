// pipe_ubench.cu -- developer microbenchmark (not part of the product): issue rates of the
// integer instructions the share-gen kernel is made of, on one B200.  Prints lane-ops/clk/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N_ITER 4096
#define ILP 8

template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t *out, uint32_t a, uint32_t b, long long *cyc) {
    uint64_t acc[ILP];
    uint32_t x[ILP], y[ILP];
    for (int i = 0; i < ILP; i++) { acc[i] = threadIdx.x + i; x[i] = a + i * 7 + threadIdx.x; y[i] = b ^ (i * 13) ^ (threadIdx.x * 3); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N_ITER; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) {  // IMAD.WIDE.U32
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x[i]), "r"(y[i]));
            } else if (MODE == 1) {  // IMAD 32
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(a));
            } else if (MODE == 2) {  // IADD3-ish
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            } else if (MODE == 3) {  // xor + rotate (LOP3 + SHF)
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
                asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(x[i]));
            } else if (MODE == 4) {  // 1 wide mad + 1 xor + 1 rot
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x[i]), "r"(y[i]));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            } else if (MODE == 5) {  // signed wide
                asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x[i]), "r"(y[i]));
            } else if (MODE == 6) {  // chacha-like QR mix: add xor rot
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
                asm volatile("xor.b32 %1, %1, %0;" : "+r"(x[i]), "+r"(y[i]));
                asm volatile("shf.l.wrap.b32 %0, %0, %0, 12;" : "+r"(y[i]));
            } else if (MODE == 7) {  // mad.lo used as add (a*1+b) + xor + rot
                asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(x[i]) : "r"(y[i]));
                asm volatile("xor.b32 %1, %1, %0;" : "+r"(x[i]), "+r"(y[i]));
                asm volatile("shf.l.wrap.b32 %0, %0, %0, 12;" : "+r"(y[i]));
            } else if (MODE == 8) {  // 64-bit add (IADD3 + IADD3.X)
                asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[i]) : "l"((uint64_t)x[i] << 20 | y[i]));
            } else if (MODE == 9) {  // prmt
                asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(x[i]) : "r"(y[i]));
            }
        }
    }
    long long t1 = clock64();
    uint64_t s = 0;
    for (int i = 0; i < ILP; i++) s += acc[i] + x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    (void)t0;
}

__global__ void clk_kernel(long long *out) {
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    long long c0 = clock64();
    while (clock64() - c0 < 2000000) {}
    long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    out[0] = c1 - c0; out[1] = (long long)(g1 - g0);
}
double sm_clock_mhz() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    clk_kernel<<<1, 1>>>(d);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return (double)h[0] / (double)h[1] * 1e3;
}

template <int MODE>
void run(const char *name, int ops_per_inner, int ctas_per_sm) {
    uint64_t *out; long long *cyc, h;
    int sms = 148;
    cudaMalloc(&out, sizeof(uint64_t) * 256 * sms * ctas_per_sm);
    cudaMalloc(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * ctas_per_sm, 256>>>(out, 3, 5, cyc);
    cudaEventRecord(e0);
    k<MODE><<<sms * ctas_per_sm, 256>>>(out, 3, 5, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double lane_ops = (double)N_ITER * ILP * ops_per_inner * 256 * ctas_per_sm;   // per SM
    double clk_mhz = sm_clock_mhz();
    printf("%-28s ctas/SM=%d  %.3f ms  lane-ops/clk/SM=%.1f  (clk %.0f MHz)\n", name, ctas_per_sm, ms,
           lane_ops / (ms * 1e-3 * clk_mhz * 1e6), clk_mhz);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int c : {2, 4}) {
        run<0>("mad.wide.u32", 1, c);
        run<5>("mad.wide.s32", 1, c);
        run<1>("mad.lo.u32", 1, c);
        run<2>("add.u32", 1, c);
        run<3>("xor+shf", 2, c);
        run<9>("prmt", 1, c);
        run<4>("mad.wide + xor", 2, c);
        run<6>("add+xor+shf", 3, c);
        run<7>("mad.lo(add)+xor+shf", 3, c);
        run<8>("add.u64", 1, c);
    }
    return 0;
}
