// pipe_ubench2.cu -- developer microbenchmark (not part of the product): can the ALU pipe
// (LOP3/SHF/IADD3) and the FMA-heavy pipe (IMAD, IMAD.WIDE) of a B200 SM sub-partition run
// concurrently, from one warp's instruction stream and from different warps?
// Prints warp-instructions per clock per SM sub-partition (SMSP).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N_ITER 2048

// NW wide-mads and NA alu ops per inner step, all independent chains (8 accumulators / 8 regs)
template <int NW, int NA, int NI>
__global__ void __launch_bounds__(256) mix(uint64_t *out, uint32_t a, uint32_t b, int role_split) {
    uint64_t acc[8];
    uint32_t x[8], y[8], z[8];
    for (int i = 0; i < 8; i++) { acc[i] = threadIdx.x + i; x[i] = a + i * 7 + threadIdx.x; y[i] = b ^ (i * 13) ^ (threadIdx.x * 3); z[i] = x[i] * 3 + 1; }
    const int warp = threadIdx.x >> 5;
    // role_split: 0 = every warp runs the mixed stream; 1 = even warps wide-mads only, odd warps alu only
    const bool do_w = role_split == 0 || (warp & 1) == 0;
    const bool do_a = role_split == 0 || (warp & 1) == 1;
#pragma unroll 1
    for (int it = 0; it < N_ITER; it++) {
        if (do_w) {
#pragma unroll
            for (int i = 0; i < NW; i++) asm volatile("{.reg .u64 t; mul.wide.u32 t, %1, %2; add.u64 %0, %0, t;}" : "+l"(acc[i % 8]) : "r"(x[i % 8]), "r"(b));
#pragma unroll
            for (int i = 0; i < NI; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(z[i % 8]) : "r"(a), "r"(b));
        }
        if (do_a) {
#pragma unroll
            for (int i = 0; i < NA; i++) {
                if (i & 1) asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(y[(i / 2) % 8]));
                else asm volatile("xor.b32 %0, %0, %1;" : "+r"(y[(i / 2) % 8]) : "r"(a));
            }
        }
    }
    uint64_t s = 0;
    for (int i = 0; i < 8; i++) s += acc[i] + x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

double sm_clock_mhz();
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


// rotate implemented on the FMA pipe: lo = x << r (IMAD.SHL), rot = hi32(x * 2^r) + lo (IMAD.HI.U32)
template <int MODE>
__global__ void __launch_bounds__(256) rotk(uint64_t *out, uint32_t a, uint32_t b) {
    uint32_t x[8];
    for (int i = 0; i < 8; i++) x[i] = a + i * 7 + threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < N_ITER; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i % 8]) : "r"(b), "r"(a));
            if (MODE == 1) asm volatile("{.reg .u32 t; mul.lo.u32 t, %0, %1; mad.hi.u32 %0, %0, %1, t;}" : "+r"(x[i % 8]) : "r"(b));
            if (MODE == 2) asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(x[i % 8]));
            if (MODE == 3) {   // half the rotates on each pipe + one xor each (ChaCha-like mix)
                if (i & 1) asm volatile("{.reg .u32 t; mul.lo.u32 t, %0, %1; mad.hi.u32 %0, %0, %1, t;}" : "+r"(x[i % 8]) : "r"(b));
                else asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(x[i % 8]));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i % 8]) : "r"(a));
            }
            if (MODE == 4) {   // all rotates on the ALU pipe + one xor each
                asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(x[i % 8]));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i % 8]) : "r"(a));
            }
        }
    }
    uint64_t s = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run_rot(const char *name, int ctas_per_sm) {
    uint64_t *out;
    const int sms = 148;
    cudaMalloc(&out, sizeof(uint64_t) * 256 * sms * ctas_per_sm);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rotk<MODE><<<sms * ctas_per_sm, 256>>>(out, 3, 128);
    cudaEventRecord(e0);
    rotk<MODE><<<sms * ctas_per_sm, 256>>>(out, 3, 128);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * sm_clock_mhz() * 1e6;
    const double steps = (double)N_ITER * 16 * (8.0 * ctas_per_sm / 4.0);
    printf("%-44s ctas/SM=%d  cycles per inner step per SMSP: %.2f\n", name, ctas_per_sm, cyc / steps);
    cudaFree(out);
}

template <int NW, int NA, int NI>
void run(const char *name, int split, int ctas_per_sm) {
    uint64_t *out;
    const int sms = 148;
    cudaMalloc(&out, sizeof(uint64_t) * 256 * sms * ctas_per_sm);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<NW, NA, NI><<<sms * ctas_per_sm, 256>>>(out, 3, 5, split);
    cudaEventRecord(e0);
    mix<NW, NA, NI><<<sms * ctas_per_sm, 256>>>(out, 3, 5, split);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double clk = sm_clock_mhz() * 1e6;
    const double warps_per_smsp = 8.0 * ctas_per_sm / 4.0;
    const double frac = split ? 0.5 : 1.0;
    const double cyc = ms * 1e-3 * clk;
    const double w = N_ITER * NW * warps_per_smsp * frac, al = N_ITER * NA * warps_per_smsp * frac, im = N_ITER * NI * warps_per_smsp * frac;
    printf("%-34s split=%d ctas/SM=%d  cycles=%.0f  per clk per SMSP: IMAD.WIDE %.3f  IMAD %.3f  ALU %.3f  | pipe-cycles if W=4,I=2,A=2: fma %.2f alu %.2f\n",
           name, split, ctas_per_sm, cyc, w / cyc, im / cyc, al / cyc, (4 * w + 2 * im) / cyc, 2 * al / cyc);
    cudaFree(out);
}

int main() {
    for (int c : {1, 4}) {
        run<8, 0, 0>("wide only", 0, c);
        run<0, 16, 0>("alu only", 0, c);
        run<0, 0, 8>("imad only", 0, c);
        run<8, 16, 0>("wide 8 + alu 16 (1:1 cycles)", 0, c);
        run<8, 8, 0>("wide 8 + alu 8", 0, c);
        run<8, 32, 0>("wide 8 + alu 32", 0, c);
        run<0, 16, 16>("imad 16 + alu 16", 0, c);
        run<4, 16, 8>("wide 4 + imad 8 + alu 16", 0, c);
        run<8, 16, 0>("split warps: wide 8 | alu 16", 1, c);
        run<8, 32, 0>("split warps: wide 8 | alu 32", 1, c);
        run<0, 16, 16>("split warps: imad 16 | alu 16", 1, c);
    }
    for (int c : {4}) {
        run_rot<0>("mad.hi.u32 (IMAD.HI)", c);
        run_rot<1>("rotate = mul.lo + mad.hi (2 FMA-pipe)", c);
        run_rot<2>("rotate = shf (ALU)", c);
        run_rot<3>("xor + rotate, rotates split ALU/FMA 1:1", c);
        run_rot<4>("xor + rotate, all ALU", c);
    }
    return 0;
}
