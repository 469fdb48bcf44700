// pipe_bench.cu -- issue throughput of the integer instructions the scan kernels are made of (per SM and clock),
// alone and mixed, to decide which shifts can move from the ALU pipe (SHF, LOP3) to the FMA pipe (IMAD, IMAD.WIDE).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_bench tools/pipe_bench.cu ; run: ./pipe_bench
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
#define UNR 8
template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned *out, unsigned mul, unsigned seed)
{
    unsigned a[UNR], b[UNR];
    unsigned long long w[UNR];
#pragma unroll
    for (int i = 0; i < UNR; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i; w[i] = a[i]; }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < UNR; i++) {
            if (MODE == 0) a[i] = __funnelshift_l(b[i], a[i], 5);                                   // SHF
            if (MODE == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(mul));   // LOP3
            if (MODE == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(mul), "r"(b[i]));      // IMAD
            if (MODE == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(mul));    // IMAD.WIDE
            if (MODE == 4) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(mul));                     // IMAD.HI
            if (MODE == 5) {                                                                                     // SHF + IMAD.WIDE, 1:1
                a[i] = __funnelshift_l(b[i], a[i], 5);
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(mul));
            }
            if (MODE == 6) {                                                                                     // LOP3 + IMAD, 1:1
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(mul));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(mul), "r"(a[i]));
            }
            if (MODE == 7) {                                                                                     // 2 LOP3 + 1 IMAD.WIDE
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(mul));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(mul));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(mul));
            }
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < UNR; i++) r ^= a[i] ^ b[i] ^ (unsigned)w[i] ^ (unsigned)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char *name, int per_iter)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    unsigned *out; cudaMalloc(&out, sms * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(out, 32, 1);
    cudaEventRecord(e0);
    k<MODE><<<sms * 8, 256>>>(out, 32, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)sms * 8 * 8 * ITER * UNR * per_iter;
    printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (at %d MHz)\n", name, ms, warp_instr / (ms * 1e-3) / (khz * 1e3) / sms, khz / 1000);
    cudaFree(out);
}
int main()
{
    run<0>("SHF", 1); run<1>("LOP3", 1); run<2>("IMAD", 1); run<3>("IMAD.WIDE", 1); run<4>("IMAD.HI", 1);
    run<5>("SHF+IMAD.WIDE", 2); run<6>("LOP3+IMAD", 2); run<7>("2 LOP3 + IMAD.WIDE", 3);
    return 0;
}
