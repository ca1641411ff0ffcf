// Developer aid: issue-rate microbenchmarks that size the roofline of the FIR kernels on B200.
// Each kernel runs a long unrolled chain of one instruction mix per thread; reports warp-instructions/clk/SM.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 4096   // x8 unrolled inner repeats
template<int MODE>
__global__ void mix_kernel(int* out, int a0, int b0, float f0)
{
    int x0 = threadIdx.x + a0, x1 = x0 * 3, x2 = x0 ^ 5, x3 = x0 + 7, x4 = x0 - 1, x5 = x0 * 7, x6 = x0 + 11, x7 = x0 ^ 9;
    float g0 = f0 + threadIdx.x, g1 = g0 * 1.1f, g2 = g0 + 2.f, g3 = g0 - 3.f, g4 = g0 * 0.5f, g5 = g0 + 5.f, g6 = g0 - 6.f, g7 = g0 * 7.f;
    double d0 = g0, d1 = g1, d2 = g2, d3 = g3;
    const int b = b0;
#pragma unroll 1
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (MODE == 0) {        // IMAD reg*imm+reg x8
            x0 = x0 * 1300 + x1; x1 = x1 * -424 + x2; x2 = x2 * 244 + x3; x3 = x3 * -164 + x4;
            x4 = x4 * 117 + x5; x5 = x5 * -86 + x6; x6 = x6 * 64 + x7; x7 = x7 * -47 + x0;
        } else if (MODE == 1) { // IADD x8 (lop to stop folding)
            asm volatile("add.s32 %0, %0, %1;" : "+r"(x0) : "r"(x1)); asm volatile("add.s32 %0, %0, %1;" : "+r"(x1) : "r"(x2));
            asm volatile("add.s32 %0, %0, %1;" : "+r"(x2) : "r"(x3)); asm volatile("add.s32 %0, %0, %1;" : "+r"(x3) : "r"(x4));
            asm volatile("add.s32 %0, %0, %1;" : "+r"(x4) : "r"(x5)); asm volatile("add.s32 %0, %0, %1;" : "+r"(x5) : "r"(x6));
            asm volatile("add.s32 %0, %0, %1;" : "+r"(x6) : "r"(x7)); asm volatile("add.s32 %0, %0, %1;" : "+r"(x7) : "r"(x0));
        } else if (MODE == 2) { // 4 IMAD + 4 IADD interleaved
            x0 = x0 * 1300 + x1; asm volatile("add.s32 %0, %0, %1;" : "+r"(x1) : "r"(x2));
            x2 = x2 * 244 + x3;  asm volatile("add.s32 %0, %0, %1;" : "+r"(x3) : "r"(x4));
            x4 = x4 * 117 + x5;  asm volatile("add.s32 %0, %0, %1;" : "+r"(x5) : "r"(x6));
            x6 = x6 * 64 + x7;   asm volatile("add.s32 %0, %0, %1;" : "+r"(x7) : "r"(x0));
        } else if (MODE == 3) { // FFMA reg*imm+reg x8
            g0 = fmaf(g0, 0.3175f, g1); g1 = fmaf(g1, -0.1037f, g2); g2 = fmaf(g2, 0.0597f, g3); g3 = fmaf(g3, -0.0401f, g4);
            g4 = fmaf(g4, 0.0287f, g5); g5 = fmaf(g5, -0.0211f, g6); g6 = fmaf(g6, 0.0157f, g7); g7 = fmaf(g7, -0.0116f, g0);
        } else if (MODE == 4) { // FADD x8
            g0 = __fadd_rn(g0, g1); g1 = __fadd_rn(g1, g2); g2 = __fadd_rn(g2, g3); g3 = __fadd_rn(g3, g4);
            g4 = __fadd_rn(g4, g5); g5 = __fadd_rn(g5, g6); g6 = __fadd_rn(g6, g7); g7 = __fadd_rn(g7, g0);
        } else if (MODE == 5) { // dp2a x8
            asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x0) : "r"(x1), "r"(b)); asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x1) : "r"(x2), "r"(b));
            asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x2) : "r"(x3), "r"(b)); asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x3) : "r"(x4), "r"(b));
            asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x4) : "r"(x5), "r"(b)); asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x5) : "r"(x6), "r"(b));
            asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x6) : "r"(x7), "r"(b)); asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x7) : "r"(x0), "r"(b));
        } else if (MODE == 6) { // DFMA x4
            d0 = fma(d0, 1.0000001, d1); d1 = fma(d1, 0.9999999, d2); d2 = fma(d2, 1.0000002, d3); d3 = fma(d3, 0.9999998, d0);
        } else if (MODE == 7) { // IMAD reg*reg+reg x8
            x0 = x0 * b + x1; x1 = x1 * b + x2; x2 = x2 * b + x3; x3 = x3 * b + x4;
            x4 = x4 * b + x5; x5 = x5 * b + x6; x6 = x6 * b + x7; x7 = x7 * b + x0;
        } else if (MODE == 8) { // 4 FFMA + 4 IADD
            g0 = fmaf(g0, 0.3175f, g1); asm volatile("add.s32 %0, %0, %1;" : "+r"(x1) : "r"(x2));
            g2 = fmaf(g2, 0.0597f, g3); asm volatile("add.s32 %0, %0, %1;" : "+r"(x3) : "r"(x4));
            g4 = fmaf(g4, 0.0287f, g5); asm volatile("add.s32 %0, %0, %1;" : "+r"(x5) : "r"(x6));
            g6 = fmaf(g6, 0.0157f, g7); asm volatile("add.s32 %0, %0, %1;" : "+r"(x7) : "r"(x0));
        } else if (MODE == 9) { // 4 FFMA + 4 FADD
            g0 = fmaf(g0, 0.3175f, g1); g1 = __fadd_rn(g1, g2); g2 = fmaf(g2, 0.0597f, g3); g3 = __fadd_rn(g3, g4);
            g4 = fmaf(g4, 0.0287f, g5); g5 = __fadd_rn(g5, g6); g6 = fmaf(g6, 0.0157f, g7); g7 = __fadd_rn(g7, g0);
        } else if (MODE == 10) { // 6 IMAD + 2 IADD
            x0 = x0 * 1300 + x1; x1 = x1 * -424 + x2; x2 = x2 * 244 + x3; asm volatile("add.s32 %0, %0, %1;" : "+r"(x3) : "r"(x4));
            x4 = x4 * 117 + x5; x5 = x5 * -86 + x6; x6 = x6 * 64 + x7; asm volatile("add.s32 %0, %0, %1;" : "+r"(x7) : "r"(x0));
        } else if (MODE == 11) { // HFMA2 x8 (packed half)
            asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x0) : "r"(b), "r"(x1)); asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x1) : "r"(b), "r"(x2));
            asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x2) : "r"(b), "r"(x3)); asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x3) : "r"(b), "r"(x4));
            asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x4) : "r"(b), "r"(x5)); asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x5) : "r"(b), "r"(x6));
            asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x6) : "r"(b), "r"(x7)); asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x7) : "r"(b), "r"(x0));
        } else if (MODE == 12) { // 4 IMAD + 4 LOP3
            x0 = x0 * 1300 + x1; asm volatile("xor.b32 %0, %0, %1;" : "+r"(x1) : "r"(x2));
            x2 = x2 * 244 + x3;  asm volatile("xor.b32 %0, %0, %1;" : "+r"(x3) : "r"(x4));
            x4 = x4 * 117 + x5;  asm volatile("xor.b32 %0, %0, %1;" : "+r"(x5) : "r"(x6));
            x6 = x6 * 64 + x7;   asm volatile("xor.b32 %0, %0, %1;" : "+r"(x7) : "r"(x0));
        } else if (MODE == 13) { // 4 DFMA(2 chains) + 8 IMAD  : is FP64 a separate pipe?
            d0 = fma(d0, 1.0000001, d1); d1 = fma(d1, 0.9999999, d0);
            x0 = x0 * 1300 + x1; x1 = x1 * -424 + x2; x2 = x2 * 244 + x3; x3 = x3 * -164 + x4;
            x4 = x4 * 117 + x5; x5 = x5 * -86 + x6; x6 = x6 * 64 + x7; x7 = x7 * -47 + x0;
        }
      }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (int) (g0 + g1 + g2 + g3 + g4 + g5 + g6 + g7) + (int) (d0 + d1 + d2 + d3);
}

template<int MODE> void run(const char* name, int per_iter, int* d_out, int sms)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 1024, blocks = sms * 2;
    mix_kernel<MODE><<<blocks, threads>>>(d_out, 1, 0x01020304, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    mix_kernel<MODE><<<blocks, threads>>>(d_out, 1, 0x01020304, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double winstr = (double) blocks * (threads / 32) * ITERS * per_iter;
    double per_s = winstr / (ms * 1e-3);
    printf("%-28s %8.3f ms  %7.2f Gwarp-instr/s  = %5.2f warp-instr/clk/SM @%d MHz(max)  lanes/clk/SM=%6.1f\n", name, ms, per_s / 1e9,
           per_s / sms / (clk * 1e3), clk / 1000, 32 * per_s / sms / (clk * 1e3));
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int* d_out; cudaMalloc(&d_out, sizeof(int) * 1024 * sms * 2);
    printf("SMs=%d\n", sms);
    run<0>("IMAD imm x8", 8, d_out, sms);
    run<7>("IMAD reg x8", 8, d_out, sms);
    run<1>("IADD x8", 8, d_out, sms);
    run<2>("4 IMAD + 4 IADD", 8, d_out, sms);
    run<10>("6 IMAD + 2 IADD", 8, d_out, sms);
    run<12>("4 IMAD + 4 LOP3", 8, d_out, sms);
    run<3>("FFMA imm x8", 8, d_out, sms);
    run<4>("FADD x8", 8, d_out, sms);
    run<9>("4 FFMA + 4 FADD", 8, d_out, sms);
    run<8>("4 FFMA + 4 IADD", 8, d_out, sms);
    run<5>("DP2A x8", 8, d_out, sms);
    run<11>("HFMA2 x8", 8, d_out, sms);
    run<6>("DFMA x4", 4, d_out, sms);
    run<13>("2 DFMA + 8 IMAD", 10, d_out, sms);
    return 0;
}
