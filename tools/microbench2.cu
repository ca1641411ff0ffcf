// Developer aid: pipe-rate / co-issue microbenchmarks (independent chains, SASS-checked opcodes), B200.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define REP 512

#define LOP(x)  asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x) : "r"(b), "r"(c))
#define SHFI(x) asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(x) : "r"(b))
#define IMAD(x) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(c), "r"(b))
#define IMADI(x) asm volatile("mad.lo.s32 %0, %0, 1300, %1;" : "+r"(x) : "r"(b))
#define IADD3(x) asm volatile("{.reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c))
#define IADD(x) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(b))
#define FFMA(g) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(g) : "f"(fb), "f"(fc))
#define FADD(g) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(g) : "f"(fb))

template<int MODE>
__global__ void k(int* out, int b, int c, float fb, float fc, long long* clk)
{
    int x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    float g0 = x0, g1 = x1, g2 = x2, g3 = x3, g4 = x4, g5 = x5, g6 = x6, g7 = x7;
    long long t0 = 0, n0 = 0;
    if (threadIdx.x == 0 && blockIdx.x == 0) { t0 = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(n0)); }
#pragma unroll 1
    for (int it = 0; it < REP; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0) { LOP(x0); LOP(x1); LOP(x2); LOP(x3); LOP(x4); LOP(x5); LOP(x6); LOP(x7); }
            if (MODE == 1) { SHFI(x0); SHFI(x1); SHFI(x2); SHFI(x3); SHFI(x4); SHFI(x5); SHFI(x6); SHFI(x7); }
            if (MODE == 2) { IMAD(x0); IMAD(x1); IMAD(x2); IMAD(x3); IMAD(x4); IMAD(x5); IMAD(x6); IMAD(x7); }
            if (MODE == 3) { IADD3(x0); IADD3(x1); IADD3(x2); IADD3(x3); IADD3(x4); IADD3(x5); IADD3(x6); IADD3(x7); }
            if (MODE == 4) { FFMA(g0); FFMA(g1); FFMA(g2); FFMA(g3); FFMA(g4); FFMA(g5); FFMA(g6); FFMA(g7); }
            if (MODE == 5) { FFMA(g0); LOP(x0); FFMA(g1); LOP(x1); FFMA(g2); LOP(x2); FFMA(g3); LOP(x3); }
            if (MODE == 6) { FFMA(g0); FFMA(g1); LOP(x0); FFMA(g2); FFMA(g3); LOP(x1); FFMA(g4); FFMA(g5); LOP(x2); FFMA(g6); FFMA(g7); LOP(x3); }
            if (MODE == 7) { IMAD(x0); LOP(x4); IMAD(x1); LOP(x5); IMAD(x2); LOP(x6); IMAD(x3); LOP(x7); }
            if (MODE == 8) { FFMA(g0); IMAD(x0); FFMA(g1); IMAD(x1); FFMA(g2); IMAD(x2); FFMA(g3); IMAD(x3); }
            if (MODE == 9) { FFMA(g0); IMAD(x0); LOP(x4); FFMA(g1); IMAD(x1); LOP(x5); FFMA(g2); IMAD(x2); LOP(x6); FFMA(g3); IMAD(x3); LOP(x7); }
            if (MODE == 10) { IMAD(x0); IADD3(x4); IMAD(x1); IADD3(x5); IMAD(x2); IADD3(x6); IMAD(x3); IADD3(x7); }
            if (MODE == 11) { IADD(x0); IADD(x1); IADD(x2); IADD(x3); IADD(x4); IADD(x5); IADD(x6); IADD(x7); }
            if (MODE == 12) { IMADI(x0); IMADI(x1); IMADI(x2); IMADI(x3); IMADI(x4); IMADI(x5); IMADI(x6); IMADI(x7); }
            if (MODE == 13) { FADD(g0); LOP(x0); FADD(g1); LOP(x1); FADD(g2); LOP(x2); FADD(g3); LOP(x3); }
            if (MODE == 14) { IMAD(x0); IMAD(x1); LOP(x4); IMAD(x2); IMAD(x3); LOP(x5); }
            if (MODE == 15) { IMAD(x0); IMAD(x1); FFMA(g0); FFMA(g1); LOP(x4); LOP(x5); IMAD(x2); IMAD(x3); FFMA(g2); FFMA(g3); LOP(x6); LOP(x7); }
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) { long long n1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(n1)); clk[0] = clock64() - t0; clk[1] = n1 - n0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (int) (g0 + g1 + g2 + g3 + g4 + g5 + g6 + g7);
}

template<int MODE> void run(const char* name, int per_u, int* d_out, long long* d_clk, int sms)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 1024, blocks = sms * 2;
    for (int i = 0; i < 3; ++i) k<MODE><<<blocks, threads>>>(d_out, 0x12345, 0x6789a, 0.999f, 0.001f, d_clk);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d_out, 0x12345, 0x6789a, 0.999f, 0.001f, d_clk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[2]; cudaMemcpy(h, d_clk, 16, cudaMemcpyDeviceToHost);
    double ghz = (double) h[0] / (double) h[1];
    double winstr = (double) blocks * (threads / 32) * REP * 4 * per_u;
    double per_clk_sm = winstr / sms / (double) h[0];     // block 0's own cycle count ~ kernel duration in SM clocks
    printf("%-34s %7.3f ms  clk=%.3f GHz  %5.2f warp-instr/clk/SM (by clock64)  %5.2f (by events @clk)\n", name, ms, ghz, per_clk_sm,
           winstr / sms / (ms * 1e-3 * ghz * 1e9));
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int* d_out; cudaMalloc(&d_out, sizeof(int) * 1024 * sms * 2);
    long long* d_clk; cudaMalloc(&d_clk, 16);
    run<0>("LOP3 x8", 8, d_out, d_clk, sms);
    run<1>("SHF x8", 8, d_out, d_clk, sms);
    run<2>("IMAD reg x8", 8, d_out, d_clk, sms);
    run<12>("IMAD imm x8", 8, d_out, d_clk, sms);
    run<3>("IADD3 (3-input) x8", 8, d_out, d_clk, sms);
    run<11>("add.s32 x8 (ptxas picks)", 8, d_out, d_clk, sms);
    run<4>("FFMA x8", 8, d_out, d_clk, sms);
    run<5>("FFMA:LOP3 1:1", 8, d_out, d_clk, sms);
    run<13>("FADD:LOP3 1:1", 8, d_out, d_clk, sms);
    run<6>("FFMA:LOP3 2:1", 12, d_out, d_clk, sms);
    run<7>("IMAD:LOP3 1:1", 8, d_out, d_clk, sms);
    run<14>("IMAD:LOP3 2:1", 6, d_out, d_clk, sms);
    run<10>("IMAD:IADD3 1:1", 8, d_out, d_clk, sms);
    run<8>("FFMA:IMAD 1:1", 8, d_out, d_clk, sms);
    run<9>("FFMA:IMAD:LOP3 1:1:1", 12, d_out, d_clk, sms);
    run<15>("IMAD:FFMA:LOP3 1:1:1 (pairs)", 12, d_out, d_clk, sms);
    return 0;
}
