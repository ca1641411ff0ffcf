// spectrum_api.cu — K5: SpectrumVis as a batched shared-memory FFT kernel + C ABI b200dsp_spectrum_*
//
// Replaces (paths relative to the reference tree):
//   SpectrumVis::feed / handleConfigure            sdrgui/dsp/spectrumvis.cpp:77-254,283-327
//   FFTWindow::create / apply                      sdrbase/dsp/fftwindow.cpp:20-73, fftwindow.h:52-84
//   FFTEngine / KissEngine::transform (kissfft)    sdrbase/dsp/kissengine.cpp:3-25, kissfft.h:44-238
//   MovingAverage2D<double> / FixedAverage2D<double>  sdrbase/util/movingaverage2d.h:40-109, fixedaverage2d.h:36-113
//
// Per frame of N samples: x[i] = (re/scalef, im/scalef) * w[i];  X = DFT_N(x) (unnormalised);  v[b] = |X[b]|^2;
// averaging in double; value = linear ? v / N^2 : (10/log2f(10)) * log2f(v) + 20*log10f(1/N);  output DC-centred
// (half swap), or with positiveOnly the first N/2 bins each written twice.
//
// B200 design: one CTA per OUTPUT frame.  The CTA walks the input frames that feed its output (1 for no/moving
// averaging, averageNb for fixed averaging), and for each: loads the int16 IQ samples straight into digit-reversed
// positions of a shared-memory buffer (scale and window fused into the load), runs in-place radix-4 passes (+ one
// radix-2 pass for odd log2 N) with twiddles from a device table, and accumulates |X|^2 per bin in double registers.
// Fixed averaging therefore never writes an intermediate spectrum to HBM (4.4 bytes per input sample end to end).
// Moving averaging writes each frame's power once (float) and a second small kernel forms the sliding sums in double.
// The FFT factorisation differs from KissFFT's, so results agree to float32 rounding (tolerance in the tests), not bitwise.
#include "common.cuh"
#include <math.h>
#include <vector>

using namespace b200dsp;

namespace {

constexpr int SPEC_MAX_N = 4096;                    // MAX_FFT_SIZE (spectrumvis.h)
constexpr int SPEC_THREADS = 256;

enum { AVG_NONE = 0, AVG_MOVING = 1, AVG_FIXED = 2 };

struct SpecParams {
    const uint32_t* in;        // new samples of this feed (packed int16 IQ)
    const uint32_t* partial;   // carried partial frame (fill samples)
    const float*    window;    // [n]
    const float2*   tw;        // [SPEC_MAX_N] exp(-2 pi i t / SPEC_MAX_N)
    const float2*   tw1;       // 4096-point kernel: [16][256] W_4096^(n2 k1) laid out so that a warp's loads are contiguous, then [16][16] W_256^(m2 j1)
    float*          out;       // [frames_out][n]
    float*          power;     // moving mode: [hist + frames][n] raw |X|^2 of every frame (hist = avg_nb - 1 carried frames first)
    const double*   fix_sum;   // fixed mode: carried partial sums [n] of the previous feed (read by CTA 0)
    double*         fix_sum_out;   // ... of this feed (written by CTA save_sums): the other half of a ping-pong pair
    int n, log2n;
    int fill;                  // samples in `partial`
    int frames;                // input frames completed by this feed
    int mode, avg_nb, linear, positive_only;
    int fix_idx;               // fixed mode: frames already accumulated in fix_sum
    int frames_out;
    int save_sums;             // fixed mode: CTA index that must write its sums back (the trailing incomplete group), -1 none
    float scalef, ofs, div, mult;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }

// digit reversal for a size-n transform computed as [radix-2 split if log2n odd] x radix-4 DIT passes
__device__ __forceinline__ int rev4(int i, int digits)
{
    int r = 0;
    for (int d = 0; d < digits; ++d) { r = (r << 2) | (i & 3); i >>= 2; }
    return r;
}

__device__ __forceinline__ float spec_value(const SpecParams& p, float v)
{
    return p.linear ? v / p.div : p.mult * log2f(v) + p.ofs;
}

// transforms the frame held (digit-reversed) in X[0..n) in place; result in natural order
__device__ void fft_inplace(float2* X, const SpecParams& p)
{
    const int n = p.n;
    const int odd = p.log2n & 1;
    const int n4 = odd ? n >> 1 : n;                 // size of the radix-4 sub-transforms
    const int tstride0 = SPEC_MAX_N / n4;
    for (int m = 1; m < n4; m <<= 2) {               // butterflies of span m inside groups of 4m
        const int tw_step = tstride0 * (n4 / (4 * m));
        for (int q = threadIdx.x; q < n / 4; q += blockDim.x) {
            const int half = odd ? (q / (n4 / 4)) : 0;          // which sub-transform
            const int qq = odd ? (q - half * (n4 / 4)) : q;
            const int u = qq & (m - 1);
            const int g = (qq - u) << 2;
            float2* x = X + half * n4 + g + u;
            const float2 a0 = x[0];
            float2 a1 = x[m], a2 = x[2 * m], a3 = x[3 * m];
            if (m > 1) {
                const float2 w1 = p.tw[u * tw_step], w2 = p.tw[2 * u * tw_step], w3 = p.tw[3 * u * tw_step];
                a1 = cmul(a1, w1); a2 = cmul(a2, w2); a3 = cmul(a3, w3);
            }
            const float2 s02 = make_float2(a0.x + a2.x, a0.y + a2.y), d02 = make_float2(a0.x - a2.x, a0.y - a2.y);
            const float2 s13 = make_float2(a1.x + a3.x, a1.y + a3.y), d13 = make_float2(a1.x - a3.x, a1.y - a3.y);
            x[0]     = make_float2(s02.x + s13.x, s02.y + s13.y);
            x[2 * m] = make_float2(s02.x - s13.x, s02.y - s13.y);
            x[m]     = make_float2(d02.x + d13.y, d02.y - d13.x);     // a0 - j a1 - a2 + j a3
            x[3 * m] = make_float2(d02.x - d13.y, d02.y + d13.x);     // a0 + j a1 - a2 - j a3
        }
        __syncthreads();
    }
    if (odd) {                                        // X[k] = E[k] + W_n^k O[k], X[k + n/2] = E[k] - W_n^k O[k]
        const int ts = SPEC_MAX_N / n;
        for (int k = threadIdx.x; k < n4; k += blockDim.x) {
            const float2 e = X[k], o = cmul(X[k + n4], p.tw[k * ts]);
            X[k] = make_float2(e.x + o.x, e.y + o.y);
            X[k + n4] = make_float2(e.x - o.x, e.y - o.y);
        }
        __syncthreads();
    }
}

// load frame f of the virtual stream [partial ++ in] into X (digit-reversed), scaled and windowed
__device__ void load_frame(float2* X, const SpecParams& p, long long f)
{
    const int n = p.n;
    const int odd = p.log2n & 1;
    const int digits = p.log2n >> 1;
    const long long base = f * n - p.fill;            // index into `in` of the frame's first sample
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const long long s = base + i;
        const uint32_t w = (s < 0) ? p.partial[(int) (s + p.fill)] : p.in[s];
        const float win = p.window[i];
        // Complex(re / m_scalef, im / m_scalef) * window  (spectrumvis.cpp:98-106)
        const float re = __fmul_rn(__fdiv_rn((float) (short) (w & 0xffffu), p.scalef), win);
        const float im = __fmul_rn(__fdiv_rn((float) ((int) w >> 16), p.scalef), win);
        int pos;
        if (odd) pos = (i & 1) * (n >> 1) + rev4(i >> 1, digits);
        else     pos = rev4(i, digits);
        X[pos] = make_float2(re, im);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// N = 4096 fast path (the SpectrumVis default maximum and BASELINE config 4): 256 threads x 16 points, three register-blocked
// radix-16 passes (4096 = 16*16*16) with two shared-memory exchanges; thread t ends up owning bins t + 256*j.
//   n = 256 n1 + n2, k = k1 + 16 k2:  X[k1 + 16 k2] = sum_{n2} W_4096^{n2 k1} ( sum_{n1} x[256 n1 + n2] W_16^{n1 k1} ) W_256^{n2 k2}
// and the 256-point transforms over n2 = 16 m1 + m2 split the same way (k2 = j1 + 16 j2).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmulc(float2 a, float wr, float wi) { return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr); }

// forward radix-4 butterfly on (a0,a1,a2,a3): y_q = sum_p a_p (-j)^{pq}
__device__ __forceinline__ void bfly4(float2& a0, float2& a1, float2& a2, float2& a3)
{
    const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = make_float2(d02.x + d13.y, d02.y - d13.x);
    a3 = make_float2(d02.x - d13.y, d02.y + d13.x);
}

// 16-point forward DFT in registers, natural order in and out (radix-4 DIF: n = a + 4p, k = q + 4r)
__device__ __forceinline__ void fft16(float2 (&v)[16])
{
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
#pragma unroll
    for (int a = 0; a < 4; ++a) bfly4(v[a], v[a + 4], v[a + 8], v[a + 12]);       // v[a + 4q] = y_q of column a
    // twiddles W_16^{a q}
    v[1 + 4]  = cmulc(v[1 + 4],  C1, -S1);        // a=1 q=1 : W^1
    v[1 + 8]  = cmulc(v[1 + 8],  R2, -R2);        // a=1 q=2 : W^2
    v[1 + 12] = cmulc(v[1 + 12], S1, -C1);        // a=1 q=3 : W^3
    v[2 + 4]  = cmulc(v[2 + 4],  R2, -R2);        // a=2 q=1 : W^2
    v[2 + 8]  = make_float2(v[2 + 8].y, -v[2 + 8].x);          // a=2 q=2 : W^4 = -j
    v[2 + 12] = cmulc(v[2 + 12], -R2, -R2);       // a=2 q=3 : W^6
    v[3 + 4]  = cmulc(v[3 + 4],  S1, -C1);        // a=3 q=1 : W^3
    v[3 + 8]  = cmulc(v[3 + 8],  -R2, -R2);       // a=3 q=2 : W^6
    v[3 + 12] = cmulc(v[3 + 12], -C1, S1);        // a=3 q=3 : W^9
#pragma unroll
    for (int q = 0; q < 4; ++q) bfly4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);   // v[4q + r] = X[q + 4r]
    // reorder to natural order: X[q + 4r] sits in v[4q + r]  -> transpose the 4x4 index
    float2 t;
#define SPEC_SWAP(i, j) t = v[i]; v[i] = v[j]; v[j] = t;
    SPEC_SWAP(1, 4) SPEC_SWAP(2, 8) SPEC_SWAP(3, 12) SPEC_SWAP(6, 9) SPEC_SWAP(7, 13) SPEC_SWAP(11, 14)
#undef SPEC_SWAP
}

// MODE: AVG_* ; LINEAR: linear output ; RCP: scalef is a power of two, so x / scalef == x * (1 / scalef) exactly
template<int MODE, bool LINEAR, bool RCP>
__global__ void __launch_bounds__(SPEC_THREADS, 2) spectrum_kernel_4096(const SpecParams p)
{
    constexpr int K1S = 16 * 17 + 1;                  // k1 stride of the pass-2 output: odd multiple => pass-3 loads hit distinct banks
    extern __shared__ float2 sm[];                    // [16 * K1S] pass-1 output [k1][n2] (4096) aliased with pass-2 output [k1][m2][17]
    float* swin = reinterpret_cast<float*>(sm + 16 * K1S);      // [4096] window
    constexpr int n = 4096, half = 2048;
    const int tid = threadIdx.x;
    const int g = blockIdx.x;
    constexpr bool fixed = (MODE == AVG_FIXED), moving = (MODE == AVG_MOVING);
    double acc[fixed ? 16 : 1];
    long long f0, f1;
    bool emit = true;
    if (fixed) {
        f0 = (long long) g * p.avg_nb - p.fix_idx;
        f1 = f0 + p.avg_nb;
        if (f0 < 0) f0 = 0;
        if (f1 > p.frames) { f1 = p.frames; emit = false; }
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[fixed ? k : 0] = (g == 0 && p.fix_idx > 0) ? p.fix_sum[tid + 256 * k] : 0.0;
    } else { f0 = g; f1 = g + 1; }
#pragma unroll
    for (int k = 0; k < 16; ++k) swin[tid + 256 * k] = p.window[tid + 256 * k];
    const float rs = 1.0f / p.scalef;
    uint32_t raw[16];
    auto fetch = [&](long long f) {
        const long long base = f * n - p.fill + tid;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const long long s = base + 256 * n1;
            raw[n1] = (s < 0) ? p.partial[(int) (s + p.fill)] : p.in[s];
        }
    };
    fetch(f0);
    float2 v[16];
    for (long long f = f0; f < f1; ++f) {
        __syncthreads();                              // swin ready (first frame) / previous frame's pass-3 reads are done
        // pass 1: n2 = tid, transform over n1; Complex(re / m_scalef, im / m_scalef) * window (spectrumvis.cpp:98-106)
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const float win = swin[256 * n1 + tid];
            const float xr = (float) (short) (raw[n1] & 0xffffu), xi = (float) ((int) raw[n1] >> 16);
            v[n1] = RCP ? make_float2(__fmul_rn(__fmul_rn(xr, rs), win), __fmul_rn(__fmul_rn(xi, rs), win))
                        : make_float2(__fmul_rn(__fdiv_rn(xr, p.scalef), win), __fmul_rn(__fdiv_rn(xi, p.scalef), win));
        }
        if (f + 1 < f1) fetch(f + 1);                 // next frame's samples: in flight during the three passes
        fft16(v);
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul(v[k1], __ldg(&p.tw1[k1 * 256 + tid]));       // W_4096^{n2 k1}: the same table values, coalesced
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) sm[k1 * 256 + tid] = v[k1];
        __syncthreads();
        // pass 2: k1 = tid >> 4, m2 = tid & 15, transform over m1 (n2 = 16 m1 + m2)
        {
            const int k1 = tid >> 4, m2 = tid & 15;
#pragma unroll
            for (int m1 = 0; m1 < 16; ++m1) v[m1] = sm[k1 * 256 + 16 * m1 + m2];
            fft16(v);
#pragma unroll
            for (int j1 = 1; j1 < 16; ++j1) v[j1] = cmul(v[j1], __ldg(&p.tw1[16 * 256 + j1 * 16 + m2]));   // W_256^{m2 j1}
            __syncthreads();                                                       // every thread has read its pass-1 values
#pragma unroll
            for (int j1 = 0; j1 < 16; ++j1) sm[k1 * K1S + m2 * 17 + j1] = v[j1];
        }
        __syncthreads();
        // pass 3: k1 = tid & 15, j1 = tid >> 4, transform over m2 -> bins tid + 256 j2
        {
            const int k1 = tid & 15, j1 = tid >> 4;
#pragma unroll
            for (int m2 = 0; m2 < 16; ++m2) v[m2] = sm[k1 * K1S + m2 * 17 + j1];
            fft16(v);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float pw = v[k].x * v[k].x + v[k].y * v[k].y;
            v[k].x = pw;                               // |X|^2 of the last frame stays in v[k].x
            if (fixed) acc[fixed ? k : 0] += (double) pw;
            if (moving) p.power[((long long) (p.avg_nb - 1) + f) * n + tid + 256 * k] = pw;
        }
    }
    if (fixed && !emit) {
        if (g == p.save_sums) {
#pragma unroll
            for (int k = 0; k < 16; ++k) p.fix_sum_out[tid + 256 * k] = acc[fixed ? k : 0];
        }
        return;
    }
    if (moving) return;
    float* out = p.out + (long long) g * n;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int b = tid + 256 * k;
        float res;
        if (fixed) {
            const double avg = acc[fixed ? k : 0] / (double) p.avg_nb;
            res = LINEAR ? v[k].x / p.div : p.mult * log2f((float) avg) + p.ofs;     // spectrumvis.cpp:198,213,222
        } else {
            res = LINEAR ? v[k].x / p.div : p.mult * log2f(v[k].x) + p.ofs;
        }
        if (p.positive_only) { if (b < half) { out[2 * b] = res; out[2 * b + 1] = res; } }
        else out[b < half ? b + half : b - half] = res;
    }
}

typedef void (*spec4096_fn)(const SpecParams);
template<int MODE> spec4096_fn pick4096_m(bool linear, bool rcp)
{
    if (linear) return rcp ? spectrum_kernel_4096<MODE, true, true> : spectrum_kernel_4096<MODE, true, false>;
    return rcp ? spectrum_kernel_4096<MODE, false, true> : spectrum_kernel_4096<MODE, false, false>;
}
spec4096_fn pick4096(int mode, bool linear, bool rcp)
{
    if (mode == AVG_FIXED) return pick4096_m<AVG_FIXED>(linear, rcp);
    if (mode == AVG_MOVING) return pick4096_m<AVG_MOVING>(linear, rcp);
    return pick4096_m<AVG_NONE>(linear, rcp);
}

__global__ void __launch_bounds__(SPEC_THREADS) spectrum_kernel(const SpecParams p)
{
    extern __shared__ float2 spec_X[];
    const int n = p.n, half = n >> 1;
    const int g = blockIdx.x;                         // output frame (or, fixed mode, group) index
    constexpr int BPT = SPEC_MAX_N / SPEC_THREADS;    // bins per thread (max)
    double acc[BPT];
    long long f0, f1;                                 // input frames [f0, f1) feeding this CTA
    bool emit = true;
    if (p.mode == AVG_FIXED && p.avg_nb > 1) {
        // group g covers input frames [g*nb - fix_idx, (g+1)*nb - fix_idx) clipped to this feed
        f0 = (long long) g * p.avg_nb - p.fix_idx;
        f1 = f0 + p.avg_nb;
        if (f0 < 0) f0 = 0;
        if (f1 > p.frames) { f1 = p.frames; emit = false; }
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
            const int b = threadIdx.x + k * SPEC_THREADS;
            acc[k] = (g == 0 && b < n && p.fix_idx > 0) ? p.fix_sum[b] : 0.0;
        }
    } else {
        f0 = g; f1 = g + 1;
    }
    float vlast[BPT];
    for (long long f = f0; f < f1; ++f) {
        load_frame(spec_X, p, f);
        fft_inplace(spec_X, p);
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
            const int b = threadIdx.x + k * SPEC_THREADS;
            if (b < n) {
                const float2 c = spec_X[b];
                const float v = c.x * c.x + c.y * c.y;
                vlast[k] = v;
                if (p.mode == AVG_FIXED && p.avg_nb > 1) acc[k] += (double) v;
                if (p.mode == AVG_MOVING && p.avg_nb > 1) p.power[((long long) (p.avg_nb - 1) + f) * n + b] = v;
            }
        }
        __syncthreads();
    }
    if (p.mode == AVG_FIXED && p.avg_nb > 1 && !emit) {
        if (g == p.save_sums) {
#pragma unroll
            for (int k = 0; k < BPT; ++k) { const int b = threadIdx.x + k * SPEC_THREADS; if (b < n) p.fix_sum_out[b] = acc[k]; }
        }
        return;
    }
    if (p.mode == AVG_MOVING && p.avg_nb > 1) return;      // the sliding sums are formed by spectrum_moving_kernel
    float* out = p.out + (long long) g * n;
#pragma unroll
    for (int k = 0; k < BPT; ++k) {
        const int b = threadIdx.x + k * SPEC_THREADS;
        if (b >= n) continue;
        float res;
        if (p.mode == AVG_FIXED && p.avg_nb > 1) {
            // spectrumvis.cpp:198,213,222: linear mode divides the LAST frame's v, not the average
            const double avg = acc[k] / (double) p.avg_nb;
            res = p.linear ? vlast[k] / p.div : p.mult * log2f((float) avg) + p.ofs;
        } else {
            res = spec_value(p, vlast[k]);
        }
        if (p.positive_only) { if (b < half) { out[2 * b] = res; out[2 * b + 1] = res; } }
        else out[b < half ? b + half : b - half] = res;
    }
}

// moving average: out frame f, bin b = value( float( sum_{k<nb} power[f + k][b] / nb ) ), sums in double
__global__ void spectrum_moving_kernel(const SpecParams p)
{
    const int n = p.n, half = n >> 1;
    const long long f = blockIdx.x;                       // frames on grid.x: a feed may hold more than 65535 of them
    for (int b = blockIdx.y * blockDim.x + threadIdx.x; b < n; b += gridDim.y * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < p.avg_nb; ++k) s += (double) p.power[(f + k) * n + b];
        const float res = spec_value(p, (float) (s / (double) p.avg_nb));
        float* out = p.out + f * n;
        if (p.positive_only) { if (b < half) { out[2 * b] = res; out[2 * b + 1] = res; } }
        else out[b < half ? b + half : b - half] = res;
    }
}

// FFTWindow::create (fftwindow.cpp:20-52, fftwindow.h:52-84): Real (float) arguments, double evaluation, float result
void make_window(int function, int n, std::vector<float>& w)
{
    const double PI = 3.14159265358979323846;
    const float fn = (float) n;
    w.resize(n);
    for (int k = 0; k < n; k++) {
        const float i = (float) k;
        double v;
        switch (function) {
        case 0: v = (2.0 / (fn - 1.0)) * ((fn - 1.0) / 2.0 - fabs(i - (fn - 1.0) / 2.0)) * 2.0; break;
        case 1: v = (0.35875 - 0.48829 * cos((2.0 * PI * i) / fn) + 0.14128 * cos((4.0 * PI * i) / fn) - 0.01168 * cos((6.0 * PI * i) / fn)) * 2.79; break;
        case 2: v = 1.0 - 1.93 * cos((2.0 * PI * i) / fn) + 1.29 * cos((4.0 * PI * i) / fn) - 0.388 * cos((6.0 * PI * i) / fn) + 0.03222 * cos((8.0 * PI * i) / fn); break;
        case 3: v = (0.54 - 0.46 * cos((2.0 * PI * i) / fn)) * 1.855; break;
        case 4: v = (0.5 - 0.5 * cos((2.0 * PI * i) / fn)) * 2.0; break;
        default: v = 1.0; break;
        }
        w[k] = (float) v;
    }
}

} // namespace

struct b200dsp_spectrum {
    int device;
    cudaStream_t stream;
    float scalef;
    int n, log2n, avg_nb, mode, window, linear;
    bool configured;
    float* d_window; float2* d_tw; float2* d_tw1;
    uint32_t* d_partial[2]; int pcur; int fill;         // carried partial frame (ping-pong)
    double* d_fix_sum; int fix_idx; int fix_cur;     // d_fix_sum: two halves of SPEC_MAX_N doubles, fix_cur = the one last written
    float* d_power; long long power_cap;                // moving mode history + frames
    float* d_tmp; long long tmp_cap;
    uint32_t* d_in; long long in_cap; float* d_out; long long out_cap;   // host-path staging
};

namespace {

int plan_counts(b200dsp_spectrum* s, long long n_samples, long long* frames, long long* frames_out)
{
    const long long fr = (s->fill + n_samples) / s->n;
    *frames = fr;
    if (s->mode == AVG_FIXED && s->avg_nb > 1) *frames_out = (s->fix_idx + fr) / s->avg_nb;
    else *frames_out = fr;
    return 0;
}

int spectrum_feed_impl(b200dsp_spectrum* s, const uint32_t* d_in, long long n_samples, int positive_only, float* d_out, long long cap_frames,
                       long long* n_frames, cudaStream_t st)
{
    int rc;
    long long frames, frames_out;
    plan_counts(s, n_samples, &frames, &frames_out);
    if (n_frames) *n_frames = frames_out;
    if (frames_out > cap_frames) return b200_fail(B200DSP_EINVAL, "spectrum_feed: output buffer too small (%lld frames needed)", frames_out);
    const int n = s->n;
    SpecParams p;
    memset(&p, 0, sizeof(p));
    p.in = d_in; p.partial = s->d_partial[s->pcur]; p.window = s->d_window; p.tw = s->d_tw; p.tw1 = s->d_tw1; p.out = d_out;
    p.fix_sum = s->d_fix_sum + (size_t) s->fix_cur * SPEC_MAX_N; p.fix_sum_out = s->d_fix_sum + (size_t) (s->fix_cur ^ 1) * SPEC_MAX_N; p.n = n; p.log2n = s->log2n; p.fill = s->fill; p.frames = (int) frames;
    p.mode = s->mode; p.avg_nb = s->avg_nb; p.linear = s->linear; p.positive_only = positive_only; p.fix_idx = s->fix_idx;
    p.frames_out = (int) frames_out; p.save_sums = -1;
    p.scalef = s->scalef; p.ofs = 20.0f * log10f(1.0f / (float) n); p.div = (float) (n * n); p.mult = 10.0f / log2f(10.0f);
    const bool fixed = (s->mode == AVG_FIXED && s->avg_nb > 1), moving = (s->mode == AVG_MOVING && s->avg_nb > 1);
    if (frames > 0) {
        long long ctas = frames;
        if (fixed) {
            ctas = (s->fix_idx + frames + s->avg_nb - 1) / s->avg_nb;          // groups touched, the last may be incomplete
            if ((s->fix_idx + frames) % s->avg_nb) { p.save_sums = (int) (ctas - 1); s->fix_cur ^= 1; }
        }
        if (moving) {
            const long long need = (frames + s->avg_nb - 1) * n;
            if (s->power_cap < need) {
                float* np = nullptr;
                if ((rc = B200_CUDA_CHECK(cudaMalloc(&np, (size_t) need * 4))) || (rc = B200_CUDA_CHECK(cudaMemsetAsync(np, 0, (size_t) need * 4, st)))) return rc;
                if (s->d_power) {
                    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(np, s->d_power, (size_t) (s->avg_nb - 1) * n * 4, cudaMemcpyDeviceToDevice, st))) ||
                        (rc = B200_CUDA_CHECK(cudaStreamSynchronize(st)))) return rc;
                    cudaFree(s->d_power);
                }
                s->d_power = np; s->power_cap = need;
            }
            p.power = s->d_power;
        }
        const int threads = SPEC_THREADS;
        int ex = 0;
        const bool rcp = (frexpf(s->scalef, &ex) == 0.5f);                     // power of two: the reciprocal multiply is exact
        const int eff_mode = fixed ? AVG_FIXED : (moving ? AVG_MOVING : AVG_NONE);
        if (n == 4096) {
            const size_t smem4096 = (size_t) 16 * (16 * 17 + 1) * sizeof(float2) + 4096 * sizeof(float);
            spec4096_fn fn = pick4096(eff_mode, s->linear != 0, rcp);
            if ((rc = B200_CUDA_CHECK(cudaFuncSetAttribute((const void*) fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem4096)))) return rc;
            fn<<<(unsigned) ctas, threads, smem4096, st>>>(p);
        }
        else           spectrum_kernel<<<(unsigned) ctas, threads, (size_t) n * sizeof(float2), st>>>(p);
        if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
        if (moving) {
            spectrum_moving_kernel<<<dim3((unsigned) frames, (n + 255) / 256), 256, 0, st>>>(p);
            if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
            // carry the last avg_nb-1 power frames to the front for the next feed (through a bounce buffer: the ranges may overlap)
            const size_t hb = (size_t) (s->avg_nb - 1) * n * 4;
            if (s->tmp_cap < (long long) hb) {
                if (s->d_tmp) cudaFree(s->d_tmp);
                s->d_tmp = nullptr; s->tmp_cap = 0;
                if ((rc = B200_CUDA_CHECK(cudaMalloc(&s->d_tmp, hb)))) return rc;
                s->tmp_cap = (long long) hb;
            }
            if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(s->d_tmp, s->d_power + frames * n, hb, cudaMemcpyDeviceToDevice, st))) ||
                (rc = B200_CUDA_CHECK(cudaMemcpyAsync(s->d_power, s->d_tmp, hb, cudaMemcpyDeviceToDevice, st)))) return rc;
        }
        if (fixed) s->fix_idx = (int) ((s->fix_idx + frames) % s->avg_nb);
    }
    // carry the trailing partial frame: new_partial = last (fill + n_samples - frames*n) samples of [partial ++ in]
    const long long new_fill = s->fill + n_samples - frames * n;
    if (new_fill > 0) {
        uint32_t* np = s->d_partial[s->pcur ^ 1];
        const long long from_in = (new_fill < n_samples) ? new_fill : n_samples;       // samples taken from the new input
        const long long from_old = new_fill - from_in;                                  // only when no frame completed
        if (from_old > 0 && (rc = B200_CUDA_CHECK(cudaMemcpyAsync(np, s->d_partial[s->pcur] + (s->fill - from_old), (size_t) from_old * 4, cudaMemcpyDeviceToDevice, st)))) return rc;
        if (from_in > 0 && (rc = B200_CUDA_CHECK(cudaMemcpyAsync(np + from_old, d_in + (n_samples - from_in), (size_t) from_in * 4, cudaMemcpyDeviceToDevice, st)))) return rc;
        s->pcur ^= 1;
    }
    s->fill = (int) new_fill;
    return 0;
}

} // namespace

extern "C" {

int b200dsp_spectrum_create(b200dsp_spectrum_t** out, float scalef)
{
    if (!out) return b200_fail(B200DSP_EINVAL, "spectrum_create: null handle pointer");
    *out = nullptr;
    int rc = b200_require_device();
    if (rc) return rc;
    b200dsp_spectrum* s = new (std::nothrow) b200dsp_spectrum();
    if (!s) return b200_fail(B200DSP_ENOMEM, "spectrum_create: out of host memory");
    memset(s, 0, sizeof(*s));
    s->device = b200_current_device();
    s->scalef = scalef;
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(s->device))) || (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)))) { delete s; return rc; }
    std::vector<float2> tw(SPEC_MAX_N);
    const double PI = 3.14159265358979323846;
    for (int t = 0; t < SPEC_MAX_N; ++t) tw[t] = make_float2((float) cos(-2.0 * PI * t / SPEC_MAX_N), (float) sin(-2.0 * PI * t / SPEC_MAX_N));
    std::vector<float2> tw1(16 * 256 + 16 * 16);
    for (int k1 = 0; k1 < 16; ++k1) for (int t = 0; t < 256; ++t) tw1[(size_t) (k1 * 256 + t)] = tw[(size_t) ((t * k1) & 4095)];
    for (int j1 = 0; j1 < 16; ++j1) for (int m2 = 0; m2 < 16; ++m2) tw1[(size_t) (16 * 256 + j1 * 16 + m2)] = tw[(size_t) ((16 * m2 * j1) & 4095)];
    if ((rc = B200_CUDA_CHECK(cudaMalloc(&s->d_tw1, tw1.size() * sizeof(float2)))) ||
        (rc = B200_CUDA_CHECK(cudaMemcpy(s->d_tw1, tw1.data(), tw1.size() * sizeof(float2), cudaMemcpyHostToDevice))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&s->d_tw, SPEC_MAX_N * sizeof(float2)))) ||
        (rc = B200_CUDA_CHECK(cudaMemcpy(s->d_tw, tw.data(), SPEC_MAX_N * sizeof(float2), cudaMemcpyHostToDevice))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&s->d_window, SPEC_MAX_N * sizeof(float)))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&s->d_partial[0], SPEC_MAX_N * 4))) || (rc = B200_CUDA_CHECK(cudaMalloc(&s->d_partial[1], SPEC_MAX_N * 4))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&s->d_fix_sum, 2 * SPEC_MAX_N * sizeof(double))))) { b200dsp_spectrum_destroy(s); return rc; }
    *out = s;
    // same defaults as the reference constructor (spectrumvis.cpp:20-34): 1024 points, Blackman-Harris, no averaging
    return b200dsp_spectrum_configure(s, 1024, 0, 0, AVG_NONE, 1, 0);
}

int b200dsp_spectrum_destroy(b200dsp_spectrum_t* s)
{
    if (!s) return 0;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->d_tw) cudaFree(s->d_tw);
    if (s->d_tw1) cudaFree(s->d_tw1);
    if (s->d_window) cudaFree(s->d_window);
    if (s->d_partial[0]) cudaFree(s->d_partial[0]);
    if (s->d_partial[1]) cudaFree(s->d_partial[1]);
    if (s->d_fix_sum) cudaFree(s->d_fix_sum);
    if (s->d_power) cudaFree(s->d_power);
    if (s->d_tmp) cudaFree(s->d_tmp);
    if (s->d_in) cudaFree(s->d_in);
    if (s->d_out) cudaFree(s->d_out);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return 0;
}

int b200dsp_spectrum_configure(b200dsp_spectrum_t* s, int fft_size, int overlap_percent, unsigned int average_nb, int averaging_mode, int window, int linear)
{
    if (!s) return b200_fail(B200DSP_EINVAL, "null handle");
    if (averaging_mode < 0 || averaging_mode > 2 || window < 0 || window > 5) return b200_fail(B200DSP_EINVAL, "spectrum_configure: bad mode/window");
    if (overlap_percent != 0) return b200_fail(B200DSP_EINVAL, "spectrum_configure: only overlap 0 is supported (the reference makes no progress otherwise, spectrumvis.cpp:91,236-239)");
    if (fft_size > SPEC_MAX_N) fft_size = SPEC_MAX_N; else if (fft_size < 64) fft_size = 64;       // spectrumvis.cpp:292-299
    int l2 = 0;
    while ((1 << l2) < fft_size) ++l2;
    if ((1 << l2) != fft_size) return b200_fail(B200DSP_EINVAL, "spectrum_configure: fft size must be a power of two");
    int rc = B200_CUDA_CHECK(cudaSetDevice(s->device));
    if (rc) return rc;
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(s->stream)))) return rc;
    s->n = fft_size; s->log2n = l2; s->avg_nb = (int) average_nb; s->mode = averaging_mode; s->window = window; s->linear = linear ? 1 : 0;
    std::vector<float> w;
    make_window(window, fft_size, w);
    if ((rc = B200_CUDA_CHECK(cudaMemcpy(s->d_window, w.data(), (size_t) fft_size * 4, cudaMemcpyHostToDevice)))) return rc;
    // handleConfigure restarts the frame buffer and both averagers (spectrumvis.cpp:317-321)
    s->fill = 0; s->fix_idx = 0; s->fix_cur = 0; s->pcur = 0;
    if ((rc = B200_CUDA_CHECK(cudaMemset(s->d_fix_sum, 0, 2 * SPEC_MAX_N * sizeof(double))))) return rc;
    if (s->d_power) { cudaFree(s->d_power); s->d_power = nullptr; s->power_cap = 0; }
    if ((rc = B200_CUDA_CHECK(cudaDeviceSynchronize()))) return rc;      // the copies above ran on the default stream
    s->configured = true;
    return 0;
}

int64_t b200dsp_spectrum_frames_for(b200dsp_spectrum_t* s, int64_t n_samples)
{
    if (!s || n_samples < 0) return -1;
    long long fr, fo;
    plan_counts(s, n_samples, &fr, &fo);
    return fo;
}

int b200dsp_spectrum_feed_dev(b200dsp_spectrum_t* s, const void* d_iq, int64_t n_samples, int positive_only, float* d_out_frames, int64_t cap_frames,
                              int64_t* n_frames, void* cuda_stream)
{
    if (!s) return b200_fail(B200DSP_EINVAL, "null handle");
    if (n_samples < 0 || (n_samples > 0 && !d_iq)) return b200_fail(B200DSP_EINVAL, "spectrum_feed: bad buffer");
    int rc = B200_CUDA_CHECK(cudaSetDevice(s->device));
    if (rc) return rc;
    long long nf = 0;
    rc = spectrum_feed_impl(s, (const uint32_t*) d_iq, n_samples, positive_only, d_out_frames, cap_frames, &nf, cuda_stream ? (cudaStream_t) cuda_stream : s->stream);
    if (n_frames) *n_frames = nf;
    return rc;
}

int b200dsp_spectrum_feed(b200dsp_spectrum_t* s, const int16_t* iq, int64_t n_samples, int positive_only, float* out_frames, int64_t cap_frames, int64_t* n_frames)
{
    if (!s) return b200_fail(B200DSP_EINVAL, "null handle");
    if (n_samples < 0 || (n_samples > 0 && !iq)) return b200_fail(B200DSP_EINVAL, "spectrum_feed: bad buffer");
    int rc = B200_CUDA_CHECK(cudaSetDevice(s->device));
    if (rc) return rc;
    long long fr, fo;
    plan_counts(s, n_samples, &fr, &fo);
    if (n_frames) *n_frames = fo;
    if (fo > cap_frames) return b200_fail(B200DSP_EINVAL, "spectrum_feed: output buffer too small (%lld frames needed)", fo);
    if (n_samples == 0) return 0;
    if (s->in_cap < n_samples) {
        if (s->d_in) cudaFree(s->d_in);
        s->d_in = nullptr; s->in_cap = 0;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&s->d_in, (size_t) n_samples * 4)))) return rc;
        s->in_cap = n_samples;
    }
    const long long out_elems = fo * s->n;
    if (s->out_cap < out_elems) {
        if (s->d_out) cudaFree(s->d_out);
        s->d_out = nullptr; s->out_cap = 0;
        if (out_elems && (rc = B200_CUDA_CHECK(cudaMalloc(&s->d_out, (size_t) out_elems * 4)))) return rc;
        s->out_cap = out_elems;
    }
    long long nf = 0;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(s->d_in, iq, (size_t) n_samples * 4, cudaMemcpyHostToDevice, s->stream))) ||
        (rc = spectrum_feed_impl(s, s->d_in, n_samples, positive_only, s->d_out, fo, &nf, s->stream))) return rc;
    if (out_elems && (rc = B200_CUDA_CHECK(cudaMemcpyAsync(out_frames, s->d_out, (size_t) out_elems * 4, cudaMemcpyDeviceToHost, s->stream)))) return rc;
    return B200_CUDA_CHECK(cudaStreamSynchronize(s->stream));
}

} // extern "C"
