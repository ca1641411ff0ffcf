// fftfilt_api.cu — SURVEY.md 8f-4, SSB / DSB / AM-synchronous back-end: b200dsp_fftfilt_* (K11)
//
// Replaces (paths relative to the reference tree):
//   fftfilt::fftfilt / create_filter / create_dsb_filter      sdrbase/dsp/fftfilt.cpp:49-166  (fsinc, _blackman: fftfilt.h:52-64)
//   fftfilt::runFilt / runSSB / runDSB                         sdrbase/dsp/fftfilt.cpp:261-360 (Fldigi's overlap-add fast convolution)
//   g_fft<float>::ComplexFFT / InverseComplexFFT               sdrbase/dsp/gfft.h (forward unnormalised, inverse scaled by 1/N)
//   callers: SSBDemod::feed (plugins/channelrx/demodssb/ssbdemod.cpp:91-92,165-175), AMDemod (amdemod.cpp:199-207)
// The reference collects flen/2 samples, zero-pads to flen, transforms, multiplies by the filter's frequency response
// (runSSB: one half-band zeroed, bin 0 kept or rejected, bin flen/2 left untouched by its loops), transforms back and
// overlap-adds: out = previous block's second half + this block's first half.  Blocks are independent up to that overlap,
// so a CTA walks a contiguous range of blocks with the overlap in shared memory; a range that starts inside the call first
// recomputes the block before it.  The transforms are radix-2 in shared memory: decimation in frequency forward (natural in,
// bit-reversed out), the multiplier table stored bit-reversed, decimation in time back (bit-reversed in, natural out), so
// no reordering pass exists.  A different FFT factorisation than g_fft: agreement is to float32 rounding (<= 1e-5 of the
// block maximum in the tests; the reference's own -ffast-math and strict builds differ by ~2e-7).
#include "common.cuh"
#include <math.h>
#include <complex>
#include <vector>

using namespace b200dsp;

namespace {

constexpr int FF_THREADS = 256;

struct FfParams {
    const float2* in;          // new input samples (device)
    const float2* pend;        // samples carried from the previous call (inptr of them)
    const float2* ovl_in;      // [flen2]
    float2* ovl_out;
    const float2* mult;        // [flen] frequency multiplier, bit-reversed order
    const float2* tw;          // [flen / 2] exp(-2 pi i k / flen)
    float2* out;
    int inptr;                 // pending samples
    int nb;                    // blocks this call
    int flen, log2n;
    int blocks_per_cta;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__global__ void __launch_bounds__(FF_THREADS) fftfilt_kernel(const FfParams p)
{
    extern __shared__ float2 ff_smem[];
    const int N = p.flen, N2 = N >> 1, tid = threadIdx.x;
    float2* x = ff_smem;               // [N]
    float2* ovl = ff_smem + N;         // [N2]
    float2* tw = ff_smem + N + N2;     // [N2]
    const int b0 = blockIdx.x * p.blocks_per_cta;
    int b1 = b0 + p.blocks_per_cta;
    if (b1 > p.nb) b1 = p.nb;
    if (b0 >= b1) return;
    for (int k = tid; k < N2; k += FF_THREADS) { tw[k] = p.tw[k]; ovl[k] = (b0 == 0) ? p.ovl_in[k] : make_float2(0.0f, 0.0f); }
    const float inv = 1.0f / (float) N;
    for (int b = (b0 == 0 ? 0 : b0 - 1); b < b1; ++b) {
        const bool emit = (b >= b0);
        __syncthreads();
        // block b = samples [b N2, (b + 1) N2) of (pending | new input), zero-padded to N
        for (int k = tid; k < N; k += FF_THREADS) {
            float2 v = make_float2(0.0f, 0.0f);
            if (k < N2) {
                const long long g = (long long) b * N2 + k;
                v = (g < p.inptr) ? p.pend[g] : p.in[g - p.inptr];
            }
            x[k] = v;
        }
        __syncthreads();
        // forward, decimation in frequency
        for (int s = p.log2n - 1; s >= 0; --s) {
            const int half = 1 << s;
            for (int k = tid; k < N2; k += FF_THREADS) {
                const int r = k & (half - 1), i = ((k >> s) << (s + 1)) + r, j = i + half;
                const float2 a = x[i], c = x[j];
                x[i] = make_float2(a.x + c.x, a.y + c.y);
                x[j] = cmul(make_float2(a.x - c.x, a.y - c.y), tw[r << (p.log2n - 1 - s)]);
            }
            __syncthreads();
        }
        for (int k = tid; k < N; k += FF_THREADS) x[k] = cmul(x[k], p.mult[k]);
        __syncthreads();
        // inverse, decimation in time (conjugate twiddles)
        for (int s = 0; s < p.log2n; ++s) {
            const int half = 1 << s;
            for (int k = tid; k < N2; k += FF_THREADS) {
                const int r = k & (half - 1), i = ((k >> s) << (s + 1)) + r, j = i + half;
                const float2 w = tw[r << (p.log2n - 1 - s)];
                const float2 t = cmul(x[j], make_float2(w.x, -w.y)), a = x[i];
                x[i] = make_float2(a.x + t.x, a.y + t.y);
                x[j] = make_float2(a.x - t.x, a.y - t.y);
            }
            __syncthreads();
        }
        // overlap and add (fftfilt.cpp:274-277)
        for (int k = tid; k < N2; k += FF_THREADS) {
            const float2 lo = x[k], hi = x[k + N2], o = ovl[k];
            if (emit) p.out[(long long) b * N2 + k] = make_float2(o.x + lo.x * inv, o.y + lo.y * inv);
            ovl[k] = make_float2(hi.x * inv, hi.y * inv);
        }
    }
    __syncthreads();
    if (b1 == p.nb) for (int k = tid; k < N2; k += FF_THREADS) p.ovl_out[k] = ovl[k];
}

// ---- flen 1024 / 2048 (the reference's SSB and DSB filter lengths: ssbdemod.h:36, ssbdemod.cpp:91-92, amdemod.cpp:72-73) ----
// The same radix-2 butterflies as above, in the same order, but FOUR stages at a time on 16 values in registers: a thread
// owns the 16 elements that differ in index bits [P, P+4) ("field P"), so a 1024-point transform is 3 passes through
// shared memory instead of 10 and the CTA is flen/16 threads.  Forward fields: top, middle ..., 0; inverse: 0, middle ..., top.
//   * the block's samples go from global memory straight into the top-field pass (its upper half is the zero padding);
//   * the last forward pass, the multiplier and the first inverse pass all work on field 0: one visit;
//   * the last inverse pass (top field) holds k and k + flen/2 in one thread: scaling, overlap-add and the store follow in
//     registers, and the overlap itself never leaves the thread's registers between blocks.
// Shared arrays are padded by one element per 16 (every pass then reads and writes without bank conflicts); the twiddles are
// stored per stage (tws[2^s - 1 + r] = exp(-2 pi i r / 2^(s+1))), so a stage's loads are contiguous or a broadcast.
__device__ __forceinline__ int ff_idx(int e) { return e + (e >> 4); }

// exp(-2 pi i q / 16), q = 0 .. 7 (the float values of the double cosines, like the table's)
__device__ __forceinline__ constexpr float ff_w16_re(int q)
{
    return q == 0 ? 1.0f : q == 1 ? 0.92387953251128674f : q == 2 ? 0.70710678118654752f : q == 3 ? 0.38268343236508977f :
           q == 4 ? 0.0f : q == 5 ? -0.38268343236508977f : q == 6 ? -0.70710678118654752f : -0.92387953251128674f;
}
__device__ __forceinline__ constexpr float ff_w16_im(int q)
{
    return q == 0 ? 0.0f : q == 1 ? -0.38268343236508977f : q == 2 ? -0.70710678118654752f : q == 3 ? -0.92387953251128674f :
           q == 4 ? -1.0f : q == 5 ? -0.92387953251128674f : q == 6 ? -0.70710678118654752f : -0.38268343236508977f;
}

// stages S_HI .. S_LO (forward, decimation in frequency) or S_LO .. S_HI (inverse, decimation in time) of field P on v[16];
// low = the element index bits below P
template<bool INV, int P, int S_LO, int S_HI>
__device__ __forceinline__ void ff_stages(float2 (&v)[16], const float2* __restrict__ tws, int low)
{
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
        const int b = INV ? bb : 3 - bb;
        const int s = P + b;
        if (s < S_LO || s > S_HI) continue;
        const float2* t = tws + ((1 << s) - 1);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            if (m & (1 << b)) continue;
            const float2 a = v[m], c = v[m | (1 << b)];
            // field 0: the twiddle exp(-2 pi i q / 16), q = r << (3 - s), is known at compile time; q = 0 and q = 4 (1 and -j) cost nothing
            const int q16 = (P == 0) ? ((m & ((1 << b) - 1)) << (3 - s)) : -1;
            float2 w;
            if (P == 0) w = make_float2(ff_w16_re(q16), ff_w16_im(q16));
            else        w = t[low + ((m & ((1 << b) - 1)) << P)];
            if (!INV) {
                const float2 d = make_float2(a.x - c.x, a.y - c.y);
                v[m] = make_float2(a.x + c.x, a.y + c.y);
                v[m | (1 << b)] = (q16 == 0) ? d : (q16 == 4) ? make_float2(d.y, -d.x) : cmul(d, w);
            } else {
                const float2 u = (q16 == 0) ? c : (q16 == 4) ? make_float2(-c.y, c.x) : cmul(c, make_float2(w.x, -w.y));
                v[m] = make_float2(a.x + u.x, a.y + u.y);
                v[m | (1 << b)] = make_float2(a.x - u.x, a.y - u.y);
            }
        }
    }
}

template<int P>
__device__ __forceinline__ void ff_field_load(const float2* x, int G, float2 (&v)[16])
{
    const int e0 = ((G >> P) << (P + 4)) | (G & ((1 << P) - 1));
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = x[ff_idx(e0 + (m << P))];
}
template<int P>
__device__ __forceinline__ void ff_field_store(float2* x, int G, const float2 (&v)[16])
{
    const int e0 = ((G >> P) << (P + 4)) | (G & ((1 << P) - 1));
#pragma unroll
    for (int m = 0; m < 16; ++m) x[ff_idx(e0 + (m << P))] = v[m];
}

// middle passes (fields strictly between 0 and the top one), forward from the top down / inverse from the bottom up
template<int LOG2N, int P>
__device__ __forceinline__ void ff_mid_forward(float2* x, const float2* tws, int G)
{
    if constexpr (P > 0) {
        float2 v[16];
        ff_field_load<P>(x, G, v);
        ff_stages<false, P, P, P + 3>(v, tws, G & ((1 << P) - 1));
        ff_field_store<P>(x, G, v);
        __syncthreads();
        ff_mid_forward<LOG2N, P - 4>(x, tws, G);
    }
}
template<int LOG2N, int S>
__device__ __forceinline__ void ff_mid_inverse(float2* x, const float2* tws, int G)
{
    if constexpr (LOG2N - S > 4) {
        float2 v[16];
        ff_field_load<S>(x, G, v);
        ff_stages<true, S, S, S + 3>(v, tws, G & ((1 << S) - 1));
        ff_field_store<S>(x, G, v);
        __syncthreads();
        ff_mid_inverse<LOG2N, S + 4>(x, tws, G);
    }
}
template<int LOG2N> constexpr int ff_last_inverse_stage() { int s = 4; while (LOG2N - s > 4) s += 4; return s; }

#ifndef FF_FAST_WARPS_PER_SM
#define FF_FAST_WARPS_PER_SM 16       // resident warps the register budget is sized for (16: 128 registers, no spills)
#endif
template<int LOG2N>
__global__ void __launch_bounds__((1 << LOG2N) / 16, 32 * FF_FAST_WARPS_PER_SM / ((1 << LOG2N) / 16)) fftfilt_fast_kernel(const FfParams p)
{
    constexpr int N = 1 << LOG2N, N2 = N / 2, NT = N / 16, TOP = LOG2N - 4;
    constexpr int F0_HI = (LOG2N % 4 == 0) ? 3 : (LOG2N % 4) - 1;      // field 0 ends the forward transform with stages F0_HI .. 0
    constexpr int SL = ff_last_inverse_stage<LOG2N>();                 // the top field ends the inverse transform with stages SL .. LOG2N-1
    static_assert(LOG2N >= 8 && LOG2N <= 12, "field layout");
    extern __shared__ float2 ff_smem[];
    float2* x = ff_smem;                      // [N + N / 16]
    float2* tws = ff_smem + N + N / 16;       // [N]: per-stage twiddles
    const int G = threadIdx.x;
    const int b0 = blockIdx.x * p.blocks_per_cta;
    int b1 = b0 + p.blocks_per_cta;
    if (b1 > p.nb) b1 = p.nb;
    if (b0 >= b1) return;
    for (int s = 0; s < LOG2N; ++s)
        for (int r = G; r < (1 << s); r += NT) tws[(1 << s) - 1 + r] = p.tw[r << (LOG2N - 1 - s)];
    __syncthreads();
    float2 ov[8];                             // overlap of output samples k = m NT + G, m < 8
#pragma unroll
    for (int m = 0; m < 8; ++m) ov[m] = (b0 == 0) ? p.ovl_in[m * NT + G] : make_float2(0.0f, 0.0f);
    const float inv = 1.0f / (float) N;
    for (int b = (b0 == 0 ? 0 : b0 - 1); b < b1; ++b) {
        const bool emit = (b >= b0);
        float2 v[16];
        // block b = samples [b N2, (b + 1) N2) of (pending | new input); elements m >= 8 of the top field are the zero padding
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const long long g = (long long) b * N2 + m * NT + G;
            v[m] = (g < p.inptr) ? p.pend[g] : p.in[g - p.inptr];
        }
        {
            const float2* t = tws + ((1 << (LOG2N - 1)) - 1);          // stage LOG2N-1 on (a, 0): a stays, the partner is a w
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m + 8] = cmul(v[m], t[G + (m << TOP)]);
        }
        ff_stages<false, TOP, TOP, LOG2N - 2>(v, tws, G);
        __syncthreads();                      // (the previous block's last pass has read x)
        ff_field_store<TOP>(x, G, v);
        __syncthreads();
        ff_mid_forward<LOG2N, TOP - 4>(x, tws, G);
        // field 0: end of the forward transform, multiplier (bit-reversed table), start of the inverse transform
        ff_field_load<0>(x, G, v);
        ff_stages<false, 0, 0, F0_HI>(v, tws, 0);
        {
            const float4* mt = reinterpret_cast<const float4*>(p.mult + 16 * G);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 mm = __ldg(mt + q);
                v[2 * q] = cmul(v[2 * q], make_float2(mm.x, mm.y));
                v[2 * q + 1] = cmul(v[2 * q + 1], make_float2(mm.z, mm.w));
            }
        }
        ff_stages<true, 0, 0, 3>(v, tws, 0);
        ff_field_store<0>(x, G, v);
        __syncthreads();
        ff_mid_inverse<LOG2N, 4>(x, tws, G);
        ff_field_load<TOP>(x, G, v);
        ff_stages<true, TOP, SL, LOG2N - 1>(v, tws, G);
        // overlap and add (fftfilt.cpp:274-277)
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const float2 lo = v[m], hi = v[m + 8], o = ov[m];
            if (emit) p.out[(long long) b * N2 + m * NT + G] = make_float2(o.x + lo.x * inv, o.y + lo.y * inv);
            ov[m] = make_float2(hi.x * inv, hi.y * inv);
        }
    }
    if (b1 == p.nb) {
#pragma unroll
        for (int m = 0; m < 8; ++m) p.ovl_out[m * NT + G] = ov[m];
    }
}

// iterative radix-2 FFT in double (host): the filter's frequency response
void host_fft(std::vector<std::complex<double>>& a)
{
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = -2.0 * 3.14159265358979323846 / (double) len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const std::complex<double> w(cos(ang * (double) k), sin(ang * (double) k));
                const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v; a[i + k + len / 2] = u - v;
            }
    }
}

// fftfilt::fsinc / _blackman (fftfilt.h:52-64): float results of double expressions
float ff_fsinc(float fc, int i, int len)
{
    const int len2 = len / 2;
    return (i == len2) ? (float) (2.0 * fc) : (float) (sin(2 * 3.14159265358979323846 * fc * (i - len2)) / (3.14159265358979323846 * (i - len2)));
}
float ff_blackman(int i, int len)
{
    return (float) (0.42 - 0.50 * cos(2.0 * 3.14159265358979323846 * i / len) + 0.08 * cos(4.0 * 3.14159265358979323846 * i / len));
}

} // namespace

struct b200dsp_fftfilt {
    int device = 0; cudaStream_t stream = nullptr;
    int flen = 0, log2n = 0;
    std::vector<std::complex<float>> filter;      // frequency response, natural order
    float2* d_mult = nullptr; int mult_key = -1;   // multiplier of the last (op, usb, get_dc), bit-reversed
    float2* d_tw = nullptr;
    float2* d_ovl[2] = { nullptr, nullptr }; int cur = 0;
    float2* d_pend[2] = { nullptr, nullptr }; int pcur = 0; int inptr = 0;
    float2* d_in = nullptr; float2* d_out = nullptr; long long cap = 0;
    int sm_count = 148;
};

namespace {

void ff_make_filter(b200dsp_fftfilt* h, int kind, float f1, float f2)
{
    const int flen = h->flen, flen2 = flen / 2;
    std::vector<std::complex<float>> t((size_t) flen, std::complex<float>(0.0f, 0.0f));
    if (kind == 0) {               // create_filter (fftfilt.cpp:107-145)
        const bool lowpass = (f2 != 0), highpass = (f1 != 0);
        for (int i = 0; i < flen2; i++) {
            float v = 0;
            if (lowpass) v += ff_fsinc(f2, i, flen2);
            if (highpass) v -= ff_fsinc(f1, i, flen2);
            t[(size_t) i] = v;
        }
        if (highpass && f2 < f1) t[(size_t) (flen2 / 2)] += 1.0f;
        for (int i = 0; i < flen2; i++) t[(size_t) i] *= ff_blackman(i, flen2);
    } else {                       // create_dsb_filter (fftfilt.cpp:148-166)
        for (int i = 0; i < flen2; i++) { t[(size_t) i] = ff_fsinc(f2, i, flen2); t[(size_t) i] *= ff_blackman(i, flen2); }
    }
    std::vector<std::complex<double>> a((size_t) flen);
    for (int i = 0; i < flen; i++) a[(size_t) i] = std::complex<double>(t[(size_t) i].real(), t[(size_t) i].imag());
    host_fft(a);
    h->filter.resize((size_t) flen);
    for (int i = 0; i < flen; i++) h->filter[(size_t) i] = std::complex<float>((float) a[(size_t) i].real(), (float) a[(size_t) i].imag());
    float scale = 0;               // "normalize the output filter for unity gain": the largest magnitude of the first half
    for (int i = 0; i < flen2; i++) { const float mag = std::abs(h->filter[(size_t) i]); if (mag > scale) scale = mag; }
    if (scale != 0) for (int i = 0; i < flen; i++) h->filter[(size_t) i] /= scale;
    h->mult_key = -1;
}

int ff_upload_mult(b200dsp_fftfilt* h, int op, int usb, int get_dc, cudaStream_t user)
{
    const int key = op * 4 + (usb ? 1 : 0) + (get_dc ? 2 : 0);
    if (key == h->mult_key) return 0;
    { const int rc0 = B200_CUDA_CHECK(cudaStreamSynchronize(user)); if (rc0) return rc0; }      // an earlier run may still be reading the table
    const int N = h->flen, N2 = N / 2;
    std::vector<std::complex<float>> m(h->filter);
    if (op == 1) {                 // runSSB (fftfilt.cpp:285-325): its loops run i = 1 .. flen2-1, bin flen2 is left as it is
        m[0] = get_dc ? h->filter[0] : std::complex<float>(0.0f, 0.0f);
        for (int i = 1; i < N2; i++) {
            if (usb) m[(size_t) (N2 + i)] = 0.0f; else m[(size_t) i] = 0.0f;
        }
        m[(size_t) N2] = 1.0f;
    } else if (op == 2) {          // runDSB (fftfilt.cpp:328-357)
        if (!get_dc) m[0] = 0.0f;
    }
    std::vector<float2> br((size_t) N);
    for (int k = 0; k < N; k++) {
        int r = 0;
        for (int b = 0; b < h->log2n; b++) if (k & (1 << b)) r |= 1 << (h->log2n - 1 - b);
        br[(size_t) k] = make_float2(m[(size_t) r].real(), m[(size_t) r].imag());       // position k of the DIF output holds frequency bitrev(k)
    }
    int rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_mult, br.data(), (size_t) N * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    if (rc) return rc;
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(h->stream)))) return rc;       // br is a local
    h->mult_key = key;
    return 0;
}

} // namespace

extern "C" {

int b200dsp_fftfilt_destroy(b200dsp_fftfilt_t* h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    float2* ptrs[] = { h->d_mult, h->d_tw, h->d_ovl[0], h->d_ovl[1], h->d_pend[0], h->d_pend[1], h->d_in, h->d_out };
    for (float2* q : ptrs) if (q) cudaFree(q);
    delete h;
    return 0;
}

int b200dsp_fftfilt_create(b200dsp_fftfilt_t** out, int kind, float f1, float f2, int len)
{
    if (!out) return b200_fail(B200DSP_EINVAL, "fftfilt_create: null handle pointer");
    *out = nullptr;
    if (kind < 0 || kind > 1 || len < 16 || len > 4096 || (len & (len - 1))) return b200_fail(B200DSP_EINVAL, "fftfilt_create: kind 0/1, len a power of two in 16..4096");
    int rc = b200_require_device();
    if (rc) return rc;
    b200dsp_fftfilt* h = new (std::nothrow) b200dsp_fftfilt();
    if (!h) return b200_fail(B200DSP_ENOMEM, "fftfilt_create: out of host memory");
    h->device = b200_current_device(); h->flen = len; h->sm_count = b200_sm_count_of(h->device);
    while ((1 << h->log2n) < len) ++h->log2n;
    const size_t N = (size_t) len, N2 = N / 2;
    std::vector<float2> tw(N2);
    for (size_t k = 0; k < N2; k++) { const double a = -2.0 * 3.14159265358979323846 * (double) k / (double) N; tw[k] = make_float2((float) cos(a), (float) sin(a)); }
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(h->device))) || (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_mult, N * 8))) || (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_tw, N2 * 8))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_ovl[0], N2 * 8))) || (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_ovl[1], N2 * 8))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_pend[0], N2 * 8))) || (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_pend[1], N2 * 8))) ||
        (rc = B200_CUDA_CHECK(cudaMemset(h->d_ovl[0], 0, N2 * 8))) || (rc = B200_CUDA_CHECK(cudaMemset(h->d_ovl[1], 0, N2 * 8))) ||
        (rc = B200_CUDA_CHECK(cudaMemcpy(h->d_tw, tw.data(), N2 * 8, cudaMemcpyHostToDevice)))) { b200dsp_fftfilt_destroy(h); return rc; }
    ff_make_filter(h, kind, f1, f2);
    *out = h;
    return 0;
}

int b200dsp_fftfilt_set_filter(b200dsp_fftfilt_t* h, int kind, float f1, float f2)
{
    if (!h || kind < 0 || kind > 1) return b200_fail(B200DSP_EINVAL, "fftfilt_set_filter: bad argument");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(h->stream)))) return rc;
    ff_make_filter(h, kind, f1, f2);           // the rings (data, ovlbuf, inptr) are kept, like create_filter on a live object
    return 0;
}

int b200dsp_fftfilt_filter(b200dsp_fftfilt_t* h, float* out_c64, int cap_samples)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    if (out_c64 && cap_samples >= h->flen) memcpy(out_c64, h->filter.data(), (size_t) h->flen * 8);
    return h->flen;
}

int64_t b200dsp_fftfilt_out_count(b200dsp_fftfilt_t* h, int64_t n_samples)
{
    if (!h || n_samples < 0) return -1;
    const int64_t N2 = h->flen / 2;
    return ((h->inptr + n_samples) / N2) * N2;
}

int b200dsp_fftfilt_run_dev(b200dsp_fftfilt_t* h, int op, int usb, int get_dc, const void* d_in_c64, int64_t n_samples, void* d_out_c64,
                            int64_t cap_samples, int64_t* n_out, void* cuda_stream)
{
    if (!h || op < 0 || op > 2 || n_samples < 0) return b200_fail(B200DSP_EINVAL, "fftfilt_run: bad argument");
    const int N = h->flen, N2 = N / 2;
    const long long total = (long long) h->inptr + n_samples;
    const long long nb = total / N2;
    if (n_out) *n_out = nb * N2;
    if (n_samples == 0) return 0;
    if (!d_in_c64 || (nb > 0 && !d_out_c64)) return b200_fail(B200DSP_EINVAL, "fftfilt_run: null buffer");
    if (nb * N2 > cap_samples) return b200_fail(B200DSP_EINVAL, "fftfilt_run: output buffer too small (%lld needed)", nb * N2);
    if (nb >= (1ll << 30)) return b200_fail(B200DSP_EINVAL, "fftfilt_run: call too long");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : h->stream;
    if ((rc = ff_upload_mult(h, op, usb, get_dc, st))) return rc;
    const float2* in = (const float2*) d_in_c64;
    if (nb > 0) {
        FfParams p;
        memset(&p, 0, sizeof(p));
        p.in = in; p.pend = h->d_pend[h->pcur]; p.ovl_in = h->d_ovl[h->cur]; p.ovl_out = h->d_ovl[h->cur ^ 1];
        p.mult = h->d_mult; p.tw = h->d_tw; p.out = (float2*) d_out_c64; p.inptr = h->inptr; p.nb = (int) nb; p.flen = N; p.log2n = h->log2n;
        const bool fast = (N == 1024 || N == 2048);
        // the fast kernel's CTAs are len/16 threads: as many ranges as CTAs fit the GPU at once
        int per_sm = 4;
        const size_t fast_smem = (size_t) (2 * N + N / 16) * sizeof(float2);          // 16.5 / 33 KB: under the 48 KB default
        if (fast) {
            const void* fn = (N == 1024) ? (const void*) fftfilt_fast_kernel<10> : (const void*) fftfilt_fast_kernel<11>;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, N / 16, fast_smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 4; }
        }
        long long ctas = (long long) h->sm_count * per_sm;
        if (ctas > (nb + 7) / 8) ctas = (nb + 7) / 8;             // a range pays one recomputed block: at least 8 blocks per CTA
        if (ctas < 1) ctas = 1;
        p.blocks_per_cta = (int) ((nb + ctas - 1) / ctas);
        ctas = (nb + p.blocks_per_cta - 1) / p.blocks_per_cta;
        if (fast) {
            if (N == 1024) fftfilt_fast_kernel<10><<<(unsigned) ctas, N / 16, fast_smem, st>>>(p);
            else           fftfilt_fast_kernel<11><<<(unsigned) ctas, N / 16, fast_smem, st>>>(p);
        } else {
            const size_t smem = (size_t) 2 * N * sizeof(float2);
            if (smem > 48 * 1024 && (rc = B200_CUDA_CHECK(cudaFuncSetAttribute((const void*) fftfilt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)))) return rc;
            fftfilt_kernel<<<(unsigned) ctas, FF_THREADS, smem, st>>>(p);
        }
        if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
        h->cur ^= 1;
    }
    // what is left waits for the next call (fftfilt's `data` fill, inptr)
    const long long rem = total - nb * N2;
    if (nb == 0) {
        rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_pend[h->pcur] + h->inptr, in, (size_t) n_samples * 8, cudaMemcpyDeviceToDevice, st));
    } else {
        if (rem > 0) rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_pend[h->pcur ^ 1], in + (n_samples - rem), (size_t) rem * 8, cudaMemcpyDeviceToDevice, st));
        h->pcur ^= 1;             // the kernel may still be reading the old pending samples
    }
    if (rc) return rc;
    h->inptr = (int) rem;
    return 0;
}

int b200dsp_fftfilt_run(b200dsp_fftfilt_t* h, int op, int usb, int get_dc, const float* in_c64, int64_t n_samples, float* out_c64,
                        int64_t cap_samples, int64_t* n_out)
{
    if (!h || op < 0 || op > 2 || n_samples < 0 || (n_samples > 0 && !in_c64)) return b200_fail(B200DSP_EINVAL, "fftfilt_run: bad argument");
    const long long m = b200dsp_fftfilt_out_count(h, n_samples);
    if (n_out) *n_out = m;
    if (n_samples == 0) return 0;
    if (m > cap_samples || (m > 0 && !out_c64)) return b200_fail(B200DSP_EINVAL, "fftfilt_run: output buffer too small (%lld needed)", m);
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    const long long need = (n_samples > m ? n_samples : m) + 8;
    if (h->cap < need) {
        if (h->d_in) cudaFree(h->d_in);
        if (h->d_out) cudaFree(h->d_out);
        h->d_in = nullptr; h->d_out = nullptr; h->cap = 0;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_in, (size_t) need * 8))) || (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_out, (size_t) need * 8)))) return rc;
        h->cap = need;
    }
    int64_t mm = 0;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_in, in_c64, (size_t) n_samples * 8, cudaMemcpyHostToDevice, h->stream))) ||
        (rc = b200dsp_fftfilt_run_dev(h, op, usb, get_dc, h->d_in, n_samples, h->d_out, h->cap, &mm, nullptr))) return rc;
    if (mm > 0 && (rc = B200_CUDA_CHECK(cudaMemcpyAsync(out_c64, h->d_out, (size_t) mm * 8, cudaMemcpyDeviceToHost, h->stream)))) return rc;
    return B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

} // extern "C"
