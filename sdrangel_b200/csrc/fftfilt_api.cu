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
        long long ctas = (long long) h->sm_count * 4;
        if (ctas > (nb + 7) / 8) ctas = (nb + 7) / 8;             // a range pays one recomputed block: at least 8 blocks per CTA
        if (ctas < 1) ctas = 1;
        p.blocks_per_cta = (int) ((nb + ctas - 1) / ctas);
        ctas = (nb + p.blocks_per_cta - 1) / p.blocks_per_cta;
        const size_t smem = (size_t) 2 * N * sizeof(float2);
        if (smem > 48 * 1024 && (rc = B200_CUDA_CHECK(cudaFuncSetAttribute((const void*) fftfilt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)))) return rc;
        fftfilt_kernel<<<(unsigned) ctas, FF_THREADS, smem, st>>>(p);
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
