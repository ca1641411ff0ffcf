// iqcorr_api.cu — C ABI of the engine-side sample corrections: b200dsp_iqcorr_*
//
// Reference: DSPDeviceSourceEngine::iqCorrections (sdrbase/dsp/dspdevicesourceengine.cpp:175-262), the step between the
// device Decimators<> and the channel sinks (called from DSPDeviceSourceEngine::work, :343-347,375-379).  DC branch
// (imbalanceCorrection == false, :254-259):
//     m_iBeta(it->real()); m_qBeta(it->imag());                 MovingAverageUtil<int32_t, int64_t, 1024>  (dspdevicesourceengine.h:106-107)
//     it->m_real -= (int32_t) m_iBeta; it->m_imag -= (int32_t) m_qBeta;
// MovingAverageUtil (sdrbase/util/movingaverage.h): running total of the last 1024 pushed samples (fewer while filling
// up, but the divisor is always N), operator T() = total / N with C++ truncation toward zero.  So
//     y[i] = (int16) (x[i] - trunc(S_i / 1024)),   S_i = sum of the (raw) samples x[max(0, i-1023) .. i] of the stream.
// The sequential running total becomes a windowed sum of raw inputs.  A block takes a tile of the stream plus a 1024-sample
// halo; a thread owns 16 consecutive samples in registers; one warp-shuffle scan of the thread totals gives every thread
// the window sum at its first sample, S(16t + k) = sum of the 64 thread totals before t + sum_{m<=k} (x_t[m] - x_{t-64}[m]),
// and the samples of thread t-64 (the ones that leave the window) come through shared memory as packed words.  The
// carried state is the last 1024 raw samples (zeros at start == the fill-up phase).
#include "common.cuh"

using namespace b200dsp;

namespace {

constexpr int DC_N = 1024;            // MovingAverageUtil<..., 1024>
constexpr int DC_TILE = 3072;         // samples per block (measured: 3072/256 threads 2.9 TB/s; 7168/512 threads 2.7 TB/s -- more, smaller CTAs per SM win)
constexpr int DC_THREADS = 256;
constexpr int DC_PER = (DC_TILE + DC_N) / DC_THREADS;     // 16 consecutive elements of (halo + tile) per thread
constexpr int DC_HALO_THREADS = DC_N / DC_PER;            // 64: thread t's window loses the samples of thread t - 64
constexpr int DC_SLOT = DC_PER + 4;                       // words per thread in the exchange array: 128-bit accesses without bank conflicts
static_assert(DC_PER == 16 && DC_HALO_THREADS == 64 && DC_THREADS % 32 == 0, "the window arithmetic below assumes 16 samples per thread and a halo of two warps");

struct DcParams {
    const uint32_t* in;        // packed int16 IQ
    uint32_t* out;
    const uint32_t* hist_in;   // last DC_N raw samples before this call (oldest first)
    uint32_t* hist_out;
    long long n;
    int in_vec, out_vec;       // in / out is 16-byte aligned: 128-bit accesses
};

// the thread's 16 samples of one tile: four 128-bit loads.  rem = how many of them belong to the call (history counts)
__device__ __forceinline__ void dc_load(const DcParams& p, long long tile, int tid, uint32_t (&raw)[DC_PER], int& rem)
{
    const long long i0 = tile * DC_TILE - DC_N + DC_PER * tid;      // stream index of the first sample (a multiple of 16; < 0: history)
    const bool hist = i0 < 0;                                       // (all 16 samples are history, or none)
    const long long left = p.n - i0;
    rem = hist ? DC_PER : (left > DC_PER ? DC_PER : (left < 0 ? 0 : (int) left));
    const uint32_t* src = hist ? p.hist_in + (DC_N + i0) : p.in + i0;
    const bool vec = hist || p.in_vec;
#pragma unroll
    for (int g = 0; g < DC_PER / 4; ++g) {
        uint4 v;
        if (vec && 4 * g + 4 <= rem) v = __ldg(reinterpret_cast<const uint4*>(src + 4 * g));
        else {
            v.x = (4 * g < rem) ? src[4 * g] : 0u;         v.y = (4 * g + 1 < rem) ? src[4 * g + 1] : 0u;
            v.z = (4 * g + 2 < rem) ? src[4 * g + 2] : 0u; v.w = (4 * g + 3 < rem) ? src[4 * g + 3] : 0u;
        }
        raw[4 * g] = v.x; raw[4 * g + 1] = v.y; raw[4 * g + 2] = v.z; raw[4 * g + 3] = v.w;
    }
}

// persistent: block b takes tiles b, b + gridDim.x, ...; the next tile's loads are issued before this tile's arithmetic
__global__ void __launch_bounds__(DC_THREADS, 4) dc_correct_kernel(const DcParams p)
{
    __shared__ __align__(16) uint32_t sraw[DC_THREADS * DC_SLOT];      // every thread's 16 packed raw samples
    __shared__ int sex_r[DC_THREADS], sex_i[DC_THREADS];               // sum of the totals of the lower lanes of its warp
    __shared__ int wre[DC_THREADS / 32], wim[DC_THREADS / 32];         // warp totals
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long tiles = (p.n + DC_TILE - 1) / DC_TILE;
    // carry: the last DC_N raw samples of the stream so far (also covers calls shorter than DC_N)
    if (blockIdx.x == 0) {
        for (int t = tid; t < DC_N; t += DC_THREADS) {
            const long long i = p.n - DC_N + t;
            p.hist_out[t] = (i < 0) ? p.hist_in[DC_N + i] : p.in[i];
        }
    }
    long long tile = blockIdx.x;
    if (tile >= tiles) return;
    uint32_t raw[DC_PER];
    int rem;
    dc_load(p, tile, tid, raw, rem);
    for (;;) {
        const long long nxt = tile + gridDim.x;
        // 1. a copy of the samples goes to shared memory for thread tid + 64
#pragma unroll
        for (int g = 0; g < DC_PER / 4; ++g)
            *reinterpret_cast<uint4*>(&sraw[tid * DC_SLOT + 4 * g]) = make_uint4(raw[4 * g], raw[4 * g + 1], raw[4 * g + 2], raw[4 * g + 3]);
        uint32_t nraw[DC_PER];
        int nrem = 0;
        if (nxt < tiles) dc_load(p, nxt, tid, nraw, nrem);
        // 2. thread totals, inclusive warp scan
        int sr = 0, si = 0;
#pragma unroll
        for (int k = 0; k < DC_PER; ++k) { sr += (int) (short) (raw[k] & 0xffffu); si += (int) raw[k] >> 16; }
        int xr = sr, xi = si;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, xr, d), b = __shfl_up_sync(0xffffffffu, xi, d);
            if (lane >= d) { xr += a; xi += b; }
        }
        sex_r[tid] = xr - sr; sex_i[tid] = xi - si;
        if (lane == 31) { wre[wid] = xr; wim[wid] = xi; }
        __syncthreads();
        // 3. outputs (tile threads): S = sum of the 1024 samples that end at the current one, y = x - trunc(S / 1024)
        if (tid >= DC_HALO_THREADS && rem > 0) {
            const int pt = tid - DC_HALO_THREADS;           // same lane, two warps down
            int Sr = wre[wid - 2] + wre[wid - 1] + (xr - sr) - sex_r[pt];
            int Si = wim[wid - 2] + wim[wid - 1] + (xi - si) - sex_i[pt];
            uint32_t* out0 = p.out + (tile * DC_TILE - DC_N + DC_PER * tid);
#pragma unroll
            for (int g = 0; g < DC_PER / 4; ++g) {
                const uint4 pv = *reinterpret_cast<const uint4*>(&sraw[pt * DC_SLOT + 4 * g]);
                const uint32_t pw[4] = { pv.x, pv.y, pv.z, pv.w };
                uint32_t ow[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const uint32_t w = raw[4 * g + m], q = pw[m];
                    const int x = (int) (short) (w & 0xffffu), y = (int) w >> 16;
                    Sr += x - (int) (short) (q & 0xffffu);
                    Si += y - ((int) q >> 16);
                    const int re = x - Sr / DC_N, im = y - Si / DC_N;         // C++ integer division: toward zero, like total / N
                    ow[m] = ((uint32_t) re & 0xffffu) | ((uint32_t) im << 16);
                }
                uint32_t* dst = out0 + 4 * g;
                if (p.out_vec && 4 * g + 4 <= rem) *reinterpret_cast<uint4*>(dst) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                else {
#pragma unroll
                    for (int m = 0; m < 4; ++m) if (4 * g + m < rem) dst[m] = ow[m];
                }
            }
        }
        if (nxt >= tiles) break;
        __syncthreads();                                    // every thread has read this tile's shared arrays
#pragma unroll
        for (int k = 0; k < DC_PER; ++k) raw[k] = nraw[k];
        rem = nrem; tile = nxt;
    }
}

// ---------------------------------------------------------------------------------------------------------
// I/Q imbalance branch (dspdevicesourceengine.cpp:219-252, the floating-point flavour: IMBALANCE_INT is not defined).
// Per sample: DC-corrected (xi, xq) / 32768; <I,I>, <I,Q> over 128 samples -> phase estimate phi = <I,Q>/<I,I>, itself averaged
// over 128 pushes (pushed only while <I,I> != 0); yq = xq - phi xi; <I,I>, <Q,Q> of (xi, yq) -> amplitude estimate
// sqrt(<I,I>/<Q,Q>), averaged over 128 pushes; zq = amp yq; output (xi, zq) * 32768 truncated to int16.  Every average
// is a MovingAverageUtil (util/movingaverage.h) whose running total adds `sample - oldest` -- a float subtraction for the
// power averages -- so the loop is sequential as written.  All estimators have finite memory (1024 + 4 x 128 samples):
// the stream is cut into segments, ONE THREAD replays the reference loop over a segment in the reference's own
// operation order (separately rounded float/double operations, IEEE division and square root), after a warm-up over the
// 2048 samples before it (from zero state; its outputs are dropped).  What a segment cannot see is the rounding drift
// the reference's running totals accumulated before the warm-up (relative 1e-7 after a million samples: an output LSB
// in ~1e-4 of the samples) and estimator pushes skipped during digital silence older than the warm-up.
// ---------------------------------------------------------------------------------------------------------
constexpr int IMB_WARM = 2048;
constexpr int IMB_SEG = 2048;

struct ImbParams {
    const uint32_t* in; uint32_t* out;
    const uint32_t* hist_in;      // the IMB_WARM raw samples before this call (oldest first)
    uint32_t* hist_out;
    uint32_t* dc_hist_out;        // the DC branch's carried history (last DC_N raw samples): both branches share m_iBeta / m_qBeta
    long long n;
};

struct ImbAvgF { float s[128]; unsigned idx; double total; };
struct ImbAvgD { double s[128]; unsigned idx; double total; };
__device__ __forceinline__ void imb_push(ImbAvgF& m, float v)        // the ring starts full of zeros == the fill-up phase (the divisor is always N)
{
    const float d = __fsub_rn(v, m.s[m.idx]);
    m.total = __dadd_rn(m.total, (double) d);
    m.s[m.idx] = v;
    m.idx = (m.idx + 1) & 127u;
}
__device__ __forceinline__ void imb_push(ImbAvgD& m, double v)
{
    m.total = __dadd_rn(m.total, __dsub_rn(v, m.s[m.idx]));
    m.s[m.idx] = v;
    m.idx = (m.idx + 1) & 127u;
}

__global__ void __launch_bounds__(32) iq_imbalance_kernel(const ImbParams p)
{
    const long long seg = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long start = seg * IMB_SEG;
    if (start < p.n) {
        const long long end = (start + IMB_SEG < p.n) ? start + IMB_SEG : p.n;
        int dr[DC_N], di[DC_N];
        unsigned didx = 0;
        long long tr = 0, ti = 0;
        ImbAvgF aII, aIQ, aII2, aQQ2;
        ImbAvgD aPhi, aAmp;
        for (int k = 0; k < DC_N; ++k) { dr[k] = 0; di[k] = 0; }
        for (int k = 0; k < 128; ++k) { aII.s[k] = 0; aIQ.s[k] = 0; aII2.s[k] = 0; aQQ2.s[k] = 0; aPhi.s[k] = 0; aAmp.s[k] = 0; }
        aII.idx = aIQ.idx = aII2.idx = aQQ2.idx = aPhi.idx = aAmp.idx = 0;
        aII.total = aIQ.total = aII2.total = aQQ2.total = aPhi.total = aAmp.total = 0.0;
        for (long long i = start - IMB_WARM; i < end; ++i) {
            const uint32_t w = (i < 0) ? p.hist_in[IMB_WARM + i] : p.in[i];
            const int re = (int) (short) (w & 0xffffu), im = (int) w >> 16;
            tr += re - dr[didx]; dr[didx] = re;
            ti += im - di[didx]; di[didx] = im;
            didx = (didx + 1) & (DC_N - 1);
            const int bi = (int) (tr / DC_N), bq = (int) (ti / DC_N);                 // operator T(): total / N, toward zero
            const float xi = __fdiv_rn((float) (re - bi), 32768.0f), xq = __fdiv_rn((float) (im - bq), 32768.0f);
            imb_push(aII, __fmul_rn(xi, xi));
            imb_push(aIQ, __fmul_rn(xi, xq));
            const double mII = aII.total / 128.0;
            if (mII != 0.0) imb_push(aPhi, __ddiv_rn(aIQ.total / 128.0, mII));
            const float yq = (float) __dsub_rn((double) xq, __dmul_rn(aPhi.total / 128.0, (double) xi));
            imb_push(aII2, __fmul_rn(xi, xi));
            imb_push(aQQ2, __fmul_rn(yq, yq));
            const double mQQ = aQQ2.total / 128.0;
            if (mQQ != 0.0) imb_push(aAmp, sqrt(__ddiv_rn(aII2.total / 128.0, mQQ)));
            const float zq = (float) __dmul_rn(aAmp.total / 128.0, (double) yq);
            if (i >= start) {
                const int ore = (int) __fmul_rn(xi, 32768.0f), oim = (int) __fmul_rn(zq, 32768.0f);      // C conversion: toward zero
                p.out[i] = ((uint32_t) ore & 0xffffu) | ((uint32_t) oim << 16);
            }
        }
    }
    // carried raw history for the next call: thread 0 of block 0
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int t = 0; t < IMB_WARM; ++t) {
            const long long i = p.n - IMB_WARM + t;
            p.hist_out[t] = (i < 0) ? p.hist_in[IMB_WARM + i] : p.in[i];
        }
        for (int t = 0; t < DC_N; ++t) {
            const long long i = p.n - DC_N + t;
            p.dc_hist_out[t] = (i < 0) ? p.hist_in[IMB_WARM + i] : p.in[i];
        }
    }
}

} // namespace

struct b200dsp_iqcorr {
    int device;
    cudaStream_t stream;
    uint32_t* d_hist[2];
    uint32_t* d_ihist[2];                                     // imbalance branch: the last IMB_WARM raw samples it has seen
    int cur, icur;
    uint32_t* d_in;  uint32_t* d_out; long long cap;       // staging for the host-pointer / in-place forms
};

namespace {

int ensure_staging(b200dsp_iqcorr* h, long long n, cudaStream_t st)
{
    if (h->cap >= n) return 0;
    int rc;
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(st)))) return rc;
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_out) cudaFree(h->d_out);
    h->d_in = nullptr; h->d_out = nullptr; h->cap = 0;
    if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_in, (size_t) n * 4))) || (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_out, (size_t) n * 4)))) return rc;
    h->cap = n;
    return 0;
}

int launch_dc(b200dsp_iqcorr* h, const uint32_t* d_in, uint32_t* d_out, long long n, cudaStream_t st)
{
    DcParams p;
    p.in = d_in; p.out = d_out; p.hist_in = h->d_hist[h->cur]; p.hist_out = h->d_hist[h->cur ^ 1]; p.n = n;
    const long long blocks = (n + DC_TILE - 1) / DC_TILE;
    p.in_vec = (((uintptr_t) d_in & 15) == 0) ? 1 : 0; p.out_vec = (((uintptr_t) d_out & 15) == 0) ? 1 : 0;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*) dc_correct_kernel, DC_THREADS, 0) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
    long long grid = (long long) b200_sm_count_of(h->device) * per_sm;      // persistent: one resident wave, tiles dealt round-robin
    if (grid > blocks) grid = blocks;
    dc_correct_kernel<<<(unsigned) grid, DC_THREADS, 0, st>>>(p);
    int rc = B200_CUDA_CHECK(cudaGetLastError());
    if (rc == 0) h->cur ^= 1;
    return rc;
}

} // namespace

extern "C" {

int b200dsp_iqcorr_create(b200dsp_iqcorr_t** out)
{
    if (!out) return b200_fail(B200DSP_EINVAL, "iqcorr_create: null handle pointer");
    *out = nullptr;
    int rc = b200_require_device();
    if (rc) return rc;
    b200dsp_iqcorr* h = new (std::nothrow) b200dsp_iqcorr();
    if (!h) return b200_fail(B200DSP_ENOMEM, "iqcorr_create: out of host memory");
    memset(h, 0, sizeof(*h));
    h->device = b200_current_device();
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(h->device))) ||
        (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)))) { delete h; return rc; }
    for (int i = 0; i < 2; ++i)
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_hist[i], DC_N * 4))) || (rc = B200_CUDA_CHECK(cudaMemset(h->d_hist[i], 0, DC_N * 4))) ||
            (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_ihist[i], IMB_WARM * 4))) || (rc = B200_CUDA_CHECK(cudaMemset(h->d_ihist[i], 0, IMB_WARM * 4)))) { b200dsp_iqcorr_destroy(h); return rc; }
    if ((rc = B200_CUDA_CHECK(cudaDeviceSynchronize()))) { b200dsp_iqcorr_destroy(h); return rc; }
    *out = h;
    return 0;
}

int b200dsp_iqcorr_destroy(b200dsp_iqcorr_t* h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int i = 0; i < 2; ++i) { if (h->d_hist[i]) cudaFree(h->d_hist[i]); if (h->d_ihist[i]) cudaFree(h->d_ihist[i]); }
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_out) cudaFree(h->d_out);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int b200dsp_iqcorr_reset(b200dsp_iqcorr_t* h)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    for (int i = 0; i < 2; ++i)
        if ((rc = B200_CUDA_CHECK(cudaMemsetAsync(h->d_hist[i], 0, DC_N * 4, h->stream))) || (rc = B200_CUDA_CHECK(cudaMemsetAsync(h->d_ihist[i], 0, IMB_WARM * 4, h->stream)))) return rc;
    return B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

static int check_mode(int imbalance) { (void) imbalance; return 0; }

static int launch_imbalance(b200dsp_iqcorr* h, const uint32_t* d_in, uint32_t* d_out, long long n, cudaStream_t st)
{
    ImbParams p;
    p.in = d_in; p.out = d_out; p.hist_in = h->d_ihist[h->icur]; p.hist_out = h->d_ihist[h->icur ^ 1];
    p.dc_hist_out = h->d_hist[h->cur ^ 1]; p.n = n;
    const long long segs = (n + IMB_SEG - 1) / IMB_SEG;
    // one thread per segment, 32 per block: the rings (12 KB per thread) live in local memory
    iq_imbalance_kernel<<<(unsigned) ((segs + 31) / 32), 32, 0, st>>>(p);
    int rc = B200_CUDA_CHECK(cudaGetLastError());
    if (rc == 0) { h->icur ^= 1; h->cur ^= 1; }
    return rc;
}

int b200dsp_iqcorr_run_dev(b200dsp_iqcorr_t* h, const void* d_in, void* d_out, int64_t n_samples, int imbalance, void* cuda_stream)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    int rc = check_mode(imbalance);
    if (rc) return rc;
    if (n_samples < 0 || (n_samples > 0 && (!d_in || !d_out))) return b200_fail(B200DSP_EINVAL, "iqcorr_run_dev: bad buffer");
    if (n_samples == 0) return 0;
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(h->device)))) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : h->stream;
    const uint32_t* src = (const uint32_t*) d_in;
    if (d_in == d_out) {          // in place (as the reference works): the halo of a tile must stay raw, so read from a copy
        if ((rc = ensure_staging(h, n_samples, st))) return rc;
        if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_in, d_in, (size_t) n_samples * 4, cudaMemcpyDeviceToDevice, st)))) return rc;
        src = h->d_in;
    }
    if (imbalance) return launch_imbalance(h, src, (uint32_t*) d_out, n_samples, st);
    return launch_dc(h, src, (uint32_t*) d_out, n_samples, st);
}

int b200dsp_iqcorr_run(b200dsp_iqcorr_t* h, int16_t* iq, int64_t n_samples, int imbalance)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    int rc = check_mode(imbalance);
    if (rc) return rc;
    if (n_samples < 0 || (n_samples > 0 && !iq)) return b200_fail(B200DSP_EINVAL, "iqcorr_run: bad buffer");
    if (n_samples == 0) return 0;
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(h->device)))) return rc;
    if ((rc = ensure_staging(h, n_samples, h->stream))) return rc;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_in, iq, (size_t) n_samples * 4, cudaMemcpyHostToDevice, h->stream))) ||
        (rc = imbalance ? launch_imbalance(h, h->d_in, h->d_out, n_samples, h->stream) : launch_dc(h, h->d_in, h->d_out, n_samples, h->stream)) ||
        (rc = B200_CUDA_CHECK(cudaMemcpyAsync(iq, h->d_out, (size_t) n_samples * 4, cudaMemcpyDeviceToHost, h->stream)))) return rc;
    return B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

} // extern "C"
