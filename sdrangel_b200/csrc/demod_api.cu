// demod_api.cu — SURVEY.md 8f-4: demodulator back-ends after Interpolator::decimate (b200dsp_demod_*) and the .sdriq record
// format (b200dsp_sdriq_*, host-side file I/O only).
//
// Replaces (paths relative to the reference tree):
//   PhaseDiscriminators::phaseDiscriminator / phaseDiscriminatorDelta / phaseDiscriminator2   sdrbase/dsp/phasediscri.h:48-96,162-194
//     (NFMDemod::feed calls phaseDiscriminatorDelta on every Interpolator::decimate output, plugins/channelrx/demodnfm/nfmdemod.cpp:165)
//   the magnitude lines of AMDemod::processOneSample                                           plugins/channelrx/demodam/amdemod.cpp:154-156,241
//   FileRecord::writeHeader / readHeader / feed                                                 sdrbase/dsp/filerecord.cpp:72-148
// The discriminators look back one (two) samples: data-parallel over a block, the look-back of the block's first samples is the
// carried state (m_m1Sample, m_m2Sample, m_prevArg).  One launch serves every channel of a pooled block
// [channel][stride] with per-channel device counts -- the layout b200dsp_bank_gather_dev(STAGE_FRONTEND) produces.
// Arithmetic is written with explicitly rounded operations in the reference's order: kinds 1-3 are bit-identical to the
// reference compiled without -ffast-math; kind 0 differs only by the device's atan2f (<= 2 ulp).
#include "common.cuh"
#include <stdio.h>
#include <vector>

using namespace b200dsp;

namespace {

constexpr int DM_STATE = 8;          // per channel: m1.re, m1.im, m2.re, m2.im, prevArg, pad

struct DemodParams {
    const float2* pool; long long stride;
    const long long* counts;
    float* out; long long out_stride;
    float* aux0; float* aux1;
    const float* st_in; float* st_out;
    float scaling;
    int kind;
    int vec;                   // strides and base pointers allow 128-bit accesses
};

// phasediscri.h:162-194
__device__ __forceinline__ float dm_atan2_approx2(float y, float x)
{
    const float PI_FLOAT = 3.14159265f, PIBY2_FLOAT = 1.5707963f;
    if (x == 0.0f) {
        if (y > 0.0f) return PIBY2_FLOAT;
        if (y == 0.0f) return 0.0f;
        return -PIBY2_FLOAT;
    }
    float at;
    const float z = __fdiv_rn(y, x);
    if (fabsf(z) < 1.0f) {
        at = __fdiv_rn(z, __fadd_rn(1.0f, __fmul_rn(__fmul_rn(0.28f, z), z)));
        if (x < 0.0f) {
            if (y < 0.0f) return __fsub_rn(at, PI_FLOAT);
            return __fadd_rn(at, PI_FLOAT);
        }
    } else {
        at = __fsub_rn(PIBY2_FLOAT, __fdiv_rn(z, __fadd_rn(__fmul_rn(z, z), 0.28f)));
        if (y < 0.0f) return __fsub_rn(at, PI_FLOAT);
    }
    return at;
}

// A thread owns 4 consecutive samples: the look-back values (previous sample's argument for kind 1) are computed once per
// thread instead of once per sample, loads and stores are 128-bit when the strides allow.
template<int KIND>
__device__ __forceinline__ void demod_run4(const DemodParams& p, const float2* __restrict__ x, long long n, const float* si, float* so, int c)
{
    const double PI = 3.14159265358979323846;
    const float2 m1s = make_float2(si[0], si[1]), m2s = make_float2(si[2], si[3]);
    const float prev_s = si[4];
    const bool vec = p.vec != 0;
    for (long long i0 = 4 * ((long long) blockIdx.x * blockDim.x + threadIdx.x); i0 < n; i0 += 4ll * gridDim.x * blockDim.x) {
        float2 s[6];                               // s[k] = x[i0 - 2 + k]
        s[0] = (i0 >= 2) ? x[i0 - 2] : m2s;        // (i0 is a multiple of 4: i0 - 2 < 0 only for i0 == 0)
        s[1] = (i0 >= 1) ? x[i0 - 1] : m1s;
        if (i0 == 0) s[0] = m2s;
        if (vec && i0 + 4 <= n) {
            const float4 a = *reinterpret_cast<const float4*>(x + i0), b = *reinterpret_cast<const float4*>(x + i0 + 2);
            s[2] = make_float2(a.x, a.y); s[3] = make_float2(a.z, a.w); s[4] = make_float2(b.x, b.y); s[5] = make_float2(b.z, b.w);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) s[2 + r] = (i0 + r < n) ? x[i0 + r] : make_float2(0.0f, 0.0f);
        }
        float o[4], a0[4], a1[4];
        float prev = 0.0f, arg = 0.0f;
        if (KIND == 1) prev = (i0 >= 1) ? dm_atan2_approx2(s[1].y, s[1].x) : prev_s;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float2 v = s[2 + r], m1 = s[1 + r], m2 = s[r];
            a0[r] = 0.0f; a1[r] = 0.0f;
            if (KIND == 0) {
                // Complex d(std::conj(m_m1Sample) * sample); atan2(d.imag(), d.real()) / M_PI * m_fmScaling     (phasediscri.h:48-53)
                const float dr = __fadd_rn(__fmul_rn(m1.x, v.x), __fmul_rn(m1.y, v.y));
                const float di = __fsub_rn(__fmul_rn(m1.x, v.y), __fmul_rn(m1.y, v.x));
                o[r] = (float) __dmul_rn(__ddiv_rn((double) atan2f(di, dr), PI), (double) p.scaling);
            } else if (KIND == 1) {
                // phaseDiscriminatorDelta (phasediscri.h:59-77)
                a0[r] = __fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y));
                arg = dm_atan2_approx2(v.y, v.x);
                float dev = (float) __ddiv_rn((double) __fsub_rn(arg, prev), PI);
                if (dev < -1.0f) dev = __fadd_rn(dev, 2.0f);
                else if (dev > 1.0f) dev = __fsub_rn(dev, 2.0f);
                a1[r] = dev;
                o[r] = __fmul_rn(dev, p.scaling);
                if (i0 + r == n - 1) so[4] = arg;
                prev = arg;
            } else if (KIND == 2) {
                // phaseDiscriminator2 (phasediscri.h:84-96)
                const float ip = __fsub_rn(v.x, m2.x), qp = __fsub_rn(v.y, m2.y);
                o[r] = __fmul_rn(__fsub_rn(__fmul_rn(m1.x, qp), __fmul_rn(m1.y, ip)), p.scaling);
            } else {
                // AMDemod::processOneSample (amdemod.cpp:154-156,241)
                const float re = __fdiv_rn(v.x, 32768.0f), im = __fdiv_rn(v.y, 32768.0f);
                a0[r] = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
                o[r] = __fsqrt_rn(a0[r]);
            }
            if (i0 + r == n - 1) {
                // the object's members after the block: kinds 0 and 1 leave m_m2Sample alone, kind 1 also m_m1Sample
                const bool upd1 = (KIND == 0 || KIND == 2);
                so[0] = upd1 ? v.x : m1s.x; so[1] = upd1 ? v.y : m1s.y;
                const float2 m2n = (KIND == 2) ? m1 : m2s;
                so[2] = m2n.x; so[3] = m2n.y;
                if (KIND != 1) so[4] = prev_s;
                so[5] = 0.0f; so[6] = 0.0f; so[7] = 0.0f;
            }
        }
        const long long ob = (long long) c * p.out_stride + i0;
        if (vec && i0 + 4 <= n) {
            *reinterpret_cast<float4*>(p.out + ob) = make_float4(o[0], o[1], o[2], o[3]);
            if (p.aux0) *reinterpret_cast<float4*>(p.aux0 + ob) = make_float4(a0[0], a0[1], a0[2], a0[3]);
            if (p.aux1) *reinterpret_cast<float4*>(p.aux1 + ob) = make_float4(a1[0], a1[1], a1[2], a1[3]);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) if (i0 + r < n) {
                p.out[ob + r] = o[r];
                if (p.aux0) p.aux0[ob + r] = a0[r];
                if (p.aux1) p.aux1[ob + r] = a1[r];
            }
        }
    }
}

__global__ void __launch_bounds__(256) demod_kernel(const DemodParams p)
{
    const int c = blockIdx.y;
    const long long n = p.counts ? p.counts[c] : 0;
    const float2* __restrict__ x = p.pool + (long long) c * p.stride;
    const float* si = p.st_in + c * DM_STATE;
    float* so = p.st_out + c * DM_STATE;
    if (n <= 0) {
        if (blockIdx.x == 0 && threadIdx.x < DM_STATE) so[threadIdx.x] = si[threadIdx.x];
        return;
    }
    if (p.kind == 0) demod_run4<0>(p, x, n, si, so, c);
    else if (p.kind == 1) demod_run4<1>(p, x, n, si, so, c);
    else if (p.kind == 2) demod_run4<2>(p, x, n, si, so, c);
    else demod_run4<3>(p, x, n, si, so, c);
}

} // namespace

struct b200dsp_demod {
    int device = 0; cudaStream_t stream = nullptr;
    int kind = 0, n_channels = 1; float scaling = 1.0f;
    float* d_state[2] = { nullptr, nullptr }; int cur = 0;
    // single-stream host form
    float2* d_in = nullptr; float* d_out = nullptr; long long cap = 0; long long* d_count = nullptr;
};

struct b200dsp_sdriq {
    FILE* f = nullptr;
    int writing = 0, header_done = 0;
    int32_t rate = 0; uint64_t center = 0; int64_t ts = 0; uint32_t sample_size = 16;
    int64_t n_samples = 0;
};

extern "C" {

int b200dsp_demod_create(b200dsp_demod_t** out, int kind, float fm_scaling, int n_channels)
{
    if (!out) return b200_fail(B200DSP_EINVAL, "demod_create: null handle pointer");
    *out = nullptr;
    if (kind < 0 || kind > 3 || n_channels < 1 || n_channels > 65535) return b200_fail(B200DSP_EINVAL, "demod_create: kind 0..3, 1..65535 channels");
    int rc = b200_require_device();
    if (rc) return rc;
    b200dsp_demod* h = new (std::nothrow) b200dsp_demod();
    if (!h) return b200_fail(B200DSP_ENOMEM, "demod_create: out of host memory");
    h->device = b200_current_device(); h->kind = kind; h->scaling = fm_scaling; h->n_channels = n_channels;
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(h->device))) || (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)))) { b200dsp_demod_destroy(h); return rc; }
    for (int k = 0; k < 2; ++k)
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_state[k], (size_t) n_channels * DM_STATE * 4))) ||
            (rc = B200_CUDA_CHECK(cudaMemset(h->d_state[k], 0, (size_t) n_channels * DM_STATE * 4)))) { b200dsp_demod_destroy(h); return rc; }
    if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_count, 8)))) { b200dsp_demod_destroy(h); return rc; }
    *out = h;
    return 0;
}

int b200dsp_demod_destroy(b200dsp_demod_t* h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    for (int k = 0; k < 2; ++k) if (h->d_state[k]) cudaFree(h->d_state[k]);
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_out) cudaFree(h->d_out);
    if (h->d_count) cudaFree(h->d_count);
    delete h;
    return 0;
}

int b200dsp_demod_reset(b200dsp_demod_t* h)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    for (int k = 0; k < 2; ++k) if ((rc = B200_CUDA_CHECK(cudaMemsetAsync(h->d_state[k], 0, (size_t) h->n_channels * DM_STATE * 4, h->stream)))) return rc;
    return B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

int b200dsp_demod_set_fm_scaling(b200dsp_demod_t* h, float fm_scaling)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    h->scaling = fm_scaling;
    return 0;
}

int b200dsp_demod_run_pool_dev(b200dsp_demod_t* h, const void* d_pool_c64, int64_t stride_samples, const int64_t* d_counts, int n_channels,
                               float* d_out, int64_t out_stride, float* d_aux0, float* d_aux1, void* cuda_stream)
{
    if (!h || !d_pool_c64 || !d_counts || !d_out || stride_samples <= 0 || out_stride <= 0) return b200_fail(B200DSP_EINVAL, "demod_run_pool: bad argument");
    if (n_channels != h->n_channels) return b200_fail(B200DSP_EINVAL, "demod_run_pool: the handle was created for %d channels, %d given", h->n_channels, n_channels);
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : h->stream;
    DemodParams p;
    memset(&p, 0, sizeof(p));
    p.pool = (const float2*) d_pool_c64; p.stride = stride_samples; p.counts = (const long long*) d_counts;
    p.out = d_out; p.out_stride = out_stride; p.aux0 = d_aux0; p.aux1 = d_aux1;
    p.st_in = h->d_state[h->cur]; p.st_out = h->d_state[h->cur ^ 1];
    p.scaling = h->scaling; p.kind = h->kind;
    p.vec = ((stride_samples & 1) == 0 && (out_stride & 3) == 0 && ((uintptr_t) d_pool_c64 & 15) == 0 && ((uintptr_t) d_out & 15) == 0 &&
             ((uintptr_t) d_aux0 & 15) == 0 && ((uintptr_t) d_aux1 & 15) == 0) ? 1 : 0;
    // CTAs per channel: enough to fill the chip, and a total that is close to a whole number of resident waves (2 x 1024
    // CTAs on 1184 slots ran 1.73 waves: the second one three quarters full)
    long long gmax = (stride_samples + 1023) / 1024;
    if (gmax < 1) gmax = 1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*) demod_kernel, 256, 0) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 8; }
    const long long slots = (long long) b200_sm_count_of(h->device) * per_sm;
    const long long want = slots / n_channels + 1;
    long long gx = want < gmax ? want : gmax;
    double best = 0.0;
    for (long long g = want; g <= 8 * want && g <= gmax; ++g) {
        const double waves = (double) (g * n_channels) / (double) slots;
        const double eff = waves / (double) (long long) (waves + 0.999999);
        if (eff > best + 0.01) { best = eff; gx = g; }
        if (eff >= 0.97) break;
    }
    demod_kernel<<<dim3((unsigned) gx, (unsigned) n_channels), 256, 0, st>>>(p);
    if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
    h->cur ^= 1;
    return 0;
}

int b200dsp_demod_run(b200dsp_demod_t* h, const float* in_c64, int64_t n, float* out, float* aux0, float* aux1)
{
    if (!h || n < 0 || (n > 0 && (!in_c64 || !out))) return b200_fail(B200DSP_EINVAL, "demod_run: bad argument");
    if (h->n_channels != 1) return b200_fail(B200DSP_EINVAL, "demod_run: single-stream call on a %d-channel handle", h->n_channels);
    if (n == 0) return 0;
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    if (h->cap < n) {
        if (h->d_in) cudaFree(h->d_in);
        if (h->d_out) cudaFree(h->d_out);
        h->d_in = nullptr; h->d_out = nullptr; h->cap = 0;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_in, (size_t) n * 8))) || (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_out, (size_t) n * 12)))) return rc;
        h->cap = n;
    }
    const long long cnt = n;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_in, in_c64, (size_t) n * 8, cudaMemcpyHostToDevice, h->stream))) ||
        (rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_count, &cnt, 8, cudaMemcpyHostToDevice, h->stream))) ||
        (rc = b200dsp_demod_run_pool_dev(h, h->d_in, h->cap, (const int64_t*) h->d_count, 1, h->d_out, h->cap, h->d_out + h->cap, h->d_out + 2 * h->cap, nullptr))) return rc;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(out, h->d_out, (size_t) n * 4, cudaMemcpyDeviceToHost, h->stream)))) return rc;
    if (aux0 && (rc = B200_CUDA_CHECK(cudaMemcpyAsync(aux0, h->d_out + h->cap, (size_t) n * 4, cudaMemcpyDeviceToHost, h->stream)))) return rc;
    if (aux1 && (rc = B200_CUDA_CHECK(cudaMemcpyAsync(aux1, h->d_out + 2 * h->cap, (size_t) n * 4, cudaMemcpyDeviceToHost, h->stream)))) return rc;
    return B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

// ---- .sdriq (host-side, no device) ----------------------------------------------------------------------------------
int b200dsp_sdriq_header_encode(int32_t sample_rate, uint64_t center_frequency, int64_t start_timestamp, uint32_t sample_size, void* out24)
{
    if (!out24) return b200_fail(B200DSP_EINVAL, "sdriq_header_encode: null buffer");
    unsigned char* o = (unsigned char*) out24;            // FileRecord::writeHeader: four raw writes, native byte order, no padding
    memcpy(o, &sample_rate, 4); memcpy(o + 4, &center_frequency, 8); memcpy(o + 12, &start_timestamp, 8); memcpy(o + 20, &sample_size, 4);
    return 0;
}

int b200dsp_sdriq_header_decode(const void* in24, int32_t* sample_rate, uint64_t* center_frequency, int64_t* start_timestamp, uint32_t* sample_size)
{
    if (!in24) return b200_fail(B200DSP_EINVAL, "sdriq_header_decode: null buffer");
    const unsigned char* i = (const unsigned char*) in24;
    int32_t r; uint64_t c; int64_t t; uint32_t s;
    memcpy(&r, i, 4); memcpy(&c, i + 4, 8); memcpy(&t, i + 12, 8); memcpy(&s, i + 20, 4);
    if (s != 16 && s != 24) s = 16;                       // "assume 16 bits if garbage (old I/Q file)", filerecord.cpp:145-147
    if (sample_rate) *sample_rate = r;
    if (center_frequency) *center_frequency = c;
    if (start_timestamp) *start_timestamp = t;
    if (sample_size) *sample_size = s;
    return 0;
}

int b200dsp_sdriq_open(b200dsp_sdriq_t** out, const char* path, int32_t* sample_rate, uint64_t* center_frequency, int64_t* start_timestamp,
                       uint32_t* sample_size, int64_t* n_samples)
{
    if (!out || !path) return b200_fail(B200DSP_EINVAL, "sdriq_open: bad argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return b200_fail(B200DSP_EINVAL, "sdriq_open: cannot open %s", path);
    unsigned char hdr[24];
    if (fread(hdr, 1, 24, f) != 24) { fclose(f); return b200_fail(B200DSP_EINVAL, "sdriq_open: %s is shorter than the 24-byte header", path); }
    b200dsp_sdriq* r = new (std::nothrow) b200dsp_sdriq();
    if (!r) { fclose(f); return b200_fail(B200DSP_ENOMEM, "sdriq_open: out of host memory"); }
    r->f = f;
    b200dsp_sdriq_header_decode(hdr, &r->rate, &r->center, &r->ts, &r->sample_size);
    fseek(f, 0, SEEK_END);
    const long long bytes = ftell(f) - 24;
    fseek(f, 24, SEEK_SET);
    r->n_samples = bytes / (r->sample_size == 24 ? 8 : 4);     // Sample = 2 x int16 (16-bit records) or 2 x int32 (24-bit records)
    if (sample_rate) *sample_rate = r->rate;
    if (center_frequency) *center_frequency = r->center;
    if (start_timestamp) *start_timestamp = r->ts;
    if (sample_size) *sample_size = r->sample_size;
    if (n_samples) *n_samples = r->n_samples;
    *out = r;
    return 0;
}

int b200dsp_sdriq_read(b200dsp_sdriq_t* r, int16_t* iq, int64_t cap_samples, int64_t* got)
{
    if (!r || r->writing || cap_samples < 0 || (cap_samples > 0 && !iq)) return b200_fail(B200DSP_EINVAL, "sdriq_read: bad argument");
    if (r->sample_size != 16) return b200_fail(B200DSP_ESTATE, "sdriq_read: %u-bit records; this build is the 16-bit Rx mode (SDR_RX_SAMP_SZ 16)", r->sample_size);
    const size_t n = fread(iq, 4, (size_t) cap_samples, r->f);
    if (got) *got = (int64_t) n;
    return 0;
}

int b200dsp_sdriq_create(b200dsp_sdriq_t** out, const char* path, int32_t sample_rate, uint64_t center_frequency, int64_t start_timestamp)
{
    if (!out || !path) return b200_fail(B200DSP_EINVAL, "sdriq_create: bad argument");
    *out = nullptr;
    FILE* f = fopen(path, "wb");
    if (!f) return b200_fail(B200DSP_EINVAL, "sdriq_create: cannot open %s", path);
    b200dsp_sdriq* w = new (std::nothrow) b200dsp_sdriq();
    if (!w) { fclose(f); return b200_fail(B200DSP_ENOMEM, "sdriq_create: out of host memory"); }
    w->f = f; w->writing = 1; w->rate = sample_rate; w->center = center_frequency; w->ts = start_timestamp; w->sample_size = 16;
    *out = w;
    return 0;
}

int b200dsp_sdriq_write(b200dsp_sdriq_t* w, const int16_t* iq, int64_t n_samples)
{
    if (!w || !w->writing || n_samples < 0 || (n_samples > 0 && !iq)) return b200_fail(B200DSP_EINVAL, "sdriq_write: bad argument");
    if (n_samples == 0) return 0;                          // FileRecord::feed: nothing to put out, not even the header (filerecord.cpp:80-90)
    if (!w->header_done) {
        unsigned char hdr[24];
        b200dsp_sdriq_header_encode(w->rate, w->center, w->ts, w->sample_size, hdr);
        if (fwrite(hdr, 1, 24, w->f) != 24) return b200_fail(B200DSP_EINVAL, "sdriq_write: write failed");
        w->header_done = 1;
    }
    if (fwrite(iq, 4, (size_t) n_samples, w->f) != (size_t) n_samples) return b200_fail(B200DSP_EINVAL, "sdriq_write: write failed");
    w->n_samples += n_samples;
    return 0;
}

int b200dsp_sdriq_close(b200dsp_sdriq_t* r)
{
    if (!r) return 0;
    if (r->f) fclose(r->f);
    delete r;
    return 0;
}

} // extern "C"
