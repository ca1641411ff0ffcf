// nco_api.cu — stand-alone NCO: b200dsp_nco_*
//
// Replaces (paths relative to the reference tree):
//   NCO::initTable / setFreq / nextPhase / nextIQ       sdrbase/dsp/nco.cpp:30-64, nco.h:40-53
//   c *= m_nco.nextIQ()  (channel plugins' per-sample mix)   plugins/channelrx/demodnfm/nfmdemod.cpp:152-153
// The reference NCO's whole state is one integer phase, advanced by a fixed integer increment BEFORE every lookup, so
// sample i of a block sees phase (phase0 + (i + 1) * inc) mod 4096: the block is data-parallel, the state update is
// host arithmetic.  The table is the reference's: T[i] = (float) cos(2 pi i / 4096) evaluated in double.
#include "common.cuh"
#include <math.h>
#include <vector>

using namespace b200dsp;

namespace {

__global__ void nco_iq_kernel(const float* __restrict__ table, int phase0, int inc, long long n, float2* __restrict__ out)
{
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const int p = (int) (((unsigned) phase0 + (unsigned) ((i + 1) & 4095) * (unsigned) inc) & 4095u);      // mod 4096: ((i+1) mod 4096) * inc is congruent
        out[i] = make_float2(table[p], -table[(p + 1024) & 4095]);
    }
}

__global__ void nco_mix_kernel(const float* __restrict__ table, int phase0, int inc, long long n, const uint32_t* __restrict__ in, float2* __restrict__ out)
{
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const uint32_t w = in[i];
        const float x = (float) (short) (w & 0xffffu), y = (float) ((int) w >> 16);
        const int p = (int) (((unsigned) phase0 + (unsigned) ((i + 1) & 4095) * (unsigned) inc) & 4095u);
        const float u = table[p], v = -table[(p + 1024) & 4095];
        out[i] = make_float2(x * u - y * v, x * v + y * u);            // std::complex<float> multiply as the reference's build evaluates it
    }
}

} // namespace

struct b200dsp_nco {
    int device = 0; cudaStream_t stream = nullptr;
    float* d_table = nullptr; float2* d_out = nullptr; long long cap = 0;
    int phase = 0, inc = 0;
};

namespace {
int advance(b200dsp_nco* h, long long n)
{
    // NCO::nextPhase n times: phase += inc, then brought back into [0, 4096) (nco.h:43-50)
    long long p = ((long long) h->phase + (long long) ((n % 4096) * (long long) h->inc)) % 4096;
    if (p < 0) p += 4096;
    h->phase = (int) p;
    return 0;
}
}

extern "C" {

int b200dsp_nco_create(b200dsp_nco_t** out)
{
    if (!out) return b200_fail(B200DSP_EINVAL, "nco_create: null handle pointer");
    *out = nullptr;
    int rc = b200_require_device();
    if (rc) return rc;
    b200dsp_nco* h = new (std::nothrow) b200dsp_nco();
    if (!h) return b200_fail(B200DSP_ENOMEM, "nco_create: out of host memory");
    h->device = b200_current_device();
    std::vector<float> t(4096);
    for (int i = 0; i < 4096; i++) t[i] = (float) cos((2.0 * 3.14159265358979323846 * i) / 4096);       // NCO::initTable, nco.cpp:30-39
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(h->device))) || (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_table, 4096 * sizeof(float)))) ||
        (rc = B200_CUDA_CHECK(cudaMemcpy(h->d_table, t.data(), 4096 * sizeof(float), cudaMemcpyHostToDevice)))) { b200dsp_nco_destroy(h); return rc; }
    *out = h;
    return 0;
}

int b200dsp_nco_destroy(b200dsp_nco_t* h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    if (h->d_table) cudaFree(h->d_table);
    if (h->d_out) cudaFree(h->d_out);
    delete h;
    return 0;
}

int b200dsp_nco_set_freq(b200dsp_nco_t* h, float freq, float sample_rate)
{
    if (!h || !(sample_rate != 0.0f)) return b200_fail(B200DSP_EINVAL, "nco_set_freq: bad argument");
    h->inc = (int) ((freq * 4096) / sample_rate);          // float arithmetic, truncation toward zero (nco.cpp:50)
    return 0;
}

int b200dsp_nco_set_phase(b200dsp_nco_t* h, int phase)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    h->phase = ((phase % 4096) + 4096) % 4096;
    return 0;
}

int b200dsp_nco_get(b200dsp_nco_t* h, int* phase, int* phase_increment)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    if (phase) *phase = h->phase;
    if (phase_increment) *phase_increment = h->inc;
    return 0;
}

int b200dsp_nco_next_iq_dev(b200dsp_nco_t* h, int64_t n, float* d_out_c64, void* cuda_stream)
{
    if (!h || n < 0 || (n > 0 && !d_out_c64)) return b200_fail(B200DSP_EINVAL, "nco_next_iq: bad argument");
    if (n == 0) return 0;
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : h->stream;
    const long long blocks = (n + 255) / 256;
    nco_iq_kernel<<<(unsigned) (blocks < 4096 ? blocks : 4096), 256, 0, st>>>(h->d_table, h->phase, h->inc, n, (float2*) d_out_c64);
    if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
    return advance(h, n);
}

int b200dsp_nco_next_iq(b200dsp_nco_t* h, int64_t n, float* out_c64)
{
    if (!h || n < 0 || (n > 0 && !out_c64)) return b200_fail(B200DSP_EINVAL, "nco_next_iq: bad argument");
    if (n == 0) return 0;
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    if (h->cap < n) {
        if (h->d_out) cudaFree(h->d_out);
        h->d_out = nullptr; h->cap = 0;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_out, (size_t) n * sizeof(float2))))) return rc;
        h->cap = n;
    }
    if ((rc = b200dsp_nco_next_iq_dev(h, n, (float*) h->d_out, nullptr))) return rc;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(out_c64, h->d_out, (size_t) n * sizeof(float2), cudaMemcpyDeviceToHost, h->stream)))) return rc;
    return B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

int b200dsp_nco_mix_dev(b200dsp_nco_t* h, const void* d_in_i16, int64_t n, float* d_out_c64, void* cuda_stream)
{
    if (!h || n < 0 || (n > 0 && (!d_in_i16 || !d_out_c64))) return b200_fail(B200DSP_EINVAL, "nco_mix: bad argument");
    if (n == 0) return 0;
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : h->stream;
    const long long blocks = (n + 255) / 256;
    nco_mix_kernel<<<(unsigned) (blocks < 4096 ? blocks : 4096), 256, 0, st>>>(h->d_table, h->phase, h->inc, n, (const uint32_t*) d_in_i16, (float2*) d_out_c64);
    if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
    return advance(h, n);
}

} // extern "C"
