// common.cu — library-level entry points of the C ABI (include/b200dsp.h) and shared plumbing.
#include "common.cuh"
#include <stdarg.h>

namespace b200dsp {

static thread_local char g_err[512] = "";
static int g_device = 0;
static std::mutex g_mu;

int b200_fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int b200_cuda_check(cudaError_t e, const char* what, const char* file, int line)
{
    if (e == cudaSuccess) return 0;
    const char* base = strrchr(file, '/');
    b200_fail(B200DSP_ECUDA, "%s: %s (%s:%d)", cudaGetErrorString(e), what, base ? base + 1 : file, line);
    cudaGetLastError();
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return B200DSP_ENODEV;
    if (e == cudaErrorMemoryAllocation) return B200DSP_ENOMEM;
    return B200DSP_ECUDA;
}

int b200_require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return b200_fail(B200DSP_ENODEV, "no CUDA device available (%s); libb200dsp has no CPU fallback",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return 0;
}

int b200_current_device() { std::lock_guard<std::mutex> g(g_mu); return g_device; }

int b200_sm_count_of(int device)
{
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
    return n;
}

} // namespace b200dsp

using namespace b200dsp;

extern "C" {

int b200dsp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int b200dsp_init(int device_ordinal)
{
    int rc = b200_require_device();
    if (rc) return rc;
    if (device_ordinal < 0 || device_ordinal >= b200dsp_device_count()) return b200_fail(B200DSP_EINVAL, "init: device %d out of range", device_ordinal);
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(device_ordinal)))) return rc;
    if ((rc = B200_CUDA_CHECK(cudaFree(0)))) return rc;
    std::lock_guard<std::mutex> g(g_mu);
    g_device = device_ordinal;
    return 0;
}

const char* b200dsp_last_error(void) { return g_err; }
const char* b200dsp_version(void) { return "b200dsp 0.1 (sm_100a)"; }
int b200dsp_sm_count(void) { return b200dsp_device_count() > 0 ? b200_sm_count_of(b200_current_device()) : 0; }

} // extern "C"
