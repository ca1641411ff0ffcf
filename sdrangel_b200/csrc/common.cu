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

namespace {
// each block notes the SM it runs on and lingers a few microseconds so the blocks of the probe are co-resident
__global__ void sm_probe_kernel(int* smids)
{
    extern __shared__ unsigned char probe_smem[];
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) { probe_smem[0] = 1; smids[blockIdx.x] = (int) smid; }
    const long long t0 = clock64();
    while (clock64() - t0 < 40000) { }
}
} // namespace

extern "C" {

// Which SMs does the block scheduler hand to the first `n_blocks` big blocks of a kernel launched on an idle GPU?  (A
// collective's CTAs launched into a gap land there; b200dsp_bank_set_reserved_sms keeps the bank's kernels off them.)
int b200dsp_probe_sm_order(int n_blocks, int threads_per_block, int* smids)
{
    int rc = b200_require_device();
    if (rc) return rc;
    if (n_blocks < 1 || n_blocks > 1024 || threads_per_block < 32 || threads_per_block > 1024 || !smids) return b200_fail(B200DSP_EINVAL, "probe_sm_order: bad argument");
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(b200_current_device())))) return rc;
    int* d = nullptr;
    const size_t smem = 120 * 1024;                       // more than half an SM's shared memory: one block per SM
    if ((rc = B200_CUDA_CHECK(cudaMalloc(&d, n_blocks * sizeof(int)))) ||
        (rc = B200_CUDA_CHECK(cudaFuncSetAttribute((const void*) sm_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem))) ||
        (rc = B200_CUDA_CHECK(cudaDeviceSynchronize()))) { if (d) cudaFree(d); return rc; }
    sm_probe_kernel<<<n_blocks, threads_per_block, smem>>>(d);
    if ((rc = B200_CUDA_CHECK(cudaGetLastError())) || (rc = B200_CUDA_CHECK(cudaDeviceSynchronize())) ||
        (rc = B200_CUDA_CHECK(cudaMemcpy(smids, d, n_blocks * sizeof(int), cudaMemcpyDeviceToHost)))) { cudaFree(d); return rc; }
    cudaFree(d);
    return 0;
}

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
