// fifo_api.cu — device-resident SampleSinkFifo: b200dsp_fifo_*
//
// Replaces (paths relative to the reference tree):
//   SampleSinkFifo::write / read / readBegin / readCommit      sdrbase/dsp/samplesinkfifo.cpp:113-231
// The reference's FIFO is the hand-over point between a device plugin's decimators (producer thread) and the engine's
// work loop (consumer: DSPDeviceSourceEngine::work, dspdevicesourceengine.cpp:325-408).  Same ring semantics -- overflow
// drops the excess, a read hands out at most two contiguous spans and is committed separately -- but the ring lives in
// device memory: the decimator kernels' output goes in with a device-to-device copy and the bank is fed from the spans
// with b200dsp_bank_feed_dev, so the samples never visit the host.  Head/tail/fill are host-side (one mutex), every copy is
// ordered on the caller's stream.
#include "common.cuh"

using namespace b200dsp;

struct b200dsp_fifo {
    int device = 0;
    uint32_t* d_data = nullptr;
    uint32_t size = 0, fill = 0, head = 0, tail = 0;
    std::mutex mu;
};

extern "C" {

int b200dsp_fifo_create(b200dsp_fifo_t** out, uint32_t size_samples)
{
    if (!out) return b200_fail(B200DSP_EINVAL, "fifo_create: null handle pointer");
    *out = nullptr;
    if (size_samples == 0 || size_samples > (1u << 30)) return b200_fail(B200DSP_EINVAL, "fifo_create: bad size");
    int rc = b200_require_device();
    if (rc) return rc;
    b200dsp_fifo* f = new (std::nothrow) b200dsp_fifo();
    if (!f) return b200_fail(B200DSP_ENOMEM, "fifo_create: out of host memory");
    f->device = b200_current_device();
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(f->device))) || (rc = B200_CUDA_CHECK(cudaMalloc(&f->d_data, (size_t) size_samples * 4)))) { delete f; return rc; }
    f->size = size_samples;
    *out = f;
    return 0;
}

int b200dsp_fifo_destroy(b200dsp_fifo_t* f)
{
    if (!f) return 0;
    cudaSetDevice(f->device);
    if (f->d_data) cudaFree(f->d_data);
    delete f;
    return 0;
}

uint32_t b200dsp_fifo_size(b200dsp_fifo_t* f) { return f ? f->size : 0; }
uint32_t b200dsp_fifo_fill(b200dsp_fifo_t* f)
{
    if (!f) return 0;
    std::lock_guard<std::mutex> g(f->mu);
    return f->fill;
}

// == SampleSinkFifo::write: min(count, size - fill) samples go in, the rest is dropped (overflow); returns the number written
int b200dsp_fifo_write(b200dsp_fifo_t* f, const void* samples, uint32_t count, int src_is_device, void* cuda_stream, uint32_t* written)
{
    if (!f || (count > 0 && !samples)) return b200_fail(B200DSP_EINVAL, "fifo_write: bad argument");
    int rc = B200_CUDA_CHECK(cudaSetDevice(f->device));
    if (rc) return rc;
    std::lock_guard<std::mutex> g(f->mu);
    const uint32_t room = f->size - f->fill;
    const uint32_t total = count < room ? count : room;
    uint32_t remaining = total;
    const uint32_t* src = (const uint32_t*) samples;
    const cudaMemcpyKind kind = src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    while (remaining > 0) {
        const uint32_t len = remaining < f->size - f->tail ? remaining : f->size - f->tail;
        if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(f->d_data + f->tail, src, (size_t) len * 4, kind, (cudaStream_t) cuda_stream)))) return rc;
        f->tail = (f->tail + len) % f->size;
        f->fill += len;
        src += len;
        remaining -= len;
    }
    if (written) *written = total;
    return 0;
}

// == SampleSinkFifo::readBegin: up to `count` samples as two contiguous device spans; nothing is consumed until _read_commit
int b200dsp_fifo_read_begin(b200dsp_fifo_t* f, uint32_t count, const void** part1, uint32_t* n1, const void** part2, uint32_t* n2, uint32_t* total_out)
{
    if (!f || !part1 || !n1 || !part2 || !n2) return b200_fail(B200DSP_EINVAL, "fifo_read_begin: null argument");
    std::lock_guard<std::mutex> g(f->mu);
    const uint32_t total = count < f->fill ? count : f->fill;       // (underflow: what there is)
    uint32_t remaining = total, head = f->head;
    *part1 = nullptr; *n1 = 0; *part2 = nullptr; *n2 = 0;
    if (remaining > 0) {
        const uint32_t len = remaining < f->size - head ? remaining : f->size - head;
        *part1 = f->d_data + head; *n1 = len;
        head = (head + len) % f->size;
        remaining -= len;
    }
    if (remaining > 0) {
        const uint32_t len = remaining < f->size - head ? remaining : f->size - head;
        *part2 = f->d_data + head; *n2 = len;
    }
    if (total_out) *total_out = total;
    return 0;
}

// == SampleSinkFifo::readCommit (at most `fill` samples); the caller orders later writes after its reads of the spans
int b200dsp_fifo_read_commit(b200dsp_fifo_t* f, uint32_t count, uint32_t* committed)
{
    if (!f) return b200_fail(B200DSP_EINVAL, "null handle");
    std::lock_guard<std::mutex> g(f->mu);
    if (count > f->fill) count = f->fill;
    f->head = (f->head + count) % f->size;
    f->fill -= count;
    if (committed) *committed = count;
    return 0;
}

// == SampleSinkFifo::read into host memory (copy + commit), synchronous on the stream
int b200dsp_fifo_read(b200dsp_fifo_t* f, void* out_host, uint32_t count, void* cuda_stream, uint32_t* read)
{
    if (!f || (count > 0 && !out_host)) return b200_fail(B200DSP_EINVAL, "fifo_read: bad argument");
    int rc = B200_CUDA_CHECK(cudaSetDevice(f->device));
    if (rc) return rc;
    const void *p1, *p2;
    uint32_t n1, n2, total;
    if ((rc = b200dsp_fifo_read_begin(f, count, &p1, &n1, &p2, &n2, &total))) return rc;
    if (n1 && (rc = B200_CUDA_CHECK(cudaMemcpyAsync(out_host, p1, (size_t) n1 * 4, cudaMemcpyDeviceToHost, (cudaStream_t) cuda_stream)))) return rc;
    if (n2 && (rc = B200_CUDA_CHECK(cudaMemcpyAsync((uint32_t*) out_host + n1, p2, (size_t) n2 * 4, cudaMemcpyDeviceToHost, (cudaStream_t) cuda_stream)))) return rc;
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t) cuda_stream)))) return rc;
    if ((rc = b200dsp_fifo_read_commit(f, total, nullptr))) return rc;
    if (read) *read = total;
    return 0;
}

} // extern "C"
