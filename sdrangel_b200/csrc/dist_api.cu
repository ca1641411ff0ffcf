// dist_api.cu — K6: the wideband baseband of a channel bank sharded over the GPUs of one box: b200dsp_dist_*
//
// The reference has no counterpart (one process, one device stream, every channel on its own host thread:
// sdrbase/dsp/dspdevicesourceengine.cpp:325-408 hands the same SampleVector span to every channel sink).  Here the channels
// are sharded by frequency block over N GPUs (one process or thread per GPU) and every GPU needs the whole baseband:
//   * device-resident source (the metric path): NCCL broadcast from the ingest GPU over NVLink;
//   * host-fed source: every rank copies ITS 1/N time slice over its own PCIe link and an in-place NCCL all-gather
//     completes the block on every GPU -- the ingest rate is N links instead of one.
// Two receive slots: the transfer of block k+1 runs on the collective stream under the kernels of block k.
// NCCL is loaded at run time (dlopen of libnccl.so.2: the copy the host process already uses, e.g. PyTorch's), so a
// single-GPU user of libb200dsp needs no NCCL at all.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <cuda.h>
#include <stdlib.h>
#include <vector>

using namespace b200dsp;

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl()
{
    static NcclApi api;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    if (api.lib) return api;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
    if (!api.lib) return api;
    api.GetUniqueId = (decltype(api.GetUniqueId)) dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank)) dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy)) dlsym(api.lib, "ncclCommDestroy");
    api.Broadcast = (decltype(api.Broadcast)) dlsym(api.lib, "ncclBroadcast");
    api.AllGather = (decltype(api.AllGather)) dlsym(api.lib, "ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString)) dlsym(api.lib, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.AllGather && api.GetErrorString;
    return api;
}

int nccl_check(ncclResult_t r, const char* what)
{
    if (r == ncclSuccess) return 0;
    return b200_fail(B200DSP_ECUDA, "NCCL: %s (%s)", nccl().GetErrorString ? nccl().GetErrorString(r) : "error", what);
}

int require_nccl()
{
    if (!nccl().ok) return b200_fail(B200DSP_ESTATE, "libnccl.so.2 not found (needed only by b200dsp_dist_*)");
    return 0;
}

} // namespace

constexpr int P2P_SLOTS = 3;
constexpr int P2P_SUB = 16;               // sub-blocks per block
struct P2pBlob { cudaIpcMemHandle_t slot[P2P_SLOTS]; cudaIpcMemHandle_t flags; };
static_assert(sizeof(P2pBlob) <= B200DSP_DIST_P2P_BLOB_BYTES, "blob size");

typedef CUresult (*StreamMemOp)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
struct DrvApi { StreamMemOp wait32 = nullptr, write32 = nullptr; bool ok = false; };
static DrvApi& drv()
{
    static DrvApi a;
    if (a.ok) return a;
    cudaDriverEntryPointQueryResult st;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &st) == cudaSuccess && f) a.wait32 = (StreamMemOp) f;
    f = nullptr;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &st) == cudaSuccess && f) a.write32 = (StreamMemOp) f;
    cudaGetLastError();
    a.ok = a.wait32 && a.write32;
    return a;
}

struct b200dsp_dist {
    int device = 0, rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t coll = nullptr, h2d = nullptr;
    uint32_t* slot[2] = { nullptr, nullptr };
    long long cap = 0;
    long long n[2] = { 0, 0 };
    cudaEvent_t ev_ready[2] = { nullptr, nullptr }, ev_free[2] = { nullptr, nullptr }, ev_src = nullptr, ev_h2d = nullptr;
    bool fed[2] = { false, false };
    // copy-engine chain (b200dsp_dist_p2p_*): the block travels rank 0 -> 1 -> ... -> N-1 in sub-blocks, forwarded by DMA
    // copies into the next rank's slot (CUDA IPC mapping), ordered by counters in device memory that the streams wait on and
    // write (stream memory operations): no SM is taken from the FIR kernels and every link carries the block once.
    bool p2p = false;
    long long p2p_cap = 0;
    uint32_t* pslot[P2P_SLOTS] = { nullptr, nullptr, nullptr };       // my receive slots
    uint32_t* nslot[P2P_SLOTS] = { nullptr, nullptr, nullptr };       // the next rank's (IPC)
    uint32_t* flags = nullptr;            // mine: [0..2] sub-blocks arrived per slot, [4..6] uses of the NEXT rank's slot it has finished
    uint32_t* nflags = nullptr;           // the next rank's flags (IPC): I write its arrived[]
    uint32_t* pflags = nullptr;           // the previous rank's flags (IPC): I write its "next freed"[]
    uint32_t* seqtab = nullptr;           // device table of the values 0..P2P_SEQ-1 (source of the 4-byte flag copies)
    cudaStream_t fwd = nullptr, fwd2 = nullptr, sig = nullptr;      // fwd2: second half of every sub-block on another copy engine
    cudaEvent_t ev_half = nullptr, ev_go = nullptr;
    cudaEvent_t ev_fwd[P2P_SLOTS] = { nullptr, nullptr, nullptr }, ev_cons[P2P_SLOTS] = { nullptr, nullptr, nullptr };
    unsigned uses[P2P_SLOTS] = { 0, 0, 0 };
    long long pn[P2P_SLOTS] = { 0, 0, 0 };
    const uint32_t* psrc[P2P_SLOTS] = { nullptr, nullptr, nullptr };   // root: the caller's source buffer stands in for the slot
};

extern "C" {

int b200dsp_dist_shard(int n_channels, int world, int rank, int* lo, int* hi)
{
    if (n_channels < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return b200_fail(B200DSP_EINVAL, "dist_shard: bad argument");
    // contiguous blocks in the order the channels were given (frequency order): the first (n mod world) ranks take one more
    const int q = n_channels / world, r = n_channels % world;
    *lo = rank * q + (rank < r ? rank : r);
    *hi = *lo + q + (rank < r ? 1 : 0);
    return 0;
}

int b200dsp_dist_unique_id(void* id_out)
{
    if (!id_out) return b200_fail(B200DSP_EINVAL, "dist_unique_id: null buffer");
    int rc = require_nccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == B200DSP_DIST_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if ((rc = nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId"))) return rc;
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

int b200dsp_dist_create(b200dsp_dist_t** out, const void* id, int rank, int world)
{
    if (!out) return b200_fail(B200DSP_EINVAL, "dist_create: null handle pointer");
    *out = nullptr;
    if (!id || world < 1 || rank < 0 || rank >= world) return b200_fail(B200DSP_EINVAL, "dist_create: bad argument");
    int rc = b200_require_device();
    if (rc) return rc;
    if ((rc = require_nccl())) return rc;
    b200dsp_dist* d = new (std::nothrow) b200dsp_dist();
    if (!d) return b200_fail(B200DSP_ENOMEM, "dist_create: out of host memory");
    d->device = b200_current_device(); d->rank = rank; d->world = world;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(d->device))) ||
        (rc = B200_CUDA_CHECK(cudaStreamCreateWithPriority(&d->coll, cudaStreamNonBlocking, -5))) ||
        (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&d->h2d, cudaStreamNonBlocking))) ||
        (rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&d->ev_src, cudaEventDisableTiming))) ||
        (rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&d->ev_h2d, cudaEventDisableTiming)))) { b200dsp_dist_destroy(d); return rc; }
    for (int i = 0; i < 2; ++i)
        if ((rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&d->ev_ready[i], cudaEventDisableTiming))) ||
            (rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&d->ev_free[i], cudaEventDisableTiming)))) { b200dsp_dist_destroy(d); return rc; }
    if ((rc = nccl_check(nccl().CommInitRank(&d->comm, world, uid, rank), "ncclCommInitRank"))) { b200dsp_dist_destroy(d); return rc; }
    *out = d;
    return 0;
}

int b200dsp_dist_destroy(b200dsp_dist_t* d)
{
    if (!d) return 0;
    cudaSetDevice(d->device);
    if (d->coll) cudaStreamSynchronize(d->coll);
    if (d->h2d) cudaStreamSynchronize(d->h2d);
    if (d->comm && nccl().ok) nccl().CommDestroy(d->comm);
    for (int i = 0; i < 2; ++i) {
        if (d->slot[i]) cudaFree(d->slot[i]);
        if (d->ev_ready[i]) cudaEventDestroy(d->ev_ready[i]);
        if (d->ev_free[i]) cudaEventDestroy(d->ev_free[i]);
    }
    if (d->fwd) cudaStreamSynchronize(d->fwd);
    if (d->fwd2) { cudaStreamSynchronize(d->fwd2); cudaStreamDestroy(d->fwd2); }
    if (d->ev_half) cudaEventDestroy(d->ev_half);
    if (d->ev_go) cudaEventDestroy(d->ev_go);
    if (d->sig) cudaStreamSynchronize(d->sig);
    for (int i = 0; i < P2P_SLOTS; ++i) {
        if (d->nslot[i]) cudaIpcCloseMemHandle(d->nslot[i]);
        if (d->pslot[i]) cudaFree(d->pslot[i]);
        if (d->ev_fwd[i]) cudaEventDestroy(d->ev_fwd[i]);
        if (d->ev_cons[i]) cudaEventDestroy(d->ev_cons[i]);
    }
    if (d->nflags) cudaIpcCloseMemHandle(d->nflags);
    if (d->pflags && d->pflags != d->nflags) cudaIpcCloseMemHandle(d->pflags);
    if (d->flags) cudaFree(d->flags);
    if (d->seqtab) cudaFree(d->seqtab);
    if (d->fwd) cudaStreamDestroy(d->fwd);
    if (d->sig) cudaStreamDestroy(d->sig);
    if (d->ev_src) cudaEventDestroy(d->ev_src);
    if (d->ev_h2d) cudaEventDestroy(d->ev_h2d);
    if (d->coll) cudaStreamDestroy(d->coll);
    if (d->h2d) cudaStreamDestroy(d->h2d);
    delete d;
    return 0;
}

int b200dsp_dist_reserve(b200dsp_dist_t* d, int64_t n_samples)
{
    if (!d || n_samples < 0) return b200_fail(B200DSP_EINVAL, "dist_reserve: bad argument");
    if (d->cap >= n_samples) return 0;
    int rc = B200_CUDA_CHECK(cudaSetDevice(d->device));
    if (rc) return rc;
    if ((rc = B200_CUDA_CHECK(cudaDeviceSynchronize()))) return rc;
    for (int i = 0; i < 2; ++i) {
        if (d->slot[i]) cudaFree(d->slot[i]);
        d->slot[i] = nullptr; d->fed[i] = false; d->n[i] = 0;
    }
    d->cap = 0;
    for (int i = 0; i < 2; ++i)
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&d->slot[i], (size_t) (n_samples + 8) * 4)))) return rc;
    d->cap = n_samples;
    return 0;
}

int b200dsp_dist_bcast_begin(b200dsp_dist_t* d, int slot, const void* d_iq, int64_t n_samples, int root, void* after_stream)
{
    if (!d || slot < 0 || slot > 1 || n_samples <= 0 || root < 0 || root >= d->world) return b200_fail(B200DSP_EINVAL, "dist_bcast_begin: bad argument");
    if (d->rank == root && !d_iq) return b200_fail(B200DSP_EINVAL, "dist_bcast_begin: the root needs a source buffer");
    int rc = B200_CUDA_CHECK(cudaSetDevice(d->device));
    if (rc) return rc;
    if ((rc = b200dsp_dist_reserve(d, n_samples))) return rc;
    if (d->fed[slot] && (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(d->coll, d->ev_free[slot], 0)))) return rc;     // its last consumer has run
    if (d->rank == root && after_stream) {
        if ((rc = B200_CUDA_CHECK(cudaEventRecord(d->ev_src, (cudaStream_t) after_stream))) ||
            (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(d->coll, d->ev_src, 0)))) return rc;
    }
    // one IQ sample = one 32-bit word (NCCL has no int16 pair type; a broadcast moves bytes)
    const void* src = (d->rank == root) ? d_iq : (const void*) d->slot[slot];
    if ((rc = nccl_check(nccl().Broadcast(src, d->slot[slot], (size_t) n_samples, ncclInt32, root, d->comm, d->coll), "ncclBroadcast"))) return rc;
    d->n[slot] = n_samples; d->fed[slot] = false;
    return B200_CUDA_CHECK(cudaEventRecord(d->ev_ready[slot], d->coll));
}

int b200dsp_dist_ingest_begin(b200dsp_dist_t* d, int slot, const int16_t* host_slice, int64_t n_samples_total)
{
    if (!d || slot < 0 || slot > 1 || !host_slice || n_samples_total <= 0) return b200_fail(B200DSP_EINVAL, "dist_ingest_begin: bad argument");
    if (n_samples_total % ((int64_t) d->world * 4)) return b200_fail(B200DSP_EINVAL, "dist_ingest_begin: the block must split into %d slices of whole 16-byte words", d->world);
    int rc = B200_CUDA_CHECK(cudaSetDevice(d->device));
    if (rc) return rc;
    if ((rc = b200dsp_dist_reserve(d, n_samples_total))) return rc;
    const long long cnt = n_samples_total / d->world;
    uint32_t* mine = d->slot[slot] + (long long) d->rank * cnt;
    if (d->fed[slot] && (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(d->h2d, d->ev_free[slot], 0)))) return rc;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(mine, host_slice, (size_t) cnt * 4, cudaMemcpyHostToDevice, d->h2d))) ||
        (rc = B200_CUDA_CHECK(cudaEventRecord(d->ev_h2d, d->h2d))) ||
        (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(d->coll, d->ev_h2d, 0)))) return rc;
    if (d->world > 1 && (rc = nccl_check(nccl().AllGather(mine, d->slot[slot], (size_t) cnt, ncclInt32, d->comm, d->coll), "ncclAllGather"))) return rc;
    d->n[slot] = n_samples_total; d->fed[slot] = false;
    return B200_CUDA_CHECK(cudaEventRecord(d->ev_ready[slot], d->coll));
}

int b200dsp_dist_feed(b200dsp_dist_t* d, int slot, b200dsp_bank_t* bank, void* stream)
{
    if (!d || !bank || slot < 0 || slot > 1 || d->n[slot] <= 0) return b200_fail(B200DSP_EINVAL, "dist_feed: bad argument or empty slot");
    int rc = B200_CUDA_CHECK(cudaSetDevice(d->device));
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t) stream : (cudaStream_t) b200dsp_bank_stream(bank);
    if ((rc = B200_CUDA_CHECK(cudaStreamWaitEvent(st, d->ev_ready[slot], 0)))) return rc;
    if ((rc = b200dsp_bank_feed_dev(bank, d->slot[slot], d->n[slot], (void*) st))) return rc;
    d->fed[slot] = true;
    return B200_CUDA_CHECK(cudaEventRecord(d->ev_free[slot], st));
}

// ---- copy-engine chain ------------------------------------------------------------------------------------------------
int b200dsp_dist_p2p_export(b200dsp_dist_t* d, int64_t n_samples, void* blob_out)
{
    if (!d || n_samples <= 0 || !blob_out) return b200_fail(B200DSP_EINVAL, "dist_p2p_export: bad argument");
    if (!drv().ok) return b200_fail(B200DSP_ESTATE, "dist_p2p_export: the driver has no stream memory operations");
    int rc = B200_CUDA_CHECK(cudaSetDevice(d->device));
    if (rc) return rc;
    if (d->pslot[0]) return b200_fail(B200DSP_ESTATE, "dist_p2p_export: already exported");
    P2pBlob blob;
    memset(&blob, 0, sizeof(blob));
    const size_t bytes = (size_t) (n_samples + 8) * 4;
    for (int i = 0; i < P2P_SLOTS; ++i)
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&d->pslot[i], bytes))) || (rc = B200_CUDA_CHECK(cudaIpcGetMemHandle(&blob.slot[i], d->pslot[i])))) return rc;
    const int nseq = 1 << 20;
    std::vector<uint32_t> seq(nseq);
    for (int i = 0; i < nseq; ++i) seq[i] = (uint32_t) i;
    if ((rc = B200_CUDA_CHECK(cudaMalloc(&d->flags, 64))) || (rc = B200_CUDA_CHECK(cudaMemset(d->flags, 0, 64))) ||
        (rc = B200_CUDA_CHECK(cudaIpcGetMemHandle(&blob.flags, d->flags))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&d->seqtab, nseq * 4))) || (rc = B200_CUDA_CHECK(cudaMemcpy(d->seqtab, seq.data(), nseq * 4, cudaMemcpyHostToDevice))) ||
        (rc = B200_CUDA_CHECK(cudaStreamCreateWithPriority(&d->fwd, cudaStreamNonBlocking, -5))) ||
        (rc = B200_CUDA_CHECK(cudaStreamCreateWithPriority(&d->sig, cudaStreamNonBlocking, -5))) ||
        (rc = B200_CUDA_CHECK(cudaStreamCreateWithPriority(&d->fwd2, cudaStreamNonBlocking, -5))) ||
        (rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&d->ev_half, cudaEventDisableTiming))) ||
        (rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&d->ev_go, cudaEventDisableTiming)))) return rc;
    for (int i = 0; i < P2P_SLOTS; ++i)
        if ((rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&d->ev_fwd[i], cudaEventDisableTiming))) ||
            (rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&d->ev_cons[i], cudaEventDisableTiming)))) return rc;
    if ((rc = B200_CUDA_CHECK(cudaDeviceSynchronize()))) return rc;
    d->p2p_cap = n_samples;
    memset(blob_out, 0, B200DSP_DIST_P2P_BLOB_BYTES);
    memcpy(blob_out, &blob, sizeof(blob));
    return 0;
}

int b200dsp_dist_p2p_import(b200dsp_dist_t* d, const void* blobs_all)
{
    if (!d || !blobs_all || !d->pslot[0]) return b200_fail(B200DSP_EINVAL, "dist_p2p_import: export first");
    int rc = B200_CUDA_CHECK(cudaSetDevice(d->device));
    if (rc) return rc;
    const char* base = (const char*) blobs_all;
    if (d->rank + 1 < d->world) {
        P2pBlob nb;
        memcpy(&nb, base + (size_t) (d->rank + 1) * B200DSP_DIST_P2P_BLOB_BYTES, sizeof(nb));
        for (int i = 0; i < P2P_SLOTS; ++i)
            if ((rc = B200_CUDA_CHECK(cudaIpcOpenMemHandle((void**) &d->nslot[i], nb.slot[i], cudaIpcMemLazyEnablePeerAccess)))) return rc;
        if ((rc = B200_CUDA_CHECK(cudaIpcOpenMemHandle((void**) &d->nflags, nb.flags, cudaIpcMemLazyEnablePeerAccess)))) return rc;
    }
    if (d->rank > 0) {
        P2pBlob pb;
        memcpy(&pb, base + (size_t) (d->rank - 1) * B200DSP_DIST_P2P_BLOB_BYTES, sizeof(pb));
        if ((rc = B200_CUDA_CHECK(cudaIpcOpenMemHandle((void**) &d->pflags, pb.flags, cudaIpcMemLazyEnablePeerAccess)))) return rc;
    }
    d->p2p = true;
    return 0;
}

namespace {
int drv_check(CUresult r, const char* what)
{
    if (r == CUDA_SUCCESS) return 0;
    return b200_fail(B200DSP_ECUDA, "driver error %d in %s", (int) r, what);
}
}

// collective, chain form of b200dsp_dist_bcast_begin (root must be rank 0): slot in [0, 3)
int b200dsp_dist_p2p_begin(b200dsp_dist_t* d, int slot, const void* d_iq, int64_t n_samples, void* after_stream)
{
    if (!d || !d->p2p || slot < 0 || slot >= P2P_SLOTS || n_samples <= 0 || n_samples > d->p2p_cap || (n_samples % (4 * P2P_SUB)))
        return b200_fail(B200DSP_EINVAL, "dist_p2p_begin: bad argument (block must be a multiple of %d samples, at most the exported size)", 4 * P2P_SUB);
    if (d->rank == 0 && !d_iq) return b200_fail(B200DSP_EINVAL, "dist_p2p_begin: rank 0 needs the source buffer");
    int rc = B200_CUDA_CHECK(cudaSetDevice(d->device));
    if (rc) return rc;
    const unsigned use = d->uses[slot]++;                       // the same on every rank: the calls are collective
    const unsigned base = use * P2P_SUB;
    if (base + P2P_SUB >= (1u << 20)) return b200_fail(B200DSP_ESTATE, "dist_p2p_begin: sequence table exhausted (%u uses of a slot)", use);
    const long long sub = n_samples / P2P_SUB;
    const uint32_t* mine = (d->rank == 0) ? (const uint32_t*) d_iq : d->pslot[slot];
    d->psrc[slot] = mine; d->pn[slot] = n_samples;
    if (d->rank == 0 && after_stream) {
        if ((rc = B200_CUDA_CHECK(cudaEventRecord(d->ev_src, (cudaStream_t) after_stream))) ||
            (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(d->fwd, d->ev_src, 0)))) return rc;
    }
    if (d->rank + 1 < d->world) {
        // the next rank's slot must have finished its previous use (its consumer and its own forwards): it says so in my flags
        if (use > 0 && (rc = drv_check(drv().wait32((CUstream) d->fwd, (CUdeviceptr) (d->flags + 4 + slot), use, CU_STREAM_WAIT_VALUE_GEQ), "cuStreamWaitValue32"))) return rc;
        // B200DSP_P2P_STRIPE=1: a sub-block goes out as two halves on two streams (two copy engines); measured slower on the
        // 8-GPU box (r02: 0.67 vs 0.56 ms per 201 MB block), so one copy per sub-block by default
        static const bool stripe = (getenv("B200DSP_P2P_STRIPE") && getenv("B200DSP_P2P_STRIPE")[0] == '1');
        const long long h1 = stripe ? ((sub / 2) & ~3ll) : sub;
        for (int j = 0; j < P2P_SUB; ++j) {
            if (d->rank > 0 && (rc = drv_check(drv().wait32((CUstream) d->fwd, (CUdeviceptr) (d->flags + slot), base + j + 1, CU_STREAM_WAIT_VALUE_GEQ), "cuStreamWaitValue32"))) return rc;
            if (h1 < sub) {
                if ((rc = B200_CUDA_CHECK(cudaEventRecord(d->ev_go, d->fwd))) || (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(d->fwd2, d->ev_go, 0))) ||
                    (rc = B200_CUDA_CHECK(cudaMemcpyAsync(d->nslot[slot] + j * sub + h1, mine + j * sub + h1, (size_t) (sub - h1) * 4, cudaMemcpyDeviceToDevice, d->fwd2))) ||
                    (rc = B200_CUDA_CHECK(cudaEventRecord(d->ev_half, d->fwd2)))) return rc;
            }
            if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(d->nslot[slot] + j * sub, mine + j * sub, (size_t) h1 * 4, cudaMemcpyDeviceToDevice, d->fwd)))) return rc;
            if (h1 < sub && (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(d->fwd, d->ev_half, 0)))) return rc;
            if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(d->nflags + slot, d->seqtab + base + j + 1, 4, cudaMemcpyDeviceToDevice, d->fwd)))) return rc;
        }
    }
    return B200_CUDA_CHECK(cudaEventRecord(d->ev_fwd[slot], d->fwd));
}

int b200dsp_dist_p2p_feed(b200dsp_dist_t* d, int slot, b200dsp_bank_t* bank, void* stream)
{
    // (bank == NULL with a stream: only wait for the block and release the slot -- transfer-rate measurements)
    if (!d || !d->p2p || (!bank && !stream) || slot < 0 || slot >= P2P_SLOTS || d->pn[slot] <= 0 || d->uses[slot] == 0) return b200_fail(B200DSP_EINVAL, "dist_p2p_feed: bad argument or empty slot");
    int rc = B200_CUDA_CHECK(cudaSetDevice(d->device));
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t) stream : (cudaStream_t) b200dsp_bank_stream(bank);
    const unsigned use = d->uses[slot] - 1;
    if (d->rank > 0 && (rc = drv_check(drv().wait32((CUstream) st, (CUdeviceptr) (d->flags + slot), use * P2P_SUB + P2P_SUB, CU_STREAM_WAIT_VALUE_GEQ), "cuStreamWaitValue32"))) return rc;
    if (bank && (rc = b200dsp_bank_feed_dev(bank, d->psrc[slot], d->pn[slot], (void*) st))) return rc;
    if ((rc = B200_CUDA_CHECK(cudaEventRecord(d->ev_cons[slot], st)))) return rc;
    if (d->rank > 0) {
        // this slot is free for its next use once the feed and the forwards have run: tell the previous rank
        if ((rc = B200_CUDA_CHECK(cudaStreamWaitEvent(d->sig, d->ev_cons[slot], 0))) || (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(d->sig, d->ev_fwd[slot], 0))) ||
            (rc = B200_CUDA_CHECK(cudaMemcpyAsync(d->pflags + 4 + slot, d->seqtab + use + 1, 4, cudaMemcpyDeviceToDevice, d->sig)))) return rc;
    }
    return 0;
}

int b200dsp_dist_slot(b200dsp_dist_t* d, int slot, const void** d_ptr, int64_t* n_samples)
{
    if (!d || slot < 0 || slot > 1 || !d_ptr) return b200_fail(B200DSP_EINVAL, "dist_slot: bad argument");
    *d_ptr = d->slot[slot];
    if (n_samples) *n_samples = d->n[slot];
    return 0;
}

int b200dsp_dist_p2p_slot(b200dsp_dist_t* d, int slot, const void** d_ptr, int64_t* n_samples)
{
    if (!d || slot < 0 || slot >= P2P_SLOTS || !d_ptr) return b200_fail(B200DSP_EINVAL, "dist_p2p_slot: bad argument");
    *d_ptr = d->psrc[slot];
    if (n_samples) *n_samples = d->pn[slot];
    return 0;
}

int b200dsp_dist_sync(b200dsp_dist_t* d)
{
    if (!d) return b200_fail(B200DSP_EINVAL, "null handle");
    int rc = B200_CUDA_CHECK(cudaSetDevice(d->device));
    if (rc) return rc;
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(d->h2d)))) return rc;
    if (d->fwd && (rc = B200_CUDA_CHECK(cudaStreamSynchronize(d->fwd)))) return rc;
    if (d->sig && (rc = B200_CUDA_CHECK(cudaStreamSynchronize(d->sig)))) return rc;
    return B200_CUDA_CHECK(cudaStreamSynchronize(d->coll));
}

} // extern "C"
