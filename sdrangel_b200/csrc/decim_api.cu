// decim_api.cu — C ABI for the half-band decimation cascades (K1/K2) : b200dsp_decim_*
// Host-side mirror of Decimators<> / DecimatorsFI / DecimatorsFF / DecimatorsIF call semantics: block rounding
// ("trailing partial block dropped", decimators.h:2862), decimation_shifts<> tables (decimators.h:79-167),
// stage rotation plan of the _inf/_sup entry points (decimators.h:463-2584), carried filter state.
#include "common.cuh"
#include "hb64_cascade.cuh"

using namespace b200dsp;

namespace {

// ---------------------------------------------------------------------------------------------------------
// element-wise entry points without a half-band stage
//   EW_COPY : decimate1                    (decimators.h:344-356, decimatorsfi.cpp:19-31, decimatorsif.h:86-98)
//   EW_HALF : float decimate2_inf/_sup     (decimatorsfi.cpp:55-93)   8 scalars -> 2 outputs, unfiltered
//   EW_DIV4 : float decimate4_inf/_sup     (decimatorsfi.cpp:95-131)  8 scalars -> 1 output, unfiltered
// ---------------------------------------------------------------------------------------------------------
enum { EW_COPY = 0, EW_HALF = 1, EW_DIV4 = 2, EW_HALF_U = 3 };

struct EwParams {
    const void* in; void* out;
    long long n_units;     // COPY: IQ samples, HALF/DIV4: groups of 8 scalars
    int kind, sup, pre, shift;
    float out_scale;
};

template<typename TIN> __device__ __forceinline__ float ew_ld(const TIN* p, long long i) { return (float) p[i]; }

template<bool IN_I16, int OUT>
__device__ __forceinline__ void ew_store(const EwParams& p, long long k, float re, float im)
{
    if (OUT == OUT_F32) {
        reinterpret_cast<float2*>(p.out)[k] = make_float2(re * p.out_scale, im * p.out_scale);
    } else {
        const int32_t a = cvt_trunc_x86(re * 32768.0f), b = cvt_trunc_x86(im * 32768.0f);
        reinterpret_cast<uint32_t*>(p.out)[k] = ((uint32_t) a & 0xffffu) | ((uint32_t) b << 16);
    }
}

template<typename TIN, int OUT>
__global__ void ew_float_kernel(const EwParams p)
{
    const TIN* in = reinterpret_cast<const TIN*>(p.in);
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long u = (long long) blockIdx.x * blockDim.x + threadIdx.x; u < p.n_units; u += stride) {
        if (p.kind == EW_COPY) {
            ew_store<sizeof(TIN) == 2, OUT>(p, u, ew_ld(in, 2 * u), ew_ld(in, 2 * u + 1));
        } else {
            float B[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) B[i] = ew_ld(in, 8 * u + i);
            if (p.kind == EW_HALF) {
                if (!p.sup) {
                    ew_store<sizeof(TIN) == 2, OUT>(p, 2 * u,     __fsub_rn(B[0], B[3]), __fadd_rn(B[1], B[2]));
                    ew_store<sizeof(TIN) == 2, OUT>(p, 2 * u + 1, __fsub_rn(B[7], B[4]), __fsub_rn(-B[5], B[6]));
                } else {
                    ew_store<sizeof(TIN) == 2, OUT>(p, 2 * u,     __fsub_rn(B[1], B[2]), __fsub_rn(-B[0], B[3]));
                    ew_store<sizeof(TIN) == 2, OUT>(p, 2 * u + 1, __fsub_rn(B[6], B[5]), __fadd_rn(B[4], B[7]));
                }
            } else {
                float xr, yi;
                LoaderDiv4<false>::combine(p.sup ? DIV4_SUP : DIV4_INF, B, xr, yi);
                ew_store<sizeof(TIN) == 2, OUT>(p, u, xr, yi);
            }
        }
    }
}

// Decimators<>::decimate1 : (qint16)(buf << pre)   decimators.h:344-356
__global__ void ew_int_copy_kernel(const EwParams p)
{
    const uint32_t* in = reinterpret_cast<const uint32_t*>(p.in);
    uint32_t* out = reinterpret_cast<uint32_t*>(p.out);
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long u = (long long) blockIdx.x * blockDim.x + threadIdx.x; u < p.n_units; u += stride) {
        const uint32_t v = in[u];
        const uint32_t re = ((uint32_t) (int32_t) (short) (v & 0xffffu)) << p.pre;
        const uint32_t im = ((uint32_t) ((int32_t) v >> 16)) << p.pre;
        out[u] = (re & 0xffffu) | (im << 16);
    }
}

// 8-bit inputs: Decimators<qint32,qint8,16,8>::decimate1 / DecimatorsU<...,Shift>::decimate1 (decimatorsu.h:218-231):
// (qint16) ((byte - Shift) << pre1)
template<bool UNSIGNED>
__global__ void ew_int8_copy_kernel(const EwParams p)
{
    const uint16_t* in = reinterpret_cast<const uint16_t*>(p.in);
    uint32_t* out = reinterpret_cast<uint32_t*>(p.out);
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long u = (long long) blockIdx.x * blockDim.x + threadIdx.x; u < p.n_units; u += stride) {
        const uint32_t v = in[u];
        const int32_t x = UNSIGNED ? (int32_t) (v & 0xffu) : (int32_t) (signed char) (v & 0xffu);
        const int32_t y = UNSIGNED ? (int32_t) (v >> 8) : (int32_t) (signed char) (v >> 8);
        const uint32_t re = (uint32_t) (x - p.shift) << p.pre, im = (uint32_t) (y - p.shift) << p.pre;
        out[u] = (re & 0xffffu) | (im << 16);
    }
}

// Decimators<>::decimate2_u (decimators.h:374-393): per 8 scalars two samples,
//   ((b0 - b3) << pre2) >> post2, ((b1 + b2 - 255) << pre2) >> post2;  ((b7 - b4) << pre2) >> post2, ((255 - b5 - b6) << pre2) >> post2
template<typename TIN>
__global__ void ew_int_half_u_kernel(const EwParams p)
{
    const TIN* in = reinterpret_cast<const TIN*>(p.in);
    uint32_t* out = reinterpret_cast<uint32_t*>(p.out);
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long u = (long long) blockIdx.x * blockDim.x + threadIdx.x; u < p.n_units; u += stride) {
        int32_t b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) b[i] = (int32_t) in[8 * u + i];
        const int32_t x0 = (int32_t) ((uint32_t) (b[0] - b[3]) << p.pre) >> p.shift, y0 = (int32_t) ((uint32_t) (b[1] + b[2] - 255) << p.pre) >> p.shift;
        const int32_t x1 = (int32_t) ((uint32_t) (b[7] - b[4]) << p.pre) >> p.shift, y1 = (int32_t) ((uint32_t) (255 - b[5] - b[6]) << p.pre) >> p.shift;
        out[2 * u] = ((uint32_t) x0 & 0xffffu) | ((uint32_t) y0 << 16);
        out[2 * u + 1] = ((uint32_t) x1 & 0xffffu) | ((uint32_t) y1 << 16);
    }
}

// split I / Q arrays -> interleaved (the Decimators<> overloads taking bufI, bufQ: decimators.h:359-371,395-417,2638-...)
template<typename T>
__global__ void interleave_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n)
{
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) { out[2 * i] = a[i]; out[2 * i + 1] = b[i]; }
}

// decimation_shifts<16, InputBits> (decimators.h:79-95 / 115-131 / 151-167), index = log2
const int PRE8[7]  = { 8, 7, 6, 5, 4, 3, 2 }, POST8[7]  = { 0, 0, 0, 0, 0, 0, 0 };
const int PRE12[7] = { 4, 3, 2, 1, 0, 0, 0 }, POST12[7] = { 0, 0, 0, 0, 0, 1, 2 };
const int PRE16[7] = { 0, 0, 0, 0, 0, 0, 0 }, POST16[7] = { 0, 1, 2, 3, 4, 5, 6 };

typedef void (*cascade_fn)(const CascadeParams);

template<typename T, int IN, int OUT, bool HASROT, bool EXACT>
cascade_fn kfn() { return hb64_cascade_kernel<T, IN, OUT, HASROT, EXACT>; }

bool is_int8(int fmt) { return fmt == B200DSP_FMT_I8 || fmt == B200DSP_FMT_U8; }

cascade_fn pick_kernel(int in_fmt, int out_fmt, bool div4, bool hasrot, bool exact, bool pre)
{
    if (in_fmt == B200DSP_FMT_I8) return hasrot ? kfn<int32_t, IN_I8, OUT_I16_SHIFT, true, false>() : kfn<int32_t, IN_I8, OUT_I16_SHIFT, false, false>();
    if (in_fmt == B200DSP_FMT_U8) return hasrot ? kfn<int32_t, IN_U8, OUT_I16_SHIFT, true, false>() : kfn<int32_t, IN_U8, OUT_I16_SHIFT, false, false>();
    if (in_fmt == B200DSP_FMT_I16 && out_fmt == B200DSP_FMT_I16) {
        if (pre) return hasrot ? kfn<int32_t, IN_I16_PRE, OUT_I16_SHIFT, true, false>() : kfn<int32_t, IN_I16_PRE, OUT_I16_SHIFT, false, false>();
        return hasrot ? kfn<int32_t, IN_I16, OUT_I16_SHIFT, true, false>() : kfn<int32_t, IN_I16, OUT_I16_SHIFT, false, false>();
    }
    if (in_fmt == B200DSP_FMT_F32 && out_fmt == B200DSP_FMT_I16) {
        if (div4) return exact ? kfn<float, IN_F32_DIV4, OUT_I16_SCALE, false, true>() : kfn<float, IN_F32_DIV4, OUT_I16_SCALE, false, false>();
        return exact ? kfn<float, IN_F32, OUT_I16_SCALE, false, true>() : kfn<float, IN_F32, OUT_I16_SCALE, false, false>();
    }
    if (in_fmt == B200DSP_FMT_F32 && out_fmt == B200DSP_FMT_F32) {
        if (div4) return exact ? kfn<float, IN_F32_DIV4, OUT_F32, false, true>() : kfn<float, IN_F32_DIV4, OUT_F32, false, false>();
        return exact ? kfn<float, IN_F32, OUT_F32, false, true>() : kfn<float, IN_F32, OUT_F32, false, false>();
    }
    if (div4) return exact ? kfn<float, IN_I16F_DIV4, OUT_F32, false, true>() : kfn<float, IN_I16F_DIV4, OUT_F32, false, false>();
    return exact ? kfn<float, IN_I16F, OUT_F32, false, true>() : kfn<float, IN_I16F, OUT_F32, false, false>();
}

// kernels of a split cascade (stages [0,3) then [3,L)): the hand-off is raw T pairs in an HBM scratch buffer
cascade_fn pick_split_a(int in_fmt, int out_fmt, bool hasrot, bool exact, bool pre)
{
    if (in_fmt == B200DSP_FMT_I8) return hasrot ? kfn<int32_t, IN_I8, OUT_I32, true, false>() : kfn<int32_t, IN_I8, OUT_I32, false, false>();
    if (in_fmt == B200DSP_FMT_U8) return hasrot ? kfn<int32_t, IN_U8, OUT_I32, true, false>() : kfn<int32_t, IN_U8, OUT_I32, false, false>();
    if (in_fmt == B200DSP_FMT_I16 && out_fmt == B200DSP_FMT_I16) {
        if (pre) return hasrot ? kfn<int32_t, IN_I16_PRE, OUT_I32, true, false>() : kfn<int32_t, IN_I16_PRE, OUT_I32, false, false>();
        return hasrot ? kfn<int32_t, IN_I16, OUT_I32, true, false>() : kfn<int32_t, IN_I16, OUT_I32, false, false>();
    }
    if (in_fmt == B200DSP_FMT_F32) return exact ? kfn<float, IN_F32, OUT_F32, false, true>() : kfn<float, IN_F32, OUT_F32, false, false>();
    return exact ? kfn<float, IN_I16F, OUT_F32, false, true>() : kfn<float, IN_I16F, OUT_F32, false, false>();
}
cascade_fn pick_split_b(int in_fmt, int out_fmt, bool hasrot, bool exact)
{
    if ((in_fmt == B200DSP_FMT_I16 || is_int8(in_fmt)) && out_fmt == B200DSP_FMT_I16)
        return hasrot ? kfn<int32_t, IN_I32, OUT_I16_SHIFT, true, false>() : kfn<int32_t, IN_I32, OUT_I16_SHIFT, false, false>();
    if (out_fmt == B200DSP_FMT_I16) return exact ? kfn<float, IN_F32, OUT_I16_SCALE, false, true>() : kfn<float, IN_F32, OUT_I16_SCALE, false, false>();
    return exact ? kfn<float, IN_F32, OUT_F32, false, true>() : kfn<float, IN_F32, OUT_F32, false, false>();
}

struct LaunchGeom { int wpb; int warps_per_sm; };

} // namespace

struct b200dsp_decim {
    int in_fmt, out_fmt, bits, exact;
    int shift;                            // DecimatorsU's Shift template argument (unsigned 8-bit input only)
    int device, sm_count;
    cudaStream_t stream, copy_stream;
    cudaEvent_t ev_h2d[2], ev_done[2];
    void* d_state[2];
    int cur;
    // device staging for the host-pointer path (double-buffered input)
    void* d_in[2]; size_t d_in_cap;
    void* d_out;   size_t d_out_cap;
    void* d_mid;   size_t d_mid_cap;      // hand-off buffer of split cascades (log2 >= 5): stage-3 outputs
};

// ---------------------------------------------------------------------------------------------------------
// call planning (pure host arithmetic)
// ---------------------------------------------------------------------------------------------------------
namespace {

struct Plan {
    bool elementwise; int ew_kind;
    int L; bool div4; int div4_kind; bool hasrot;
    signed char rot[8];
    int pre, post; float out_scale;
    long long consumed_scalars;   // input scalars actually used
    long long n0;                 // stage-0 samples into the cascade
    long long n_out;
};

int make_plan(int in_fmt, int out_fmt, int bits, int log2, int mode, long long len, Plan* pl, bool split = false)
{
    if (log2 < 0 || log2 > 6 || mode < 0 || mode > 3 || len < 0) return B200DSP_EINVAL;
    memset(pl, 0, sizeof(*pl));
    const bool is_int = ((in_fmt == B200DSP_FMT_I16 || is_int8(in_fmt)) && out_fmt == B200DSP_FMT_I16);
    if (mode == B200DSP_MODE_U) {
        // Decimators<>::decimate2_u (decimators.h:374-393): unfiltered /2 of offset-255 data, signed integer inputs only
        if (log2 != 1 || !is_int || in_fmt == B200DSP_FMT_U8) return B200DSP_EINVAL;
        const int* PRE = bits == 8 ? PRE8 : bits == 12 ? PRE12 : PRE16;
        const int* POST = bits == 8 ? POST8 : bits == 12 ? POST12 : POST16;
        pl->pre = PRE[1]; pl->post = POST[1];
        pl->elementwise = true; pl->ew_kind = EW_HALF_U;
        pl->consumed_scalars = (len / 8) * 8;
        pl->n_out = pl->consumed_scalars / 4;
        pl->out_scale = 1.0f;
        return 0;
    }
    const int N = 1 << log2;
    pl->out_scale = 1.0f;
    if (in_fmt == B200DSP_FMT_I16 && out_fmt == B200DSP_FMT_F32)     // decimation_scale<InputBits>::scaleIn
        pl->out_scale = bits == 8 ? 1.0f / 128.0f : bits == 12 ? 1.0f / 2048.0f : 1.0f / 32768.0f;
    if (is_int) {
        const int* PRE = bits == 8 ? PRE8 : bits == 12 ? PRE12 : PRE16;
        const int* POST = bits == 8 ? POST8 : bits == 12 ? POST12 : POST16;
        pl->pre = PRE[log2]; pl->post = POST[log2];
        if (log2 == 0) {
            pl->elementwise = true; pl->ew_kind = EW_COPY;
            pl->n_out = len / 2; pl->consumed_scalars = 2 * pl->n_out;
            return 0;
        }
        // scalars per loop iteration: _cen 2N (N>=8), 16 (N=4), 8 (N=2); _inf/_sup 4N
        // (the split-I/Q _cen overloads take N samples per array per output for every N: decimators.h:2641,2707)
        const long long blk = (mode == B200DSP_MODE_CEN) ? ((N >= 8 || split) ? 2 * N : (N == 4 ? 16 : 8)) : 4 * N;
        const long long nblk = len / blk;
        pl->consumed_scalars = nblk * blk;
        pl->n0 = pl->consumed_scalars / 2;
        pl->L = log2;
        pl->n_out = pl->n0 >> log2;
        if (mode != B200DSP_MODE_CEN) {
            // stage 1 Inf(+j)/Sup(-j); stages 2..L-1 the opposite; last stage centred (N>=8);
            // N=4: (first, opposite); N=2: (first)
            const int first = (mode == B200DSP_MODE_INF) ? +1 : -1;
            pl->hasrot = true;
            pl->rot[1] = (signed char) first;
            if (log2 == 2) pl->rot[2] = (signed char) -first;
            else for (int s = 2; s <= log2 - 1; ++s) pl->rot[s] = (signed char) -first;
        }
        return 0;
    }
    // float cascades
    if (log2 == 0) {
        pl->elementwise = true; pl->ew_kind = EW_COPY;
        pl->n_out = len / 2; pl->consumed_scalars = 2 * pl->n_out;
        return 0;
    }
    if (mode == B200DSP_MODE_CEN) {
        const long long blk = 2 * N;
        pl->consumed_scalars = (len / blk) * blk;
        pl->n0 = pl->consumed_scalars / 2;
        pl->L = log2;
        pl->n_out = pl->n0 >> log2;
        return 0;
    }
    if (log2 == 1) {
        pl->elementwise = true; pl->ew_kind = EW_HALF; pl->div4_kind = (mode == B200DSP_MODE_SUP) ? 1 : 0;
        pl->consumed_scalars = (len / 8) * 8;
        pl->n_out = pl->consumed_scalars / 4;
        return 0;
    }
    {
        const long long blk = 2 * N;
        pl->consumed_scalars = (len / blk) * blk;
        if (log2 == 2) {
            pl->elementwise = true; pl->ew_kind = EW_DIV4; pl->div4_kind = (mode == B200DSP_MODE_SUP) ? 1 : 0;
            pl->n_out = pl->consumed_scalars / 8;
            return 0;
        }
        pl->div4 = true;
        pl->div4_kind = (mode == B200DSP_MODE_INF) ? DIV4_INF : (N >= 16 ? DIV4_SUP16 : DIV4_SUP);
        pl->n0 = pl->consumed_scalars / 8;
        pl->L = log2 - 2;
        pl->n_out = pl->n0 >> pl->L;
    }
    return 0;
}

// best (warps per block, resident warps per SM) for a cascade kernel with L stages
// (the shared-memory opt-in and the occupancy belong to one device: the cache is keyed by it)
LaunchGeom pick_geom(int device, cascade_fn fn, int L)
{
    static std::mutex mu;
    static std::map<std::pair<std::pair<int, const void*>, int>, LaunchGeom> cache;
    std::lock_guard<std::mutex> g(mu);
    auto key = std::make_pair(std::make_pair(device, (const void*) fn), L);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    LaunchGeom best = { 1, 1 };
    cudaFuncSetAttribute((const void*) fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    for (int wpb = 1; wpb <= 8; ++wpb) {
        const size_t smem = (size_t) wpb * L * HB_STAGE_BYTES;
        if (smem > 227 * 1024) break;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*) fn, wpb * 32, smem) != cudaSuccess) { cudaGetLastError(); continue; }
        if (nb * wpb > best.warps_per_sm || (nb * wpb == best.warps_per_sm && wpb < best.wpb)) { best.wpb = wpb; best.warps_per_sm = nb * wpb; }
    }
    cache[key] = best;
    return best;
}

// one kernel over stages [base, base + L) of the plan; `first`/`final` = this segment reads the call's input / writes its output
int launch_segment(b200dsp_decim* h, cascade_fn fn, const Plan& pl, int base, int L, const void* d_in, void* d_out, long long n0,
                   bool first, bool final, cudaStream_t st)
{
    const LaunchGeom g = pick_geom(h->device, fn, L);
    CascadeParams p;
    memset(&p, 0, sizeof(p));
    p.in = d_in; p.out = d_out;
    const size_t esz = 4;
    p.state_in = (const char*) h->d_state[h->cur] + (size_t) base * 128 * esz;
    p.state_out = (char*) h->d_state[h->cur ^ 1] + (size_t) base * 128 * esz;
    p.n0 = n0; p.n_out = n0 >> L; p.L = L;
    p.pre = first ? pl.pre : 0; p.post = final ? pl.post : 0;
    p.in_mul = 1 << pl.pre; p.in_add = -(h->shift << pl.pre);
    p.out_scale = final ? pl.out_scale : 1.0f; p.div4 = pl.div4_kind;
    for (int s = 1; s <= L && s + base < 8; ++s) p.rot[s] = pl.rot[s + base];
    p.opq_zero = 0; p.opq_one = 1; p.opq_mone = -1;
    p.copy_end = final ? (HB_MAX_STAGES - base) * 128 : L * 128;
    const long long U = (long long) HB_IN << (L - 1);
    const long long sp_total = (n0 + U - 1) / U;
    const long long max_warps = (long long) h->sm_count * g.warps_per_sm;
    long long slice_sp = (sp_total + max_warps - 1) / max_warps;
    if (slice_sp < 1) slice_sp = 1;
    const long long n_slices = (sp_total + slice_sp - 1) / slice_sp;
    p.slice_sp = (int) slice_sp; p.n_slices = (int) n_slices;
    int wpb = g.wpb;
    if (n_slices < (long long) h->sm_count * wpb) {       // few slices: spread them over the SMs
        wpb = (int) ((n_slices + h->sm_count - 1) / h->sm_count);
        if (wpb < 1) wpb = 1;
    }
    const unsigned grid = (unsigned) ((n_slices + wpb - 1) / wpb);
    const size_t smem = (size_t) wpb * L * HB_STAGE_BYTES;
    fn<<<grid, wpb * 32, smem, st>>>(p);
    return B200_CUDA_CHECK(cudaGetLastError());
}

int launch_plan(b200dsp_decim* h, const Plan& pl, const void* d_in, void* d_out, cudaStream_t st)
{
    if (pl.n_out <= 0) return 0;
    if (pl.elementwise) {
        EwParams ep;
        ep.in = d_in; ep.out = d_out; ep.kind = pl.ew_kind; ep.pre = pl.pre; ep.out_scale = pl.out_scale; ep.shift = h->shift;
        ep.sup = pl.div4_kind;   // HALF/DIV4: 0 inf, 1 sup
        ep.n_units = pl.ew_kind == EW_COPY ? pl.n_out : pl.consumed_scalars / 8;
        const int threads = 256;
        long long blocks = (ep.n_units + threads - 1) / threads;
        if (blocks > (long long) h->sm_count * 16) blocks = (long long) h->sm_count * 16;
        if (blocks < 1) blocks = 1;
        if (pl.ew_kind == EW_HALF_U) {
            ep.shift = pl.post;               // (the 8-bit Shift does not enter decimate2_u: the 255 offset is in the formula)
            if (h->in_fmt == B200DSP_FMT_I8) ew_int_half_u_kernel<signed char><<<(unsigned) blocks, threads, 0, st>>>(ep);
            else ew_int_half_u_kernel<int16_t><<<(unsigned) blocks, threads, 0, st>>>(ep);
        }
        else if (h->in_fmt == B200DSP_FMT_I8) ew_int8_copy_kernel<false><<<(unsigned) blocks, threads, 0, st>>>(ep);
        else if (h->in_fmt == B200DSP_FMT_U8) ew_int8_copy_kernel<true><<<(unsigned) blocks, threads, 0, st>>>(ep);
        else if (h->in_fmt == B200DSP_FMT_I16 && h->out_fmt == B200DSP_FMT_I16) ew_int_copy_kernel<<<(unsigned) blocks, threads, 0, st>>>(ep);
        else if (h->in_fmt == B200DSP_FMT_F32 && h->out_fmt == B200DSP_FMT_I16) ew_float_kernel<float, OUT_I16_SCALE><<<(unsigned) blocks, threads, 0, st>>>(ep);
        else if (h->in_fmt == B200DSP_FMT_F32) ew_float_kernel<float, OUT_F32><<<(unsigned) blocks, threads, 0, st>>>(ep);
        else ew_float_kernel<int16_t, OUT_F32><<<(unsigned) blocks, threads, 0, st>>>(ep);
        return B200_CUDA_CHECK(cudaGetLastError());
    }
    // Cascades of 5 or 6 stages are run as stages [0,3) + [3,L): per-warp shared memory drops from L*3.5 KB to <= 10.5 KB, so
    // 16+ warps stay resident per SM instead of 10-12; the hand-off costs 2 x 8 B per 8 input samples of HBM traffic.
    const bool split = (pl.L >= 5) && !pl.div4;
    int rc;
    if (!split) {
        rc = launch_segment(h, pick_kernel(h->in_fmt, h->out_fmt, pl.div4, pl.hasrot, h->exact != 0, pl.pre != 0), pl, 0, pl.L, d_in, d_out,
                            pl.n0, true, true, st);
    } else {
        const long long n_mid = pl.n0 >> 3;
        const size_t need = (size_t) n_mid * 8;
        if (h->d_mid_cap < need) {
            if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(st)))) return rc;
            if (h->d_mid) cudaFree(h->d_mid);
            h->d_mid = nullptr; h->d_mid_cap = 0;
            if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_mid, need)))) return rc;
            h->d_mid_cap = need;
        }
        rc = launch_segment(h, pick_split_a(h->in_fmt, h->out_fmt, pl.hasrot, h->exact != 0, pl.pre != 0), pl, 0, 3, d_in, h->d_mid, pl.n0, true, false, st);
        if (rc == 0)
            rc = launch_segment(h, pick_split_b(h->in_fmt, h->out_fmt, pl.hasrot, h->exact != 0), pl, 3, pl.L - 3, h->d_mid, d_out, n_mid, false, true, st);
    }
    if (rc == 0) h->cur ^= 1;
    return rc;
}

bool valid_formats(int in_fmt, int out_fmt)
{
    if (is_int8(in_fmt)) return out_fmt == B200DSP_FMT_I16;
    return (in_fmt == B200DSP_FMT_I16 || in_fmt == B200DSP_FMT_F32) && (out_fmt == B200DSP_FMT_I16 || out_fmt == B200DSP_FMT_F32);
}

size_t in_elem_bytes(int fmt) { return is_int8(fmt) ? 1 : fmt == B200DSP_FMT_I16 ? 2 : 4; }

} // namespace

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int64_t b200dsp_decim_out_count(int in_fmt, int out_fmt, int log2_decim, int mode, int64_t len_scalars)
{
    Plan pl;
    if (!valid_formats(in_fmt, out_fmt)) return -1;
    if (make_plan(in_fmt, out_fmt, is_int8(in_fmt) ? 8 : 12, log2_decim, mode, len_scalars, &pl) != 0) return -1;
    return pl.n_out;
}

int b200dsp_decim_create(b200dsp_decim_t** out, int in_fmt, int out_fmt, int input_bits)
{
    if (!out) return b200_fail(B200DSP_EINVAL, "decim_create: null handle pointer");
    *out = nullptr;
    if (!valid_formats(in_fmt, out_fmt))
        return b200_fail(B200DSP_EINVAL, "decim_create: bad sample format (8-bit inputs decimate to int16 only)");
    if (input_bits != 8 && input_bits != 12 && input_bits != 16)
        return b200_fail(B200DSP_EINVAL, "decim_create: input_bits must be 8, 12 or 16");
    if (is_int8(in_fmt) && input_bits != 8)
        return b200_fail(B200DSP_EINVAL, "decim_create: 8-bit sample formats take input_bits = 8");
    int rc = b200_require_device();
    if (rc) return rc;
    b200dsp_decim* h = new (std::nothrow) b200dsp_decim();
    if (!h) return b200_fail(B200DSP_ENOMEM, "decim_create: out of host memory");
    memset(h, 0, sizeof(*h));
    h->in_fmt = in_fmt; h->out_fmt = out_fmt; h->bits = input_bits;
    h->shift = (in_fmt == B200DSP_FMT_U8) ? 127 : 0;     // rtlsdrthread.h:55
    h->device = b200_current_device();
    h->sm_count = b200_sm_count_of(h->device);
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(h->device))) ||
        (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking))) ||
        (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)))) { delete h; return rc; }
    for (int i = 0; i < 2; ++i) {
        if ((rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming))) ||
            (rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming))) ||
            (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_state[i], B200DSP_DECIM_STATE_ELEMS * 4))) ||
            (rc = B200_CUDA_CHECK(cudaMemset(h->d_state[i], 0, B200DSP_DECIM_STATE_ELEMS * 4)))) { b200dsp_decim_destroy(h); return rc; }
    }
    if ((rc = B200_CUDA_CHECK(cudaDeviceSynchronize()))) { b200dsp_decim_destroy(h); return rc; }   // memsets ran on the default stream
    *out = h;
    return 0;
}

int b200dsp_decim_destroy(b200dsp_decim_t* h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    for (int i = 0; i < 2; ++i) {
        if (h->d_state[i]) cudaFree(h->d_state[i]);
        if (h->d_in[i]) cudaFree(h->d_in[i]);
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    }
    if (h->d_out) cudaFree(h->d_out);
    if (h->d_mid) cudaFree(h->d_mid);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    delete h;
    return 0;
}

int b200dsp_decim_set_shift(b200dsp_decim_t* h, int shift)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    if (h->in_fmt != B200DSP_FMT_U8) return b200_fail(B200DSP_ESTATE, "decim_set_shift: only the unsigned 8-bit input format has a shift");
    if (shift < 0 || shift > 255) return b200_fail(B200DSP_EINVAL, "decim_set_shift: shift must be in [0, 255]");
    h->shift = shift;
    return 0;
}

int b200dsp_decim_set_exact_float(b200dsp_decim_t* h, int exact)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    h->exact = exact ? 1 : 0;
    return 0;
}

int b200dsp_decim_run_dev(b200dsp_decim_t* h, int log2_decim, int mode, const void* d_in, int64_t len_scalars,
                          void* d_out, int64_t* n_out, void* cuda_stream)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    Plan pl;
    if (make_plan(h->in_fmt, h->out_fmt, h->bits, log2_decim, mode, len_scalars, &pl)) return b200_fail(B200DSP_EINVAL, "decim_run: bad log2/mode/len");
    if (n_out) *n_out = pl.n_out;
    if (pl.n_out == 0) return 0;
    if (!d_in || !d_out) return b200_fail(B200DSP_EINVAL, "decim_run: null buffer");
    if (((uintptr_t) d_in & 15) || ((uintptr_t) d_out & 7)) return b200_fail(B200DSP_EINVAL, "decim_run_dev: d_in must be 16-byte and d_out 8-byte aligned");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    return launch_plan(h, pl, d_in, d_out, cuda_stream ? (cudaStream_t) cuda_stream : h->stream);
}

// Host-pointer form.  Large calls are cut into sub-calls (whole blocks, state carried => identical result) so the
// H2D copy of chunk i+1 overlaps the kernel of chunk i.
int b200dsp_decim_run(b200dsp_decim_t* h, int log2_decim, int mode, const void* in, int32_t len_scalars, void* out, int32_t* n_out)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    Plan whole;
    if (make_plan(h->in_fmt, h->out_fmt, h->bits, log2_decim, mode, len_scalars, &whole)) return b200_fail(B200DSP_EINVAL, "decim_run: bad log2/mode/len");
    if (n_out) *n_out = (int32_t) whole.n_out;
    if (whole.n_out == 0) return 0;
    if (!in || !out) return b200_fail(B200DSP_EINVAL, "decim_run: null buffer");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    const size_t ib = in_elem_bytes(h->in_fmt), ob = in_elem_bytes(h->out_fmt) * 2;
    // chunk: multiple of the largest block (512 scalars covers every mode's block size) and of the kernel's alignment needs
    const long long CHUNK_SCALARS = 8ll << 20;
    const long long total = whole.consumed_scalars;
    const long long chunk = total < CHUNK_SCALARS ? total : CHUNK_SCALARS;
    const size_t need_in = (size_t) chunk * ib;
    if (h->d_in_cap < need_in) {
        for (int i = 0; i < 2; ++i) { if (h->d_in[i]) cudaFree(h->d_in[i]); h->d_in[i] = nullptr; }
        h->d_in_cap = 0;
        for (int i = 0; i < 2; ++i) if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_in[i], need_in)))) return rc;
        h->d_in_cap = need_in;
    }
    const size_t need_out = (size_t) whole.n_out * ob;
    if (h->d_out_cap < need_out) {
        if (h->d_out) cudaFree(h->d_out);
        h->d_out = nullptr; h->d_out_cap = 0;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_out, need_out)))) return rc;
        h->d_out_cap = need_out;
    }
    long long done = 0, out_done = 0;
    int slot = 0;
    int nchunks = 0;
    while (done < total) {
        const long long n = (total - done) < chunk ? (total - done) : chunk;
        Plan pl;
        make_plan(h->in_fmt, h->out_fmt, h->bits, log2_decim, mode, n, &pl);
        if (nchunks >= 2 && (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(h->copy_stream, h->ev_done[slot], 0)))) return rc;
        if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_in[slot], (const char*) in + (size_t) done * ib, (size_t) n * ib, cudaMemcpyHostToDevice, h->copy_stream))) ||
            (rc = B200_CUDA_CHECK(cudaEventRecord(h->ev_h2d[slot], h->copy_stream))) ||
            (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(h->stream, h->ev_h2d[slot], 0)))) return rc;
        char* dout = (char*) h->d_out + (size_t) out_done * ob;
        if ((rc = launch_plan(h, pl, h->d_in[slot], dout, h->stream))) return rc;
        if ((rc = B200_CUDA_CHECK(cudaEventRecord(h->ev_done[slot], h->stream)))) return rc;
        if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync((char*) out + (size_t) out_done * ob, dout, (size_t) pl.n_out * ob, cudaMemcpyDeviceToHost, h->stream)))) return rc;
        done += n; out_done += pl.n_out; slot ^= 1; ++nchunks;
    }
    return B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

// == the Decimators<> overloads on separate I and Q arrays (decimators.h:359-371 decimate1, :395-417 decimate2_u, :2638-2700 ...
//    :3888-... decimateN_cen): `len_per_array` samples in each of in_i, in_q.  The arrays are interleaved on the device
//    and take the same kernels as the interleaved entry points (the reference's two forms compute the same thing).
int b200dsp_decim_run_split(b200dsp_decim_t* h, int log2_decim, int mode, const void* in_i, const void* in_q, int32_t len_per_array, void* out, int32_t* n_out)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    if (len_per_array < 0 || len_per_array > (1 << 29)) return b200_fail(B200DSP_EINVAL, "decim_run_split: bad length");
    if (mode != B200DSP_MODE_CEN && mode != B200DSP_MODE_U) return b200_fail(B200DSP_EINVAL, "decim_run_split: the reference defines the split overloads for decimate1, decimate2_u and decimateN_cen only");
    Plan whole;
    if (make_plan(h->in_fmt, h->out_fmt, h->bits, log2_decim, mode, 2ll * len_per_array, &whole, true)) return b200_fail(B200DSP_EINVAL, "decim_run_split: bad log2/mode/len");
    if (n_out) *n_out = (int32_t) whole.n_out;
    if (whole.n_out == 0) return 0;
    if (!in_i || !in_q || !out) return b200_fail(B200DSP_EINVAL, "decim_run_split: null buffer");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    const size_t ib = in_elem_bytes(h->in_fmt), ob = in_elem_bytes(h->out_fmt) * 2;
    const long long n = whole.consumed_scalars / 2;                 // samples per array that take part
    void *d_i = nullptr, *d_q = nullptr, *d_x = nullptr, *d_o = nullptr;
    if ((rc = B200_CUDA_CHECK(cudaMalloc(&d_i, (size_t) n * ib))) || (rc = B200_CUDA_CHECK(cudaMalloc(&d_q, (size_t) n * ib))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&d_x, (size_t) n * ib * 2 + 64))) || (rc = B200_CUDA_CHECK(cudaMalloc(&d_o, (size_t) whole.n_out * ob + 64)))) goto done;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(d_i, in_i, (size_t) n * ib, cudaMemcpyHostToDevice, h->stream))) ||
        (rc = B200_CUDA_CHECK(cudaMemcpyAsync(d_q, in_q, (size_t) n * ib, cudaMemcpyHostToDevice, h->stream)))) goto done;
    {
        const unsigned blocks = (unsigned) ((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
        if (ib == 1) interleave_kernel<unsigned char><<<blocks, 256, 0, h->stream>>>((const unsigned char*) d_i, (const unsigned char*) d_q, (unsigned char*) d_x, n);
        else if (ib == 2) interleave_kernel<uint16_t><<<blocks, 256, 0, h->stream>>>((const uint16_t*) d_i, (const uint16_t*) d_q, (uint16_t*) d_x, n);
        else interleave_kernel<uint32_t><<<blocks, 256, 0, h->stream>>>((const uint32_t*) d_i, (const uint32_t*) d_q, (uint32_t*) d_x, n);
        if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) goto done;
    }
    if ((rc = launch_plan(h, whole, d_x, d_o, h->stream))) goto done;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(out, d_o, (size_t) whole.n_out * ob, cudaMemcpyDeviceToHost, h->stream)))) goto done;
    rc = B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
done:
    if (d_i) cudaFree(d_i);
    if (d_q) cudaFree(d_q);
    if (d_x) cudaFree(d_x);
    if (d_o) cudaFree(d_o);
    return rc;
}

int b200dsp_decim_get_state(b200dsp_decim_t* h, void* state_host)
{
    if (!h || !state_host) return b200_fail(B200DSP_EINVAL, "null argument");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(h->stream)))) return rc;
    return B200_CUDA_CHECK(cudaMemcpy(state_host, h->d_state[h->cur], B200DSP_DECIM_STATE_ELEMS * 4, cudaMemcpyDeviceToHost));
}

int b200dsp_decim_set_state(b200dsp_decim_t* h, const void* state_host)
{
    if (!h || !state_host) return b200_fail(B200DSP_EINVAL, "null argument");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(h->stream)))) return rc;
    return B200_CUDA_CHECK(cudaMemcpy(h->d_state[h->cur], state_host, B200DSP_DECIM_STATE_ELEMS * 4, cudaMemcpyHostToDevice));
}

int b200dsp_decim_reset(b200dsp_decim_t* h)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(h->stream)))) return rc;
    return B200_CUDA_CHECK(cudaMemset(h->d_state[h->cur], 0, B200DSP_DECIM_STATE_ELEMS * 4));
}

int b200dsp_decim_sync(b200dsp_decim_t* h)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    return B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

} // extern "C"
