// tx_api.cu — the Tx mirror of the path: b200dsp_interps_* (K8) and b200dsp_upchan_* (K9)
//
// Replaces (paths relative to the reference tree):
//   Interpolators<T,SdrBits,OutputBits>::interpolate{1,2,4,...,64}_cen        sdrbase/dsp/interpolators.h:104-617
//   interpolation_shifts<16,{8,12,16}>                                        sdrbase/dsp/interpolators.h:30-102
//   IntHalfbandFilterEO1<order>::myInterpolate / doInterpolateFIR             sdrbase/dsp/inthalfbandfiltereo1.h:601-622,797-815
//   IntHalfbandFilterEO1<96>::workInterpolate{Center,LowerHalf,UpperHalf}     sdrbase/dsp/inthalfbandfiltereo1.h:98-127,291-355,490-554
//   UpChannelizer::pull / applyConfiguration / createFilterChain              sdrbase/dsp/upchannelizer.cpp:51-104,175-209,252-327
//   coefficient tables HBFIRFilterTraits<16|32|64|96>::hbCoeffs               sdrbase/dsp/hbfiltertraits.cpp
//
// Closed forms (probe-verified against the compiled reference, oracle/ref_capi_tx.cpp).  An interpolating half-band of
// order O keeps the last O/2 inputs x; input n produces two outputs
//     y[2n]   = x[n - O/4]                                                       (the centre tap: a pure delay)
//     y[2n+1] = ( sum_{i < O/4} h[i] * ( x[n - i] + x[n - O/2 + 1 + i] ) ) >> (hbShift - 1)       int32, wrapping
// Interpolators<>: x0 = sample << pre, stages of order 64, 32, 16, 16, 16, 16 on int32 streams, out = (T) (x_L >> post).
// UpChannelizer: stages of order 96 on int16 streams (the >> 15 result is stored into a Sample: wraps), stage 0 at the output
// rate; a stage consumes its input on its odd calls only and the consumed sample is the one the next stage produced on ITS
// previous call, so stage s sees the stream of stage s+1 delayed by one sample (zero first); lower/upper-half stages multiply
// output k by (-j)^(k+1) / (+j)^(k+1) with int16 wrap of -(-32768).
#include "common.cuh"
#include <vector>

using namespace b200dsp;

namespace {

// ---------------------------------------------------------------------------------------------------------
// coefficients: (int32_t) (c * 4096) for orders 16/32/64, (int32_t) (c * 65536) for order 96 (hbfiltertraits.cpp), values
// checked against the compiled reference tables in tests/test_tx_gpu.py / test_oracle_port.py
// ---------------------------------------------------------------------------------------------------------
template<int H> struct HbTaps;          // H = order / 2 = ring length
template<> struct HbTaps<8>  { static __host__ __device__ constexpr int h(int i) { constexpr int t[4] = { -21, 95, -311, 1260 }; return t[i]; } };
template<> struct HbTaps<16> { static __host__ __device__ constexpr int h(int i) { constexpr int t[8] = { -7, 15, -33, 65, -117, 207, -401, 1294 }; return t[i]; } };
template<> struct HbTaps<32> { static __host__ __device__ constexpr int h(int i) { constexpr int t[16] = { -1, 2, -5, 8, -12, 17, -25, 35, -47, 64, -86, 117, -164, 244, -424, 1300 }; return t[i]; } };
template<> struct HbTaps<48> { static __host__ __device__ constexpr int h(int i) { constexpr int t[24] = { -1, 3, -6, 11, -19, 31, -47, 70, -99, 139, -189, 254, -335, 436, -563, 722, -923, 1181, -1525, 2004, -2730, 3990, -6842, 20823 }; return t[i]; } };

// Pipe placement: ptxas turns 2-input integer adds into IMAD.IADD and piles them onto the FMA-heavy pipe
// next to the multiply-adds; a 3-input add with an opaque zero stays an IADD3 on the ALU pipe, so the pre-adds and the
// multiply-adds of the FIR share the two 64-lane integer pipes evenly.
__device__ __forceinline__ int tx_add3(int a, int b, int z)
{
    int r;
    asm("{.reg .s32 t; add.s32 t, %1, %2; add.s32 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(z));
    return r;
}

// four consecutive FIR outputs from the register window w[q] = x[n0 - H + q], q in [0, H + 4):
//   fir[r] = sum_i h[i] * ( w[H + r - i] + w[r + 1 + i] )          mid[r] = w[H / 2 + r]
template<int H>
__device__ __forceinline__ void hbint_group(const int (&w)[H + 4], int (&fir)[4], int zero)
{
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int acc = 0;
#pragma unroll
        for (int i = 0; i < H / 2; ++i) acc += HbTaps<H>::h(i) * tx_add3(w[H + r - i], w[r + 1 + i], zero);
        fir[r] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------
// K8: Interpolators<> cascade, fused: a CTA walks a contiguous range of tiles of TN = 4096 >> L samples (4096 output samples
// per tile); the stage streams of the tile live in shared memory as [level][component][history + new] int32 arrays, a level's
// newest H samples slide to the front as the next tile's history.  A thread item = (4 consecutive inputs of one level, one
// component): one aligned register window (H/4 + 1 128-bit shared loads), 4 FIR + 4 centre outputs.  Lanes 0-15 own the
// real component of 16 consecutive groups, lanes 16-31 the imaginary one: the last stage exchanges through one shuffle per
// value and each lane stores 4 packed output samples, a warp 512 contiguous bytes.  A CTA whose range starts inside the
// stream first runs the previous tile with zero history and its stores off (FIR: finite memory, 42 input samples deep).
// ---------------------------------------------------------------------------------------------------------
constexpr int IP_THREADS = 256;
constexpr int IP_LEVELS = 6;
constexpr int IP_STATE_WORDS = IP_LEVELS * 2 * 32;       // per level and component the last 32 (16, 8) stage inputs, oldest first

struct InterpParams {
    const uint32_t* in;        // packed int16 IQ samples
    void* out;                 // int16 or int8 interleaved I,Q
    const int* st_in;          // [6][2][32]
    int* st_out;
    long long n;               // samples to consume
    int L, pre, post;
    int out_i8;
    int quirk110;              // interpolate64_cen writes only scalars 0..109 of each block of 128 (interpolators.h, its last loop)
    int tiles, tiles_per_cta;
    int opq_zero;              // 0, passed as data (tx_add3)
};

__host__ __device__ __forceinline__ constexpr int ip_hist(int l) { return l == 0 ? 32 : (l == 1 ? 16 : 8); }
// level l: [2][len_l] ints, len_l = hist_l + (TN << l) + 4 (window overrun of the last group), a multiple of 4; levels are laid
// out back to back, so the offsets are closed forms (no per-thread table in local memory)
__host__ __device__ __forceinline__ constexpr int ip_len(int l, int TN) { return ip_hist(l) + (TN << l) + 4; }
__host__ __device__ __forceinline__ constexpr int ip_off(int l, int TN)
{
    // hist_0 + ... + hist_{l-1} = 0, 32, 48, 56, ...
    return 2 * ((l == 0 ? 0 : (l == 1 ? 32 : 32 + 8 * l)) + 4 * l + TN * ((1 << l) - 1));
}

template<int H, bool LAST>
__device__ __forceinline__ void ip_stage(const InterpParams& p, int* cur, int cur_len, int* nxt, int nxt_len, int nxt_hist,
                                         int n_new, long long out_sample0, bool store)
{
    // cur: [2][cur_len] with history at [0, H), new samples at [H, H + n_new); n_new a multiple of 4 (tiles are) or ragged (last tile)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = lane >> 4;
    const int groups = (n_new + 3) >> 2;
    for (int gb = warp * 16; gb < groups; gb += (IP_THREADS / 32) * 16) {
        const int g = gb + (lane & 15);
        const bool act = g < groups;
        const int n0 = 4 * (act ? g : 0);
        int w[H + 4];
        const int4* src = reinterpret_cast<const int4*>(cur + c * cur_len + n0);
#pragma unroll
        for (int q = 0; q < (H + 4) / 4; ++q) { const int4 v = src[q]; w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w; }
        int fir[4];
        hbint_group<H>(w, fir, p.opq_zero);
        int y[8];
#pragma unroll
        for (int r = 0; r < 4; ++r) { y[2 * r] = w[H / 2 + r]; y[2 * r + 1] = fir[r] >> 11; }
        if (!LAST) {
            if (act) {
                int4* dst = reinterpret_cast<int4*>(nxt + c * nxt_len + nxt_hist + 2 * n0);
                dst[0] = make_int4(y[0], y[1], y[2], y[3]);
                dst[1] = make_int4(y[4], y[5], y[6], y[7]);
            }
        } else {
            // exchange: lanes 0-15 (real) take output samples 0..3 of the group, lanes 16-31 (imaginary) samples 4..7
            int mine[4], other[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int give = c ? y[k] : y[4 + k];              // what the partner needs from me
                mine[k] = (c ? y[4 + k] : y[k]) >> p.post;
                other[k] = __shfl_xor_sync(0xffffffffu, give, 16) >> p.post;
            }
            if (act && store) {
                const long long s0 = out_sample0 + 2ll * n0 + 4 * c;      // first of this lane's 4 output samples (stream index within the call)
                const int valid = 2 * n_new - (2 * n0 + 4 * c);           // output samples of this tile from s0 on
                int lim = valid < 4 ? valid : 4;
                if (p.quirk110) {                                         // samples 55..63 of every 64 are never written
                    const int pos = (int) (s0 & 63);
                    const int q = 55 - pos;
                    if (q < lim) lim = q;
                }
                if (p.out_i8) {
                    uint16_t v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int re = c ? other[k] : mine[k], im = c ? mine[k] : other[k];
                        v[k] = (uint16_t) ((re & 0xff) | ((im & 0xff) << 8));
                    }
                    uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) + s0;
                    if (lim == 4) *reinterpret_cast<uint2*>(dst) = make_uint2((uint32_t) v[0] | ((uint32_t) v[1] << 16), (uint32_t) v[2] | ((uint32_t) v[3] << 16));
                    else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) if (k < lim) dst[k] = v[k];
                    }
                } else {
                    uint32_t v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int re = c ? other[k] : mine[k], im = c ? mine[k] : other[k];
                        v[k] = ((uint32_t) re & 0xffffu) | ((uint32_t) im << 16);
                    }
                    uint32_t* dst = reinterpret_cast<uint32_t*>(p.out) + s0;
                    if (lim == 4) *reinterpret_cast<uint4*>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
                    else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) if (k < lim) dst[k] = v[k];
                    }
                }
            }
        }
    }
}

// The last stage: a thread owns 4 consecutive inputs of BOTH components (two register windows), so the 8 output samples are
// packed in registers (no exchange, no selects) and leave as one or two 128-bit stores.
template<int H>
__device__ __forceinline__ void ip_stage_last(const InterpParams& p, const int* cur, int cur_len, int n_new, long long out_sample0, bool store)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int groups = (n_new + 3) >> 2;
    const int shm = p.post, shf = 11 + p.post;          // (acc >> 11) >> post == acc >> (11 + post): arithmetic shifts compose
    for (int g = warp * 32 + lane; g < groups; g += IP_THREADS) {
        const int n0 = 4 * g;
        int wr[H + 4], wi[H + 4];
        const int4* sr = reinterpret_cast<const int4*>(cur + n0);
        const int4* si = reinterpret_cast<const int4*>(cur + cur_len + n0);
#pragma unroll
        for (int q = 0; q < (H + 4) / 4; ++q) {
            const int4 a = sr[q], b = si[q];
            wr[4 * q] = a.x; wr[4 * q + 1] = a.y; wr[4 * q + 2] = a.z; wr[4 * q + 3] = a.w;
            wi[4 * q] = b.x; wi[4 * q + 1] = b.y; wi[4 * q + 2] = b.z; wi[4 * q + 3] = b.w;
        }
        int fr[4], fi[4];
        hbint_group<H>(wr, fr, p.opq_zero);
        hbint_group<H>(wi, fi, p.opq_zero);
        if (!store) continue;
        const long long s0 = out_sample0 + 2ll * n0;               // first of this thread's 8 output samples (index within the call)
        int lim = 2 * n_new - 2 * n0;                               // output samples of this tile from s0 on
        if (lim > 8) lim = 8;
        if (p.quirk110) { const int q = 55 - (int) (s0 & 63); if (q < lim) lim = q; }      // samples 55..63 of every 64 are never written
        if (p.out_i8) {
            uint32_t h16[8];                                        // one output sample = (int8 re, int8 im)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                h16[2 * r] = __byte_perm((uint32_t) (wr[H / 2 + r] >> shm), (uint32_t) (wi[H / 2 + r] >> shm), 0x0040);
                h16[2 * r + 1] = __byte_perm((uint32_t) (fr[r] >> shf), (uint32_t) (fi[r] >> shf), 0x0040);
            }
            uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) + s0;
            if (lim == 8) {
                *reinterpret_cast<uint4*>(dst) = make_uint4(__byte_perm(h16[0], h16[1], 0x5410), __byte_perm(h16[2], h16[3], 0x5410),
                                                            __byte_perm(h16[4], h16[5], 0x5410), __byte_perm(h16[6], h16[7], 0x5410));
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) if (k < lim) dst[k] = (uint16_t) (h16[k] & 0xffffu);
            }
        } else {
            uint32_t v[8];                                          // one output sample = (int16 re, int16 im)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                v[2 * r] = __byte_perm((uint32_t) (wr[H / 2 + r] >> shm), (uint32_t) (wi[H / 2 + r] >> shm), 0x5410);
                v[2 * r + 1] = __byte_perm((uint32_t) (fr[r] >> shf), (uint32_t) (fi[r] >> shf), 0x5410);
            }
            uint32_t* dst = reinterpret_cast<uint32_t*>(p.out) + s0;
            if (lim == 8) {
                reinterpret_cast<uint4*>(dst)[0] = make_uint4(v[0], v[1], v[2], v[3]);
                reinterpret_cast<uint4*>(dst)[1] = make_uint4(v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) if (k < lim) dst[k] = v[k];
            }
        }
    }
}

// the levels of one tile, unrolled at compile time: every offset and length is a constant of the instantiation
template<int L, int l>
struct IpLevels {
    static __device__ __forceinline__ void run(const InterpParams& p, int* smem, int nv, long long os0, bool store)
    {
        constexpr int TN = 4096 >> L, H = ip_hist(l);
        int* cur = smem + ip_off(l, TN);
        const int n_new = nv << l;
        if constexpr (l == L - 1) {
            ip_stage_last<H>(p, cur, ip_len(l, TN), n_new, os0, store);
            __syncthreads();
        } else {
            ip_stage<H, false>(p, cur, ip_len(l, TN), smem + ip_off(l + 1, TN), ip_len(l + 1, TN), ip_hist(l + 1), n_new, os0, store);
            __syncthreads();
            IpLevels<L, l + 1>::run(p, smem, nv, os0, store);
        }
    }
};

#ifndef IP_MINB
#define IP_MINB 4
#endif
template<int L>
__global__ void __launch_bounds__(IP_THREADS, IP_MINB) interps_cascade_kernel(const InterpParams p)
{
    extern __shared__ __align__(16) int ip_smem[];
    const int tid = threadIdx.x;
    constexpr int TN = 4096 >> L;
    const int t_begin = blockIdx.x * p.tiles_per_cta;
    int t_end = t_begin + p.tiles_per_cta;
    if (t_end > p.tiles) t_end = p.tiles;
    if (t_begin >= t_end) return;
    // histories: the handle's state for the stream's first tile, zeros (+ one warm-up tile) elsewhere
    for (int e = tid; e < IP_LEVELS * 2 * 32; e += IP_THREADS) {
        const int l = e >> 6, c = (e >> 5) & 1, q = e & 31;
        if (l < L && q < ip_hist(l)) ip_smem[ip_off(l, TN) + c * ip_len(l, TN) + q] = (t_begin == 0) ? p.st_in[e] : 0;
    }
    __syncthreads();
    for (int t = (t_begin == 0 ? 0 : t_begin - 1); t < t_end; ++t) {
        const bool store = (t >= t_begin);
        const long long i0 = (long long) t * TN;
        long long rem = p.n - i0;
        const int nv = rem < TN ? (int) rem : TN;                  // samples of this tile (ragged only for the stream's last tile)
        // level 0: sample << pre
        for (int i = tid; i < TN; i += IP_THREADS) {
            const uint32_t wd = (i < nv) ? __ldg(p.in + i0 + i) : 0u;
            ip_smem[32 + i] = (int) (short) (wd & 0xffffu) << p.pre;
            ip_smem[ip_len(0, TN) + 32 + i] = ((int) wd >> 16) << p.pre;
        }
        __syncthreads();
        IpLevels<L, 0>::run(p, ip_smem, nv, i0 << L, store);
        // slide: the newest hist_l samples of every level become the next tile's history (read, barrier, write: they may overlap)
        int keep = 0, kdst = -1;
        if (tid < IP_LEVELS * 2 * 32) {
            const int l = tid >> 6, c = (tid >> 5) & 1, q = tid & 31;
            if (l < L && q < ip_hist(l)) { kdst = ip_off(l, TN) + c * ip_len(l, TN) + q; keep = ip_smem[kdst + (nv << l)]; }
        }
        int keep2 = 0, kdst2 = -1;                                  // 384 entries, 256 threads: a second round
        {
            const int e = tid + IP_THREADS;
            if (e < IP_LEVELS * 2 * 32) {
                const int l = e >> 6, c = (e >> 5) & 1, q = e & 31;
                if (l < L && q < ip_hist(l)) { kdst2 = ip_off(l, TN) + c * ip_len(l, TN) + q; keep2 = ip_smem[kdst2 + (nv << l)]; }
            }
        }
        __syncthreads();
        if (kdst >= 0) ip_smem[kdst] = keep;
        if (kdst2 >= 0) ip_smem[kdst2] = keep2;
        __syncthreads();
    }
    if (t_end == p.tiles) {
        // the stream's last tile: the histories are the state the next call starts from; stages this call did not run keep theirs
        for (int e = tid; e < IP_LEVELS * 2 * 32; e += IP_THREADS) {
            const int l = e >> 6, c = (e >> 5) & 1, q = e & 31;
            p.st_out[e] = (l < L && q < ip_hist(l)) ? ip_smem[ip_off(l, TN) + c * ip_len(l, TN) + q] : p.st_in[e];
        }
    }
}

// interpolate1: buf = sample >> post1 (interpolators.h:118-127)
__global__ void interps_copy_kernel(const uint32_t* __restrict__ in, void* __restrict__ out, long long n, int post, int out_i8)
{
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const uint32_t w = in[i];
        const int re = (int) (short) (w & 0xffffu) >> post, im = ((int) w >> 16) >> post;
        if (out_i8) reinterpret_cast<uint16_t*>(out)[i] = (uint16_t) ((re & 0xff) | ((im & 0xff) << 8));
        else reinterpret_cast<uint32_t*>(out)[i] = ((uint32_t) re & 0xffffu) | ((uint32_t) im << 16);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K9: one UpChannelizer stage (order 96) over a block of calls.  U[j] = u[cb - 48 + j]: the 48 inputs the stage consumed
// last, the sample pending at its input (st_in, 49 words), then the outputs the next stage (or the modulator) produces
// during this block (in_new).  Call t of the block is stage call k = ks + t and works on consumption index dm = (t + (ks&1)) >> 1:
//     even k:  y = U[dm + 23]                      odd k:  y = wrap16( ( sum_i h[i] (U[dm + 47 - i] + U[dm + i]) ) >> 15 )
// then the lower/upper-half rotation (-+j)^(k+1).  A CTA takes 512 consumption indices (1024 calls); thread item =
// (4 consecutive indices, one component): a 52-value register window from 13 aligned 128-bit shared loads; lanes 0-15 hold
// the real parts, lanes 16-31 the imaginary parts of 16 groups, exchanged by shuffle for the rotation and the packed stores.
// ---------------------------------------------------------------------------------------------------------
constexpr int UP_THREADS = 256;
#ifndef UP_DMT_DEF
#define UP_DMT_DEF 512
#endif
constexpr int UP_DMT = UP_DMT_DEF;           // consumption indices per CTA (a multiple of 512: 16 groups of 4 per warp and trip)
static_assert(UP_DMT % 512 == 0, "a trip of the 8 warps covers 512 consumption indices");
constexpr int UP_LEN = 48 + UP_DMT + 8;      // shared array per component (560 + pad)
constexpr int UP_STATE_WORDS = 49;

struct UpParams {
    const uint32_t* st_in;     // [49] packed int16 IQ
    const uint32_t* in_new;    // [n_new]
    uint32_t* out;             // [ns]
    uint32_t* st_out;          // [49]
    int kphase;                // ks & 3
    int ns;                    // calls in this block
    int n_new;                 // consumptions in this block = ((ks & 1) + ns) >> 1
    int mode;                  // 0 centre, 1 lower half, 2 upper half
    int opq_zero;              // 0, passed as data (tx_add3)
};

__device__ __forceinline__ uint32_t up_u(const UpParams& p, int j)
{
    return j < UP_STATE_WORDS ? p.st_in[j] : (j - UP_STATE_WORDS < p.n_new ? __ldg(p.in_new + (j - UP_STATE_WORDS)) : 0u);
}

__global__ void __launch_bounds__(UP_THREADS) upchan_stage_kernel(const UpParams p)
{
    __shared__ __align__(16) int su[2][UP_LEN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int par = p.kphase & 1;
    const int dm_total = ((par + p.ns - 1) >> 1) + 1;          // consumption indices touched by the block's calls
    const int dmt0 = blockIdx.x * UP_DMT;
    // shared index a <-> U[dmt0 + a]; thread windows reach a < 4 * 128 + 52 = 564
    for (int a = tid; a < UP_LEN; a += UP_THREADS) {
        const uint32_t w = up_u(p, dmt0 + a);
        su[0][a] = (int) (short) (w & 0xffffu);
        su[1][a] = (int) w >> 16;
    }
    if (blockIdx.x == 0 && tid < UP_STATE_WORDS) p.st_out[tid] = up_u(p, p.n_new + tid);
    __syncthreads();
    const int c = lane >> 4;
    const int sg = (p.mode == 1) ? -1 : 1;
#pragma unroll 1
    for (int gb = warp * 16; gb < UP_DMT / 4; gb += (UP_THREADS / 32) * 16) {
        const int g = gb + (lane & 15);
        const int dm0 = dmt0 + 4 * g;
        if (dmt0 + 4 * gb >= dm_total) break;                   // warp-uniform: nothing left in this tile
        int w[52];
        const int4* src = reinterpret_cast<const int4*>(&su[c][4 * g]);
#pragma unroll
        for (int q = 0; q < 13; ++q) { const int4 v = src[q]; w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w; }
        int y[8];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int acc = 0;
#pragma unroll
            for (int i = 0; i < 24; ++i) acc += HbTaps<48>::h(i) * (w[47 + r - i] + w[r + i]);      // (the 3-input-add form of K8 measured 5 % slower here)
            y[2 * r] = w[23 + r];
            y[2 * r + 1] = (int) (short) (acc >> 15);              // stored into a Sample: int16 wrap
        }
        // lanes 0-15 finish calls 0..3 of the group, lanes 16-31 calls 4..7
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int give = c ? y[k] : y[4 + k];
            const int mine = c ? y[4 + k] : y[k];
            const int oth = __shfl_xor_sync(0xffffffffu, give, 16);
            int re = c ? oth : mine, im = c ? mine : oth;
            const int t = 2 * dm0 + 4 * c + k - par;               // call index within the block
            if (p.mode != 0) {
                const int r4 = (p.kphase + t + 1) & 3;             // (k + 1) mod 4 with k = ks + t; t >= -1
                const int a = re, b = im;
                if (r4 == 1) { re = -sg * b; im = sg * a; }
                else if (r4 == 2) { re = -a; im = -b; }
                else if (r4 == 3) { re = sg * b; im = -sg * a; }
            }
            if (t >= 0 && t < p.ns) p.out[t] = ((uint32_t) re & 0xffffu) | ((uint32_t) im << 16);
        }
    }
}

__global__ void copy_words_kernel2(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, long long n)
{
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) dst[i] = src[i];
}

int shifts_for(int output_bits, int log2, int* pre, int* post)
{
    // interpolation_shifts<16, OutputBits> (interpolators.h:47-102): pre = min(log2, 3); post = pre + (16 - OutputBits)
    if (log2 < 0 || log2 > 6) return -1;
    *pre = log2 < 3 ? log2 : 3;
    *post = *pre + (16 - output_bits);
    return 0;
}

} // namespace

struct b200dsp_interps {
    int device = 0; cudaStream_t stream = nullptr;
    int out_fmt = B200DSP_FMT_I16, bits = 16;
    int* d_state[2] = { nullptr, nullptr }; int cur = 0;
    uint32_t* d_in = nullptr; long long in_cap = 0;
    void* d_out = nullptr; long long out_cap = 0;        // bytes
    int sm_count = 148;
};

struct b200dsp_upchan {
    int device = 0; cudaStream_t stream = nullptr;
    std::vector<int> modes;                 // stage 0 first (output rate)
    std::vector<long long> calls;           // work() calls so far per stage
    uint32_t* d_state[2] = { nullptr, nullptr }; int cur = 0;     // [stages][49]
    std::vector<uint32_t*> d_mid; std::vector<long long> mid_cap; // per stage: outputs of the block (input of the stage above)
    uint32_t* d_src = nullptr; long long src_cap = 0;
    uint32_t* d_out = nullptr; long long out_cap = 0;
    uint32_t* d_sample_in = nullptr;        // m_sampleIn: the modulator sample pulled last; survives applyConfiguration (upchannelizer.h:110)
    int in_rate = 0, ofs = 0, out_rate = 0;
};

namespace {

int up_alloc_state(b200dsp_upchan* h)
{
    int rc;
    for (int k = 0; k < 2; ++k) { if (h->d_state[k]) cudaFree(h->d_state[k]); h->d_state[k] = nullptr; }
    for (auto p : h->d_mid) if (p) cudaFree(p);
    h->d_mid.assign(h->modes.size(), nullptr); h->mid_cap.assign(h->modes.size(), 0);
    h->calls.assign(h->modes.size(), 0);
    h->cur = 0;
    const size_t S = h->modes.size();
    if (S == 0) return 0;
    for (int k = 0; k < 2; ++k) {
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_state[k], S * UP_STATE_WORDS * 4)))) return rc;
        if ((rc = B200_CUDA_CHECK(cudaMemsetAsync(h->d_state[k], 0, S * UP_STATE_WORDS * 4, h->stream)))) return rc;
        // the last stage's pending input is m_sampleIn
        if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_state[k] + (S - 1) * UP_STATE_WORDS + 48, h->d_sample_in, 4, cudaMemcpyDeviceToDevice, h->stream)))) return rc;
    }
    return B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

// calls per stage for a block of n_out pulls: ns[0] = n_out, ns[s+1] = consumptions of stage s; ns[S] = modulator samples pulled
void up_counts(const b200dsp_upchan* h, long long n_out, std::vector<long long>& ns)
{
    const size_t S = h->modes.size();
    ns.assign(S + 1, 0);
    ns[0] = n_out;
    for (size_t s = 0; s < S; ++s) ns[s + 1] = (h->calls[s] + ns[s]) / 2 - h->calls[s] / 2;
}

} // namespace

extern "C" {

// ---- Interpolators<> -----------------------------------------------------------------------------------------------
int b200dsp_interps_create(b200dsp_interps_t** out, int out_fmt, int output_bits)
{
    if (!out) return b200_fail(B200DSP_EINVAL, "interps_create: null handle pointer");
    *out = nullptr;
    if (!((out_fmt == B200DSP_FMT_I16 && (output_bits == 12 || output_bits == 16)) || (out_fmt == B200DSP_FMT_I8 && output_bits == 8)))
        return b200_fail(B200DSP_EINVAL, "interps_create: unsupported format (I16 with 12/16 output bits, I8 with 8)");
    int rc = b200_require_device();
    if (rc) return rc;
    b200dsp_interps* h = new (std::nothrow) b200dsp_interps();
    if (!h) return b200_fail(B200DSP_ENOMEM, "interps_create: out of host memory");
    h->device = b200_current_device(); h->out_fmt = out_fmt; h->bits = output_bits;
    h->sm_count = b200_sm_count_of(h->device);
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(h->device))) || (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)))) { b200dsp_interps_destroy(h); return rc; }
    for (int k = 0; k < 2; ++k)
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_state[k], IP_STATE_WORDS * 4))) || (rc = B200_CUDA_CHECK(cudaMemset(h->d_state[k], 0, IP_STATE_WORDS * 4)))) { b200dsp_interps_destroy(h); return rc; }
    *out = h;
    return 0;
}

int b200dsp_interps_destroy(b200dsp_interps_t* h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    for (int k = 0; k < 2; ++k) if (h->d_state[k]) cudaFree(h->d_state[k]);
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_out) cudaFree(h->d_out);
    delete h;
    return 0;
}

int b200dsp_interps_reset(b200dsp_interps_t* h)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    for (int k = 0; k < 2; ++k) if ((rc = B200_CUDA_CHECK(cudaMemsetAsync(h->d_state[k], 0, IP_STATE_WORDS * 4, h->stream)))) return rc;
    return B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

int64_t b200dsp_interps_in_count(int log2_interp, int64_t len_scalars)
{
    if (log2_interp < 0 || log2_interp > 6 || len_scalars < 0) return -1;
    return len_scalars / (2ll << log2_interp);           // the reference loops run while pos + 2N <= len
}

int b200dsp_interps_run_dev(b200dsp_interps_t* h, int log2_interp, const void* d_samples, void* d_buf, int64_t len_scalars, int64_t* n_consumed, void* cuda_stream)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    int pre = 0, post = 0;
    if (shifts_for(h->bits, log2_interp, &pre, &post)) return b200_fail(B200DSP_EINVAL, "interps_run: log2_interp %d not in 0..6", log2_interp);
    if (len_scalars < 0) return b200_fail(B200DSP_EINVAL, "interps_run: negative length");
    const long long n = len_scalars / (2ll << log2_interp);
    if (n_consumed) *n_consumed = n;
    if (n == 0) return 0;
    if (!d_samples || !d_buf) return b200_fail(B200DSP_EINVAL, "interps_run: null buffer");
    if (((uintptr_t) d_samples & 3) || ((uintptr_t) d_buf & 15)) return b200_fail(B200DSP_EINVAL, "interps_run: device buffers must be aligned (samples 4, buf 16 bytes)");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : h->stream;
    const int i8 = (h->out_fmt == B200DSP_FMT_I8) ? 1 : 0;
    if (log2_interp == 0) {
        const long long blocks = (n + 255) / 256;
        interps_copy_kernel<<<(unsigned) (blocks < 8192 ? blocks : 8192), 256, 0, st>>>((const uint32_t*) d_samples, d_buf, n, post, i8);
        return B200_CUDA_CHECK(cudaGetLastError());
    }
    InterpParams p;
    memset(&p, 0, sizeof(p));
    p.in = (const uint32_t*) d_samples; p.out = d_buf; p.st_in = h->d_state[h->cur]; p.st_out = h->d_state[h->cur ^ 1];
    p.n = n; p.L = log2_interp; p.pre = pre; p.post = post; p.out_i8 = i8; p.quirk110 = (log2_interp == 6) ? 1 : 0;
    const int TN = 4096 >> log2_interp;
    const long long tiles = (n + TN - 1) / TN;
    if (tiles >= (1ll << 31)) return b200_fail(B200DSP_EINVAL, "interps_run: call too long");
    size_t smem = 0;
    for (int l = 0; l < log2_interp; ++l) smem += 2 * (size_t) ((l == 0 ? 32 : (l == 1 ? 16 : 8)) + (TN << l) + 4) * 4;
    typedef void (*ip_fn)(const InterpParams);
    const ip_fn fns[7] = { nullptr, interps_cascade_kernel<1>, interps_cascade_kernel<2>, interps_cascade_kernel<3>,
                           interps_cascade_kernel<4>, interps_cascade_kernel<5>, interps_cascade_kernel<6> };
    int per_sm = 0;         // ranges = the CTAs the GPU holds at once (4 by the launch bound, 5 where registers and shared memory allow)
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*) fns[log2_interp], IP_THREADS, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 4; }
    long long ctas = (long long) h->sm_count * per_sm;
    if (ctas > (tiles + 7) / 8) ctas = (tiles + 7) / 8;          // a range pays one warm-up tile: at least 8 tiles per CTA
    if (ctas < 1) ctas = 1;
    p.tiles = (int) tiles; p.tiles_per_cta = (int) ((tiles + ctas - 1) / ctas);
    ctas = (tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    switch (log2_interp) {
    case 1: interps_cascade_kernel<1><<<(unsigned) ctas, IP_THREADS, smem, st>>>(p); break;
    case 2: interps_cascade_kernel<2><<<(unsigned) ctas, IP_THREADS, smem, st>>>(p); break;
    case 3: interps_cascade_kernel<3><<<(unsigned) ctas, IP_THREADS, smem, st>>>(p); break;
    case 4: interps_cascade_kernel<4><<<(unsigned) ctas, IP_THREADS, smem, st>>>(p); break;
    case 5: interps_cascade_kernel<5><<<(unsigned) ctas, IP_THREADS, smem, st>>>(p); break;
    default: interps_cascade_kernel<6><<<(unsigned) ctas, IP_THREADS, smem, st>>>(p); break;
    }
    if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
    h->cur ^= 1;
    return 0;
}

int b200dsp_interps_run(b200dsp_interps_t* h, int log2_interp, const int16_t* samples_iq, void* buf, int32_t len_scalars, int32_t* n_consumed)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    if (log2_interp < 0 || log2_interp > 6 || len_scalars < 0) return b200_fail(B200DSP_EINVAL, "interps_run: bad argument");
    const long long n = len_scalars / (2ll << log2_interp);
    if (n_consumed) *n_consumed = (int32_t) n;
    if (n == 0) return 0;
    if (!samples_iq || !buf) return b200_fail(B200DSP_EINVAL, "interps_run: null buffer");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    const size_t osz = (h->out_fmt == B200DSP_FMT_I8) ? 1 : 2;
    const long long out_scalars = n * (2ll << log2_interp);
    if (h->in_cap < n) {
        if (h->d_in) cudaFree(h->d_in);
        h->d_in = nullptr; h->in_cap = 0;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_in, (size_t) n * 4)))) return rc;
        h->in_cap = n;
    }
    if (h->out_cap < (long long) (out_scalars * osz)) {
        if (h->d_out) cudaFree(h->d_out);
        h->d_out = nullptr; h->out_cap = 0;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_out, (size_t) out_scalars * osz)))) return rc;
        h->out_cap = out_scalars * osz;
    }
    int64_t nc = 0;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_in, samples_iq, (size_t) n * 4, cudaMemcpyHostToDevice, h->stream))) ||
        (rc = b200dsp_interps_run_dev(h, log2_interp, h->d_in, h->d_out, out_scalars, &nc, nullptr))) return rc;
    if (log2_interp == 6) {
        // interpolate64_cen leaves scalars 110..127 of every block of 128 untouched: copy back only what the reference writes
        rc = B200_CUDA_CHECK(cudaMemcpy2DAsync(buf, 128 * osz, h->d_out, 128 * osz, 110 * osz, (size_t) n, cudaMemcpyDeviceToHost, h->stream));
    } else {
        rc = B200_CUDA_CHECK(cudaMemcpyAsync(buf, h->d_out, (size_t) out_scalars * osz, cudaMemcpyDeviceToHost, h->stream));
    }
    if (rc) return rc;
    return B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

// ---- UpChannelizer -----------------------------------------------------------------------------------------------
int b200dsp_upchan_create(b200dsp_upchan_t** out)
{
    if (!out) return b200_fail(B200DSP_EINVAL, "upchan_create: null handle pointer");
    *out = nullptr;
    int rc = b200_require_device();
    if (rc) return rc;
    b200dsp_upchan* h = new (std::nothrow) b200dsp_upchan();
    if (!h) return b200_fail(B200DSP_ENOMEM, "upchan_create: out of host memory");
    h->device = b200_current_device();
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(h->device))) || (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_sample_in, 4))) || (rc = B200_CUDA_CHECK(cudaMemset(h->d_sample_in, 0, 4)))) { b200dsp_upchan_destroy(h); return rc; }
    *out = h;
    return 0;
}

int b200dsp_upchan_destroy(b200dsp_upchan_t* h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    for (int k = 0; k < 2; ++k) if (h->d_state[k]) cudaFree(h->d_state[k]);
    for (auto p : h->d_mid) if (p) cudaFree(p);
    if (h->d_src) cudaFree(h->d_src);
    if (h->d_out) cudaFree(h->d_out);
    if (h->d_sample_in) cudaFree(h->d_sample_in);
    delete h;
    return 0;
}

int b200dsp_upchan_set_path(b200dsp_upchan_t* h, const int* modes, int n_modes)
{
    if (!h || n_modes < 0 || n_modes > 30 || (n_modes > 0 && !modes)) return b200_fail(B200DSP_EINVAL, "upchan_set_path: bad argument");
    for (int i = 0; i < n_modes; ++i) if (modes[i] < 0 || modes[i] > 2) return b200_fail(B200DSP_EINVAL, "upchan_set_path: mode %d not in 0..2", modes[i]);
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(h->stream)))) return rc;
    h->modes.assign(modes, modes + n_modes);
    return up_alloc_state(h);                    // applyConfiguration frees and rebuilds the chain: fresh filters (upchannelizer.cpp:188-194)
}

int b200dsp_upchan_configure(b200dsp_upchan_t* h, int output_rate_hz, int requested_rate_hz, int center_offset_hz, int* in_rate_hz, int* residual_offset_hz)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    if (output_rate_hz <= 0) return b200_fail(B200DSP_EINVAL, "upchan_configure: output rate must be positive (the reference aborts the configuration)");
    // UpChannelizer::createFilterChain (upchannelizer.cpp:252-327) is DownChannelizer's selection with the same float32/double mix
    int modes[32], rate = 0, ofs = 0;
    const int S = b200dsp_filter_chain(output_rate_hz, requested_rate_hz, center_offset_hz, &rate, &ofs, modes, 32);
    if (S < 0) return S;
    int rc = b200dsp_upchan_set_path(h, modes, S);
    if (rc) return rc;
    h->out_rate = output_rate_hz; h->in_rate = rate; h->ofs = ofs;
    if (in_rate_hz) *in_rate_hz = rate;
    if (residual_offset_hz) *residual_offset_hz = ofs;
    return 0;
}

int b200dsp_upchan_path(b200dsp_upchan_t* h, int* modes, int cap)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    for (size_t i = 0; i < h->modes.size() && (int) i < cap; ++i) if (modes) modes[i] = h->modes[i];
    return (int) h->modes.size();
}

int64_t b200dsp_upchan_source_count(b200dsp_upchan_t* h, int64_t n_out)
{
    if (!h || n_out < 0) return -1;
    std::vector<long long> ns;
    up_counts(h, n_out, ns);
    return ns.back();
}

int b200dsp_upchan_pull_dev(b200dsp_upchan_t* h, const void* d_source, int64_t n_source, void* d_out, int64_t n_out, void* cuda_stream)
{
    if (!h || n_out < 0 || n_source < 0) return b200_fail(B200DSP_EINVAL, "upchan_pull: bad argument");
    if (n_out == 0) return 0;
    if (n_out >= (1ll << 30)) return b200_fail(B200DSP_EINVAL, "upchan_pull: block too long");
    std::vector<long long> ns;
    up_counts(h, n_out, ns);
    const size_t S = h->modes.size();
    if (n_source < ns[S]) return b200_fail(B200DSP_EINVAL, "upchan_pull: %lld output samples pull %lld modulator samples, %lld given", (long long) n_out, ns[S], (long long) n_source);
    if (!d_out || (ns[S] > 0 && !d_source)) return b200_fail(B200DSP_EINVAL, "upchan_pull: null buffer");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : h->stream;
    if (S == 0) {                            // no stage: the modulator's samples go straight through (upchannelizer.cpp:58-61)
        const long long blocks = (n_out + 255) / 256;
        copy_words_kernel2<<<(unsigned) (blocks < 4096 ? blocks : 4096), 256, 0, st>>>((const uint32_t*) d_source, (uint32_t*) d_out, n_out);
        return B200_CUDA_CHECK(cudaGetLastError());
    }
    for (size_t s = 1; s < S; ++s) {         // d_mid[s]: outputs of stage s in this block
        if (h->mid_cap[s] < ns[s]) {
            if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(st)))) return rc;
            if (h->d_mid[s]) cudaFree(h->d_mid[s]);
            h->d_mid[s] = nullptr; h->mid_cap[s] = 0;
            const long long cap = ns[s] + (ns[s] >> 2) + 64;
            if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_mid[s], (size_t) cap * 4)))) return rc;
            h->mid_cap[s] = cap;
        }
    }
    const uint32_t* sin = h->d_state[h->cur];
    uint32_t* sout = h->d_state[h->cur ^ 1];
    for (int s = (int) S - 1; s >= 0; --s) {
        UpParams p;
        memset(&p, 0, sizeof(p));
        p.st_in = sin + (size_t) s * UP_STATE_WORDS; p.st_out = sout + (size_t) s * UP_STATE_WORDS;
        p.in_new = (s == (int) S - 1) ? (const uint32_t*) d_source : h->d_mid[s + 1];
        p.out = (s == 0) ? (uint32_t*) d_out : h->d_mid[s];
        p.kphase = (int) (h->calls[s] & 3); p.ns = (int) ns[s]; p.n_new = (int) ns[s + 1]; p.mode = h->modes[s];
        if (ns[s] > 0) {
            const int par = p.kphase & 1;
            const int dm_total = ((par + p.ns - 1) >> 1) + 1;
            upchan_stage_kernel<<<(unsigned) ((dm_total + UP_DMT - 1) / UP_DMT), UP_THREADS, 0, st>>>(p);
        } else {
            // a stage that is not called in this block keeps its state: copy it to the other half of the ping-pong
            rc = B200_CUDA_CHECK(cudaMemcpyAsync(sout + (size_t) s * UP_STATE_WORDS, sin + (size_t) s * UP_STATE_WORDS, UP_STATE_WORDS * 4, cudaMemcpyDeviceToDevice, st));
            if (rc) return rc;
        }
        if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
    }
    if (ns[S] > 0 && (rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_sample_in, (const uint32_t*) d_source + (ns[S] - 1), 4, cudaMemcpyDeviceToDevice, st)))) return rc;
    for (size_t s = 0; s < S; ++s) h->calls[s] += ns[s];
    h->cur ^= 1;
    return 0;
}

int b200dsp_upchan_pull(b200dsp_upchan_t* h, const int16_t* source_iq, int64_t n_source, int16_t* out_iq, int64_t n_out)
{
    if (!h || n_out < 0 || n_source < 0) return b200_fail(B200DSP_EINVAL, "upchan_pull: bad argument");
    if (n_out == 0) return 0;
    std::vector<long long> ns;
    up_counts(h, n_out, ns);
    const long long need = ns.back();
    if (n_source < need) return b200_fail(B200DSP_EINVAL, "upchan_pull: %lld output samples pull %lld modulator samples, %lld given", (long long) n_out, need, (long long) n_source);
    if (!out_iq || (need > 0 && !source_iq)) return b200_fail(B200DSP_EINVAL, "upchan_pull: null buffer");
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    if (h->src_cap < need) {
        if (h->d_src) cudaFree(h->d_src);
        h->d_src = nullptr; h->src_cap = 0;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_src, (size_t) (need + 64) * 4)))) return rc;
        h->src_cap = need + 64;
    }
    if (h->out_cap < n_out) {
        if (h->d_out) cudaFree(h->d_out);
        h->d_out = nullptr; h->out_cap = 0;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_out, (size_t) (n_out + 64) * 4)))) return rc;
        h->out_cap = n_out + 64;
    }
    if (need > 0 && (rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_src, source_iq, (size_t) need * 4, cudaMemcpyHostToDevice, h->stream)))) return rc;
    if ((rc = b200dsp_upchan_pull_dev(h, h->d_src, need, h->d_out, n_out, nullptr))) return rc;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(out_iq, h->d_out, (size_t) n_out * 4, cudaMemcpyDeviceToHost, h->stream)))) return rc;
    return B200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

} // extern "C"
