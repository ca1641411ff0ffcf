// hb64_cascade.cuh — K1/K2: fused multi-stage order-64 half-band decimation cascade for sm_100a.
//
// Replaces (paths relative to the reference tree):
//   Decimators<qint32,qint16,16,N>::decimate{2..64}_{inf,sup,cen}   sdrbase/dsp/decimators.h:463-3886
//   DecimatorsFI/FF/IF::decimate{2..64}_cen (+ FI/FF/IF inf/sup)     sdrbase/dsp/decimatorsfi.cpp:33-1172 ...
//   IntHalfbandFilterEO<qint32,qint32,64>::myDecimate*/doFIR        sdrbase/dsp/inthalfbandfiltereo.h:565-692,832-870
//   IntHalfbandFilterEOF<64>::myDecimate/doFIR                      sdrbase/dsp/inthalfbandfiltereof.h:63-71,141-188
//
// Arithmetic, per component, stage input x, stage output y:
//   int  : y[k] = ( sum_{i<16} h[i]*(x[2k+1-2i] + x[2k+1-62+2i]) + (x[2k+1-31] << 11) ) >> 11     (int32 wrap, arithmetic >>)
//   float: y[k] = (((0 + hF[0]*(a0+b0)) + hF[1]*(a1+b1)) ... ) + 0.5f*x[2k+1-31]
// Inf/Sup stages multiply stage input n by (sigma*j)^((n+1)&3) before filtering.
//
// B200 design (not the reference's ring buffers):
//   * The stream is cut into time slices; ONE WARP owns one slice and runs ALL stages of the cascade on it, with
//     the stage hand-off buffers in shared memory private to the warp.  No block barriers, only __syncwarp();
//     10-15 independent warps per SM hide each other's latencies.  A slice other than the first starts one
//     "superphase" early with zero history and discards those outputs (FIR: finite memory => exact), the first
//     slice starts from the handle's carried filter state.
//   * Stage inputs are kept de-interleaved by component AND by sample parity ("even"/"odd" arrays): a half-band
//     output only touches odd-phase samples (32 taps) plus one even-phase centre sample, so each lane reads one
//     contiguous 44-word window with 128-bit shared loads and slides it in registers over R=12 consecutive outputs.
//     Lane stride = 12 words => the 128-bit loads and 64-bit stores are bank-conflict free.
//   * A lane = (component, 12 consecutive outputs).  16 lanes x 12 = 192 outputs per component per item; stage s
//     runs once every 2^(s-1) phases, when its input batch (384 samples) is complete, so every executed
//     instruction has all 32 lanes busy at every stage depth.
//   * Inf/Sup rotation is deferred into the consumer's coefficients: odd-phase samples only change sign
//     (alternating), even-phase (centre) samples swap component with a sign, so rotation costs zero instructions.
//   * The instruction mix per output is 16 IADD (ALU pipe) + 16 IMAD (FMA pipe) + 1 shift; the kernel is
//     issue-bound by design (DESIGN.md roofline), global traffic is one 128-bit load per 4 samples.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200dsp {

constexpr int HB_R      = 12;                 // outputs per lane per item
constexpr int HB_LPC    = 16;                 // lanes per component
constexpr int HB_BATCH  = HB_R * HB_LPC;      // 192 outputs per component per item
constexpr int HB_IN     = 2 * HB_BATCH;       // 384 stage-input samples per item
constexpr int HB_HIST   = 32;                 // history entries kept in front of each parity array
constexpr int HB_ARR    = HB_HIST + HB_BATCH; // 224 words per (component, parity) array
constexpr int HB_STAGE_WORDS = 4 * HB_ARR;    // [comp][parity][HB_ARR]
constexpr int HB_STAGE_BYTES = HB_STAGE_WORDS * 4;
constexpr int HB_MAX_STAGES  = 6;
constexpr int HB_STATE_ELEMS = 64;            // per stage, per component (canonical carried state)

enum : int { IN_I16 = 0, IN_F32 = 1, IN_I16F = 2, IN_F32_DIV4 = 3, IN_I16F_DIV4 = 4, IN_I16_PRE = 5, IN_I32 = 6, IN_I8 = 7, IN_U8 = 8 };
enum : int { DIV4_INF = 0, DIV4_SUP = 1, DIV4_SUP16 = 2 };   // /4 front-end flavours (decimatorsfi.cpp:95-367)
enum : int { OUT_I16_SHIFT = 0, OUT_I16_SCALE = 1, OUT_F32 = 2, OUT_I32 = 3 };

struct CascadeParams {
    const void* in;          // interleaved IQ, int16 or float
    void*       out;         // interleaved IQ, int16 or float
    const void* state_in;    // T[6][2][64] canonical carried state (as the reference's rings hold it)
    void*       state_out;   // T[6][2][64]
    long long   n0;          // stage-0 IQ samples this call feeds into the cascade (after a /4 front-end)
    long long   n_out;       // n0 >> L
    int         L;           // number of half-band stages, 1..6
    int         slice_sp;    // superphases per slice
    int         n_slices;
    int         pre, post;   // integer pre/post shifts (decimation_shifts<>)
    int         in_mul, in_add;   // 8-bit inputs: sample = byte * in_mul + in_add == (byte - Shift) << pre  (decimatorsu.h:241-249)
    float       out_scale;   // float output scale (IF: 1/2^(bits-1))
    int         div4;        // DIV4_* flavour for the /4 front-end loaders
    int         opq_zero, opq_one, opq_mone;   // 0, 1, -1 passed as data so ptxas keeps the 3-input / multiply forms
    int         copy_end;    // state elements [L*128, copy_end) are carried over unchanged by the last slice (stages this call does not use)
    signed char rot[8];      // rot[s], s = 1..L : 0 centred, +1 Inf/LowerHalf (+j), -1 Sup/UpperHalf (-j)
};

// order-64 half-band taps: hbfiltertraits.cpp:136-153 (trunc(c * 4096)) and :173-190 (float)
__device__ __forceinline__ constexpr int hb64_h(int i)
{
    constexpr int h[16] = { -1, 2, -5, 8, -12, 17, -25, 35, -47, 64, -86, 117, -164, 244, -424, 1300 };
    return h[i];
}
__device__ __forceinline__ constexpr float hb64_hf(int i)
{
    constexpr float h[16] = {
        (float) -0.00046530503347925404, (float) 0.00071204906245268839, (float) -0.0012303473710125559,
        (float) 0.0019716520179919018,   (float) -0.0029947484165425580, (float) 0.0043703902150498061,
        (float) -0.0061858352927315653,  (float) 0.0085554408639278122,  (float) -0.011639792444518736,
        (float) 0.015685222110674839,    (float) -0.021107083223807829,  (float) 0.028685084689002990,
        (float) -0.040095617393092191,   (float) 0.059721592320069267,   (float) -0.10369820548136352,
        (float) 0.31750143940288489 };
    return h[i];
}

template<typename T> struct vec4;
template<> struct vec4<int32_t> { using type = int4; };
template<> struct vec4<float>   { using type = float4; };
template<typename T> struct vec2;
template<> struct vec2<int32_t> { using type = int2; };
template<> struct vec2<float>   { using type = float2; };

__device__ __forceinline__ int32_t neg_wrap(int32_t v) { return (int32_t) (0u - (uint32_t) v); }
__device__ __forceinline__ float   neg_wrap(float v)   { return -v; }

// multiply (re,im) by j^q, q mod 4
template<typename T>
__device__ __forceinline__ void rot_jq(T& re, T& im, int q)
{
    T a = re, b = im;
    switch (q & 3) {
    case 1: re = neg_wrap(b); im = a; break;
    case 2: re = neg_wrap(a); im = neg_wrap(b); break;
    case 3: re = b; im = neg_wrap(a); break;
    default: break;
    }
}

// ---------------------------------------------------------------------------------------------------------
// One item: 12 consecutive outputs of one component of one stage.
//   w  : 44-word window of the component's odd-phase array, w[t] = XO[12j - 32 + t]
//   c  : 16-word window of the centre source (even-phase array, own component if ROT == 0, the OTHER component
//        otherwise), c[t] = XE[12j - 16 + t]
//   csgn: +1/-1 runtime sign of the centre term for rotated stages (depends on component and sigma)
// Output r (k = 12j + r):  XO[k-i] = w[32+r-i],  XO[k-31+i] = w[1+r+i],  XE[k-15] = c[1+r]
// ---------------------------------------------------------------------------------------------------------
// Pipe placement (B200, measured: profiles/r01_microbench2_pipes.txt): IMAD issues on the FMA-heavy pipe and IADD3 on the
// ALU pipe, each 64 lanes/clk/SM; ptxas by itself turns most 2-input adds into IMAD.IADD and overloads the heavy pipe
// (ncu r01 v1: fmaheavy 74 % busy, alu 37 %).  The pre-adds are therefore written as explicit 3-input adds with an
// opaque zero (=> IADD3, ALU pipe) except the first HB_XH taps per output, which use an opaque-one multiply-add
// (=> IMAD, heavy pipe), so both pipes carry the same load.
struct IntOpaque { int zero, one, mone; };
#ifndef HB_XH
#define HB_XH 1
#endif
__device__ __forceinline__ uint32_t add_alu(uint32_t a, uint32_t b, int z)
{
    uint32_t r;
    asm("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(z));
    return r;
}
__device__ __forceinline__ uint32_t sub_alu(uint32_t a, uint32_t b, int z)
{
    uint32_t r;
    asm("{.reg .u32 t; sub.u32 t, %1, %2; add.u32 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(z));
    return r;
}
__device__ __forceinline__ uint32_t mad_fma(uint32_t a, int m, uint32_t b)     // a*m + b with a register multiplier
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(m), "r"(b));
    return r;
}

template<bool ROT, bool EXACT>
__device__ __forceinline__ void hb64_item(const int32_t (&w)[44], const int32_t (&c)[16], int csgn, const IntOpaque& q, int32_t (&y)[HB_R])
{
#pragma unroll
    for (int r = 0; r < HB_R; ++r) {
        uint32_t acc;
        if (!ROT) {
            acc = (uint32_t) c[1 + r] << 11;
            // tap 0: h = -1  =>  acc - a - b in one 3-input add
            asm("{.reg .u32 t; sub.u32 t, %0, %1; sub.u32 %0, t, %2;}" : "+r"(acc) : "r"(w[32 + r]), "r"(w[1 + r]));
#pragma unroll
            for (int i = 1; i < 16; ++i) {
                const uint32_t a = (uint32_t) w[32 + r - i], b = (uint32_t) w[1 + r + i];
                const uint32_t s = (i <= HB_XH) ? mad_fma(a, q.one, b) : add_alu(a, b, q.zero);
                acc += (uint32_t) hb64_h(i) * s;
            }
        } else {
            // y = s_k * [ sum (-1)^i h_i (XO[k-i] - XO[k-31+i]) + csgn * 2048 * XEother[k-15] ],  s_k = (-1)^(k+1)
            const int sk = (r & 1) ? 1 : -1;
            acc = (uint32_t) c[1 + r] * (uint32_t) (csgn * sk * 2048);
            // tap 0: coefficient -s_k  =>  acc -/+ a +/- b in one 3-input add
            if (sk > 0) asm("{.reg .u32 t; sub.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(acc) : "r"(w[32 + r]), "r"(w[1 + r]));
            else        asm("{.reg .u32 t; add.u32 t, %0, %1; sub.u32 %0, t, %2;}" : "+r"(acc) : "r"(w[32 + r]), "r"(w[1 + r]));
#pragma unroll
            for (int i = 1; i < 16; ++i) {
                const int g = ((i & 1) ? -sk : sk) * hb64_h(i);
                const uint32_t a = (uint32_t) w[32 + r - i], b = (uint32_t) w[1 + r + i];
                const uint32_t d = (i <= HB_XH) ? mad_fma(b, q.mone, a) : sub_alu(a, b, q.zero);
                acc += (uint32_t) g * d;
            }
        }
        y[r] = (int32_t) acc >> 11;
    }
}

template<bool ROT, bool EXACT>
__device__ __forceinline__ void hb64_item(const float (&w)[44], const float (&c)[16], int csgn, const IntOpaque&, float (&y)[HB_R])
{
#pragma unroll
    for (int r = 0; r < HB_R; ++r) {
        float acc = 0.0f;
        const int sk = ROT ? ((r & 1) ? 1 : -1) : 1;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float g = ROT ? (((i & 1) ? -sk : sk) * hb64_hf(i)) : hb64_hf(i);
            if (EXACT) {
                const float s = ROT ? __fsub_rn(w[32 + r - i], w[1 + r + i]) : __fadd_rn(w[32 + r - i], w[1 + r + i]);
                acc = __fadd_rn(acc, __fmul_rn(s, g));
            } else {
                const float s = ROT ? (w[32 + r - i] - w[1 + r + i]) : (w[32 + r - i] + w[1 + r + i]);
                acc = fmaf(s, g, acc);
            }
        }
        const float cm = ROT ? (0.5f * (float) (csgn * sk)) : 0.5f;
        if (EXACT) acc = __fadd_rn(acc, __fmul_rn(c[1 + r], cm));
        else       acc = fmaf(c[1 + r], cm, acc);
        y[r] = acc;
    }
}

// load the two register windows of lane (comp, j) from the stage-input buffer Xin
template<typename T>
__device__ __forceinline__ void hb64_load_windows(const T* __restrict__ Xin, int comp, int j, bool rot, T (&w)[44], T (&c)[16])
{
    using V4 = typename vec4<T>::type;
    const T* xo = Xin + (comp * 2 + 1) * HB_ARR + HB_R * j;                       // odd array, position 12j
    const T* xe = Xin + ((rot ? (comp ^ 1) : comp) * 2 + 0) * HB_ARR + HB_R * j + 16;   // even array, position 12j+16
#pragma unroll
    for (int q = 0; q < 11; ++q) {
        V4 v = *reinterpret_cast<const V4*>(xo + 4 * q);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        V4 v = *reinterpret_cast<const V4*>(xe + 4 * q);
        c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
    }
}

// keep the last 64 samples of a consumed batch as the next batch's history: positions [192,224) -> [0,32) of 4 arrays
template<typename T>
__device__ __forceinline__ typename vec4<T>::type hb64_tail_load(const T* Xin, int lane)
{
    return *reinterpret_cast<const typename vec4<T>::type*>(Xin + (lane >> 3) * HB_ARR + 4 * (lane & 7) + HB_BATCH);
}
template<typename T>
__device__ __forceinline__ void hb64_tail_store(T* Xin, int lane, const typename vec4<T>::type& v)
{
    *reinterpret_cast<typename vec4<T>::type*>(Xin + (lane >> 3) * HB_ARR + 4 * (lane & 7)) = v;
}

// store 12 outputs of lane (comp, j) as the next stage's input: even k -> even array, odd k -> odd array
template<typename T>
__device__ __forceinline__ void hb64_store_next(T* Xout, int comp, int j, int fill, const T (&y)[HB_R])
{
    using V2 = typename vec2<T>::type;
    T* xe = Xout + (comp * 2 + 0) * HB_ARR + HB_HIST + fill + 6 * j;
    T* xo = xe + HB_ARR;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        V2 e, o;
        e.x = y[4 * q]; e.y = y[4 * q + 2];
        o.x = y[4 * q + 1]; o.y = y[4 * q + 3];
        *reinterpret_cast<V2*>(xe + 2 * q) = e;
        *reinterpret_cast<V2*>(xo + 2 * q) = o;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Loaders: a phase's 384 stage-0 samples go global -> registers (fetch, issued one phase AHEAD so the HBM latency
// hides behind the previous phase's arithmetic) -> X0 positions [32, 224) of the 4 arrays (store).
// Samples at or beyond n0 read as zero.  (pos, n0 are multiples of 4 for int16 input, of 2 for float input.)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int4 ldg_nc_v4(const void* p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// sign-extend the low / high half of a packed int16 pair: one PRMT each (byte selector 8|n replicates byte n's sign)
__device__ __forceinline__ int32_t sext_lo16(int32_t v)
{
    int32_t r;
    asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ int32_t sext_hi16(int32_t v)
{
    int32_t r;
    asm("prmt.b32 %0, %1, 0, 0xbb32;" : "=r"(r) : "r"(v));
    return r;
}

template<int IN, typename T> struct Loader;

// int16 IQ -> int32 (Decimators<>: value << pre, decimators.h:2866-2878).  PRE = whether a pre-shift is applied.
template<bool PRE> struct LoaderI16 {
    static constexpr int NV = 3;                // int4 registers per lane per phase
    __device__ static __forceinline__ void fetch(const CascadeParams& p, long long pos, int lane, int4 (&v)[NV])
    {
        const int32_t* in = reinterpret_cast<const int32_t*>(p.in) + pos + 4 * lane;
        if (pos + HB_IN <= p.n0) {              // warp-uniform fast path: whole batch in range
#pragma unroll
            for (int q = 0; q < NV; ++q) v[q] = ldg_nc_v4(in + 128 * q);
        } else {
#pragma unroll
            for (int q = 0; q < NV; ++q)
                v[q] = (pos + 4 * (lane + 32 * q) < p.n0) ? ldg_nc_v4(in + 128 * q) : make_int4(0, 0, 0, 0);
        }
    }
    __device__ static __forceinline__ void store(const CascadeParams& p, int32_t* X0, int lane, const int4 (&v)[NV])
    {
        const int pre = PRE ? p.pre : 0;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            int32_t* a = X0 + HB_HIST + 2 * (lane + 32 * q);
            *reinterpret_cast<int2*>(a)              = make_int2(sext_lo16(v[q].x) << pre, sext_lo16(v[q].z) << pre);
            *reinterpret_cast<int2*>(a + HB_ARR)     = make_int2(sext_lo16(v[q].y) << pre, sext_lo16(v[q].w) << pre);
            *reinterpret_cast<int2*>(a + 2 * HB_ARR) = make_int2(sext_hi16(v[q].x) << pre, sext_hi16(v[q].z) << pre);
            *reinterpret_cast<int2*>(a + 3 * HB_ARR) = make_int2(sext_hi16(v[q].y) << pre, sext_hi16(v[q].w) << pre);
        }
    }
};
template<> struct Loader<IN_I16, int32_t>     : LoaderI16<false> {};
template<> struct Loader<IN_I16_PRE, int32_t> : LoaderI16<true> {};

// 8-bit IQ -> int32: Decimators<qint32,qint8,16,8> (signed, HackRF: hackrfinputthread.h:57) and
// DecimatorsU<qint32,quint8,16,8,Shift> (unsigned minus Shift, RTL-SDR: rtlsdrthread.h:55, decimatorsu.h:241-249).
// (byte - Shift) << pre is one multiply-add per scalar: byte * 2^pre - (Shift << pre).
__device__ __forceinline__ int2 ldg_nc_v2(const void* p)
{
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
template<bool UNSIGNED, int B> __device__ __forceinline__ int32_t ext_byte(int32_t v)
{
    int32_t r;
    if (UNSIGNED) asm("prmt.b32 %0, %1, 0, %2;" : "=r"(r) : "r"(v), "n"(0x4440 + B));             // zero-extend byte B
    else          asm("prmt.b32 %0, %1, 0, %2;" : "=r"(r) : "r"(v), "n"(0x8880 + 0x1111 * B));    // sign-extend byte B
    return r;
}
template<bool UNSIGNED> struct LoaderI8 {
    static constexpr int NV = 3;                // one 8-byte load (4 IQ samples) per register set; .z/.w stay unused
    __device__ static __forceinline__ void fetch(const CascadeParams& p, long long pos, int lane, int4 (&v)[NV])
    {
        const int16_t* in = reinterpret_cast<const int16_t*>(p.in) + pos + 4 * lane;      // one IQ sample = 2 bytes
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            int2 t = make_int2(0, 0);
            if (pos + HB_IN <= p.n0 || pos + 4 * (lane + 32 * q) < p.n0) t = ldg_nc_v2(in + 128 * q);
            v[q].x = t.x; v[q].y = t.y;
        }
    }
    __device__ static __forceinline__ void store(const CascadeParams& p, int32_t* X0, int lane, const int4 (&v)[NV])
    {
        const int m = p.in_mul, c = p.in_add;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            int32_t* a = X0 + HB_HIST + 2 * (lane + 32 * q);
            *reinterpret_cast<int2*>(a)              = make_int2(ext_byte<UNSIGNED, 0>(v[q].x) * m + c, ext_byte<UNSIGNED, 0>(v[q].y) * m + c);
            *reinterpret_cast<int2*>(a + HB_ARR)     = make_int2(ext_byte<UNSIGNED, 2>(v[q].x) * m + c, ext_byte<UNSIGNED, 2>(v[q].y) * m + c);
            *reinterpret_cast<int2*>(a + 2 * HB_ARR) = make_int2(ext_byte<UNSIGNED, 1>(v[q].x) * m + c, ext_byte<UNSIGNED, 1>(v[q].y) * m + c);
            *reinterpret_cast<int2*>(a + 3 * HB_ARR) = make_int2(ext_byte<UNSIGNED, 3>(v[q].x) * m + c, ext_byte<UNSIGNED, 3>(v[q].y) * m + c);
        }
    }
};
template<> struct Loader<IN_I8, int32_t> : LoaderI8<false> {};
template<> struct Loader<IN_U8, int32_t> : LoaderI8<true> {};

// float IQ -> float (DecimatorsFI/FF *_cen, decimatorsfi.cpp:33-53,369-1172)
template<> struct Loader<IN_F32, float> {
    static constexpr int NV = 6;
    __device__ static __forceinline__ void fetch(const CascadeParams& p, long long pos, int lane, int4 (&v)[NV])
    {
        const float2* in = reinterpret_cast<const float2*>(p.in) + pos + 2 * lane;
        if (pos + HB_IN <= p.n0) {
#pragma unroll
            for (int q = 0; q < NV; ++q) v[q] = ldg_nc_v4(in + 64 * q);
        } else {
#pragma unroll
            for (int q = 0; q < NV; ++q)
                v[q] = (pos + 2 * (lane + 32 * q) < p.n0) ? ldg_nc_v4(in + 64 * q) : make_int4(0, 0, 0, 0);
        }
    }
    __device__ static __forceinline__ void store(const CascadeParams&, float* X0, int lane, const int4 (&v)[NV])
    {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            float* a = X0 + HB_HIST + (lane + 32 * q);
            a[0]          = __int_as_float(v[q].x);
            a[2 * HB_ARR] = __int_as_float(v[q].y);
            a[HB_ARR]     = __int_as_float(v[q].z);
            a[3 * HB_ARR] = __int_as_float(v[q].w);
        }
    }
};

// int32 IQ pairs -> int32: second half of a split integer cascade (the first half's raw stage outputs)
template<> struct Loader<IN_I32, int32_t> {
    static constexpr int NV = 6;
    __device__ static __forceinline__ void fetch(const CascadeParams& p, long long pos, int lane, int4 (&v)[NV])
    {
        Loader<IN_F32, float>::fetch(p, pos, lane, v);
    }
    __device__ static __forceinline__ void store(const CascadeParams&, int32_t* X0, int lane, const int4 (&v)[NV])
    {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            int32_t* a = X0 + HB_HIST + (lane + 32 * q);
            a[0] = v[q].x; a[2 * HB_ARR] = v[q].y; a[HB_ARR] = v[q].z; a[3 * HB_ARR] = v[q].w;
        }
    }
};

// int16 IQ -> float (DecimatorsIF *_cen: raw integer values enter the cascade, the 2^-k scale is applied at the
// output, which is exact; decimatorsif.h:142-163)
template<> struct Loader<IN_I16F, float> {
    static constexpr int NV = 3;
    __device__ static __forceinline__ void fetch(const CascadeParams& p, long long pos, int lane, int4 (&v)[NV])
    {
        LoaderI16<false>::fetch(p, pos, lane, v);
    }
    __device__ static __forceinline__ void store(const CascadeParams&, float* X0, int lane, const int4 (&v)[NV])
    {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            float* a = X0 + HB_HIST + 2 * (lane + 32 * q);
            *reinterpret_cast<float2*>(a)              = make_float2((float) sext_lo16(v[q].x), (float) sext_lo16(v[q].z));
            *reinterpret_cast<float2*>(a + HB_ARR)     = make_float2((float) sext_lo16(v[q].y), (float) sext_lo16(v[q].w));
            *reinterpret_cast<float2*>(a + 2 * HB_ARR) = make_float2((float) sext_hi16(v[q].x), (float) sext_hi16(v[q].z));
            *reinterpret_cast<float2*>(a + 3 * HB_ARR) = make_float2((float) sext_hi16(v[q].y), (float) sext_hi16(v[q].w));
        }
    }
};

// FI/FF/IF inf/sup: unfiltered /4 rotate-and-add of 4 input samples per stage-0 sample
// (decimatorsfi.cpp:95-367; association as written there: N=4,8 vs N>=16 differ for sup's imaginary part).
// 4x the input bytes per stage-0 sample: fetched and combined in `store` (no register prefetch, NV = 0).
template<bool FROM_I16>
struct LoaderDiv4 {
    static constexpr int NV = 1;
    __device__ static __forceinline__ void combine(int VARIANT, const float (&B)[8], float& xr, float& yi)
    {
        if (VARIANT == DIV4_INF) {
            xr = __fsub_rn(__fadd_rn(__fsub_rn(B[0], B[3]), B[7]), B[4]);
            yi = __fsub_rn(__fadd_rn(__fsub_rn(B[1], B[5]), B[2]), B[6]);
        } else {                       // sup
            xr = __fadd_rn(__fsub_rn(__fsub_rn(B[1], B[2]), B[5]), B[6]);
            if (VARIANT == DIV4_SUP16) yi = __fsub_rn(__fsub_rn(__fadd_rn(B[4], B[7]), B[0]), B[3]);
            else              yi = __fadd_rn(__fadd_rn(__fsub_rn(-B[0], B[3]), B[4]), B[7]);
        }
    }
    // the position travels in the register slot; the data is read in store()
    __device__ static __forceinline__ void fetch(const CascadeParams&, long long pos, int, int4 (&v)[NV])
    {
        v[0].x = (int) (pos & 0xffffffffll); v[0].y = (int) (pos >> 32);
    }
    __device__ static __forceinline__ void store(const CascadeParams& p, float* X0, int lane, const int4 (&v)[NV])
    {
        const long long pos = ((long long) v[0].y << 32) | (unsigned int) v[0].x;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float B[3][8];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int i = lane + 32 * (q + 3 * h);       // stage-0 sample within the batch
                const long long s = pos + i;
                if (FROM_I16) {
                    int4 w = (s < p.n0) ? ldg_nc_v4(reinterpret_cast<const int4*>(p.in) + s) : make_int4(0, 0, 0, 0);
                    B[q][0] = (float) sext_lo16(w.x); B[q][1] = (float) sext_hi16(w.x);
                    B[q][2] = (float) sext_lo16(w.y); B[q][3] = (float) sext_hi16(w.y);
                    B[q][4] = (float) sext_lo16(w.z); B[q][5] = (float) sext_hi16(w.z);
                    B[q][6] = (float) sext_lo16(w.w); B[q][7] = (float) sext_hi16(w.w);
                } else {
                    int4 v0 = make_int4(0, 0, 0, 0), v1 = v0;
                    if (s < p.n0) {
                        v0 = ldg_nc_v4(reinterpret_cast<const int4*>(p.in) + 2 * s);
                        v1 = ldg_nc_v4(reinterpret_cast<const int4*>(p.in) + 2 * s + 1);
                    }
                    B[q][0] = __int_as_float(v0.x); B[q][1] = __int_as_float(v0.y);
                    B[q][2] = __int_as_float(v0.z); B[q][3] = __int_as_float(v0.w);
                    B[q][4] = __int_as_float(v1.x); B[q][5] = __int_as_float(v1.y);
                    B[q][6] = __int_as_float(v1.z); B[q][7] = __int_as_float(v1.w);
                }
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int i = lane + 32 * (q + 3 * h);
                float xr, yi;
                combine(p.div4, B[q], xr, yi);
                float* a = X0 + (i & 1) * HB_ARR + HB_HIST + (i >> 1);
                a[0] = xr;
                a[2 * HB_ARR] = yi;
            }
        }
    }
};
template<> struct Loader<IN_F32_DIV4, float>  : LoaderDiv4<false> {};
template<> struct Loader<IN_I16F_DIV4, float> : LoaderDiv4<true> {};

// ---------------------------------------------------------------------------------------------------------
// Epilogues: final-stage outputs -> interleaved IQ in global memory.
// ---------------------------------------------------------------------------------------------------------
// x86 cvttss2si semantics (the reference's (FixReal) cast of a float/double product): out of range -> INT_MIN
__device__ __forceinline__ int32_t cvt_trunc_x86(float v)
{
    return (fabsf(v) < 2147483648.0f) ? __float2int_rz(v) : (int32_t) 0x80000000;
}

template<int OUT, typename T> struct Epilogue;

template<> struct Epilogue<OUT_I16_SHIFT, int32_t> {        // Decimators<>: (qint16)(v >> post), decimators.h:2960-2962
    static constexpr int WORDS = 1;
    __device__ static __forceinline__ int32_t conv(const CascadeParams& p, int32_t v) { return v >> p.post; }
};
template<> struct Epilogue<OUT_I16_SCALE, float> {          // DecimatorsFI: (FixReal)(v * 32768), decimatorsfi.cpp:1168-1169
    static constexpr int WORDS = 1;
    __device__ static __forceinline__ int32_t conv(const CascadeParams&, float v) { return cvt_trunc_x86(v * 32768.0f); }
};
template<> struct Epilogue<OUT_I32, int32_t> {              // first half of a split integer cascade: raw stage outputs
    static constexpr int WORDS = 2;
    __device__ static __forceinline__ int32_t conv(const CascadeParams&, int32_t v) { return v; }
};
template<> struct Epilogue<OUT_F32, float> {                // DecimatorsFF (scale 1) / DecimatorsIF (scale 1/2^(bits-1))
    static constexpr int WORDS = 2;
    __device__ static __forceinline__ int32_t conv(const CascadeParams& p, float v) { return __float_as_int(v * p.out_scale); }
};

// lane (comp 0, j) holds re of outputs 12j..12j+11, lane (comp 1, j) holds im.  After one shuffle each lane owns
// 6 complete IQ samples: comp 0 -> outputs 12j..12j+5, comp 1 -> outputs 12j+6..12j+11.
// `out` points at the warp's first (warm-up) output; krel/lo/hi are relative to it.
template<int OUT, typename T>
__device__ __forceinline__ void hb64_epilogue(const CascadeParams& p, void* out, const T (&y)[HB_R], int comp, int j,
                                              int krel, int lo, int hi)
{
    using E = Epilogue<OUT, T>;
    int32_t mine[6], other[6];
#pragma unroll
    for (int t = 0; t < 6; ++t) {
        const int32_t a = E::conv(p, y[t]), b = E::conv(p, y[6 + t]);
        const int32_t send = comp ? a : b;
        mine[t] = comp ? b : a;
        other[t] = __shfl_xor_sync(0xffffffffu, send, 16);
    }
    const int k0 = krel + HB_R * j + 6 * comp;      // first of this lane's 6 output samples
    if (k0 >= hi || k0 + 6 <= lo) return;
    const bool full = (k0 >= lo) && (k0 + 6 <= hi);
    if (E::WORDS == 1) {
        uint32_t wds[6];
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const int32_t re = comp ? other[t] : mine[t], im = comp ? mine[t] : other[t];
            wds[t] = ((uint32_t) re & 0xffffu) | ((uint32_t) im << 16);
        }
        uint32_t* o = reinterpret_cast<uint32_t*>(out) + k0;
        if (full) {
#pragma unroll
            for (int t = 0; t < 3; ++t) *reinterpret_cast<uint2*>(o + 2 * t) = make_uint2(wds[2 * t], wds[2 * t + 1]);
        } else {
#pragma unroll
            for (int t = 0; t < 6; ++t) if (k0 + t >= lo && k0 + t < hi) o[t] = wds[t];
        }
    } else {
        int2* o = reinterpret_cast<int2*>(out) + k0;
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const int32_t re = comp ? other[t] : mine[t], im = comp ? mine[t] : other[t];
            if (full || (k0 + t >= lo && k0 + t < hi)) o[t] = make_int2(re, im);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Carried state <-> warp-private history.  Canonical state (what the reference's rings hold): for stage buffer s,
// component c, t = 0..63 chronological, the ROTATED stage inputs.  Internal buffers hold UNROTATED samples.
// ---------------------------------------------------------------------------------------------------------
template<typename T>
__device__ __noinline__ void hb64_state_load(const CascadeParams& p, T* X, int lane)
{
    const T* st = reinterpret_cast<const T*>(p.state_in);
    for (int s = 0; s < p.L; ++s) {
        const int sigma = p.rot[s + 1];
        for (int t = lane; t < 64; t += 32) {
            T re = st[(s * 2 + 0) * 64 + t], im = st[(s * 2 + 1) * 64 + t];
            if (sigma) rot_jq(re, im, -sigma * ((t + 1) & 3));       // n = t - 64, phase (n+1)&3 = (t+1)&3
            T* a = X + s * HB_STAGE_WORDS + (t & 1) * HB_ARR + (t >> 1);
            a[0] = re;
            a[2 * HB_ARR] = im;
        }
    }
}

template<typename T>
__device__ __noinline__ void hb64_state_zero(const CascadeParams& p, T* X, int lane)
{
    for (int s = 0; s < p.L; ++s)
        for (int a = 0; a < 4; ++a)
            X[s * HB_STAGE_WORDS + a * HB_ARR + lane] = (T) 0;
}

// buffer s holds stage-s samples [b0 - 64, b0 + 384); write the 64 samples that end at n_s (exclusive)
template<typename T>
__device__ __noinline__ void hb64_state_save(const CascadeParams& p, const T* Xs, int s, long long b0, long long n_s, int lane)
{
    T* st = reinterpret_cast<T*>(p.state_out);
    const int sigma = p.rot[s + 1];
    for (int t = lane; t < 64; t += 32) {
        const long long idx = n_s - 64 + t;
        const int rel = (int) (idx - b0);                 // [-64, 384)
        const T* a = Xs + (rel & 1) * HB_ARR + HB_HIST + (rel >> 1);
        T re = a[0], im = a[2 * HB_ARR];
        if (sigma) rot_jq(re, im, sigma * (int) ((idx + 1) & 3));
        st[(s * 2 + 0) * 64 + t] = re;
        st[(s * 2 + 1) * 64 + t] = im;
    }
}

// ---------------------------------------------------------------------------------------------------------
// The kernel.  blockDim.x = 32 * warps; dynamic shared memory = warps * L * HB_STAGE_BYTES.
// ---------------------------------------------------------------------------------------------------------
// last slice only: when the batch of buffer b consumed at this point holds the stream's final stage-b sample, save state
template<typename T>
__device__ __noinline__ void hb64_state_check(const CascadeParams& p, const T* Xin, int b, long long pos_after, int lane)
{
    const long long done = pos_after >> b;                 // stage-b samples consumed so far (call-absolute)
    const long long n_b = p.n0 >> b;
    if (n_b > done - HB_IN && n_b <= done) hb64_state_save<T>(p, Xin, b, done - HB_IN, n_b, lane);
}

template<typename T, int IN, int OUT, bool HASROT, bool EXACT>
__global__ void __launch_bounds__(256, 2) hb64_cascade_kernel(const CascadeParams p)
{
    extern __shared__ __align__(16) unsigned char hb64_smem[];
    using LD = Loader<IN, T>;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int w = blockIdx.x * (blockDim.x >> 5) + wib;
    if (w >= p.n_slices) return;

    const int L = p.L;
    T* X = reinterpret_cast<T*>(hb64_smem) + (size_t) wib * L * HB_STAGE_WORDS;
    const int comp = lane >> 4, j = lane & 15;
    const IntOpaque opq = { p.opq_zero, p.opq_one, p.opq_mone };

    const long long U = (long long) HB_IN << (L - 1);           // stage-0 samples per superphase
    const long long sp_total = (p.n0 + U - 1) / U;
    const long long a0 = (long long) w * p.slice_sp;
    long long a1 = a0 + p.slice_sp;
    if (a1 > sp_total) a1 = sp_total;
    const bool first = (w == 0), last = (w == p.n_slices - 1);
    const long long sp_begin = first ? 0 : a0 - 1;              // one warm-up superphase, outputs discarded

    if (first) hb64_state_load<T>(p, X, lane);
    else       hb64_state_zero<T>(p, X, lane);

    // outputs relative to the first one this warp produces (warm-up included): 32-bit bookkeeping in the loop
    const long long k_first = sp_begin * HB_BATCH;
    const int lo = (int) (a0 * HB_BATCH - k_first);
    const long long hi_abs = (a1 * HB_BATCH < p.n_out) ? a1 * HB_BATCH : p.n_out;
    const int hi = (int) (hi_abs - k_first);
    void* out = reinterpret_cast<char*>(p.out) + k_first * (Epilogue<OUT, T>::WORDS * 4);
    const int nph = (int) ((a1 - sp_begin) << (L - 1));
    long long pos = sp_begin * U;

    int4 pre[LD::NV];
    LD::fetch(p, pos, lane, pre);
    for (int ph = 0; ph < nph; ++ph) {
        LD::store(p, X, lane, pre);
        pos += HB_IN;
        if (ph + 1 < nph) LD::fetch(p, pos, lane, pre);         // next phase's input: in flight during this phase's math
        __syncwarp();
        for (int s = 1; s <= L; ++s) {
            if (((ph + 1) & ((1 << (s - 1)) - 1)) != 0) break;
            T* Xin = X + (s - 1) * HB_STAGE_WORDS;
            const int sigma = HASROT ? (int) p.rot[s] : 0;
            T wv[44], cv[16], y[HB_R];
            hb64_load_windows<T>(Xin, comp, j, sigma != 0, wv, cv);
            const typename vec4<T>::type tail = hb64_tail_load<T>(Xin, lane);
            if (last) hb64_state_check<T>(p, Xin, s - 1, pos, lane);
            if (HASROT && sigma != 0) hb64_item<true, EXACT>(wv, cv, comp ? sigma : -sigma, opq, y);
            else                      hb64_item<false, EXACT>(wv, cv, 0, opq, y);
            __syncwarp();                                       // every lane has read its windows of Xin
            hb64_tail_store<T>(Xin, lane, tail);
            if (s < L) {
                const int fill = ((((ph + 1) >> (s - 1)) - 1) & 1) * (HB_BATCH / 2);
                hb64_store_next<T>(X + s * HB_STAGE_WORDS, comp, j, fill, y);
            } else {
                hb64_epilogue<OUT, T>(p, out, y, comp, j, (((ph + 1) >> (L - 1)) - 1) * HB_BATCH, lo, hi);
            }
            __syncwarp();
        }
    }
    if (last) {     // stages not used by this call keep their state
        const T* si = reinterpret_cast<const T*>(p.state_in);
        T* so = reinterpret_cast<T*>(p.state_out);
        for (int e = L * 128 + lane; e < p.copy_end; e += 32) so[e] = si[e];
    }
}

} // namespace b200dsp
