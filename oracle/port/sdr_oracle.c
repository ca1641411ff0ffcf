/* ORACLE (test infrastructure; never shipped, never on the product path).
 * Plain-C restatement of the reference hot path; see sdr_oracle.h for scope and pinning.
 * Citations are file:line under /root/reference.  Written from the closed forms of SURVEY.md
 * Appendix A, not from the reference's ring-buffer code.
 */
#define _GNU_SOURCE
#include "sdr_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <complex.h>

/* ------------------------------------------------------------------------------------------------
 * sdrbench input generators (sdrbench/mainbench.cpp:30-31,76-79): std::mt19937 default seed 5489,
 * libstdc++ uniform_int_distribution<int16>(-2048,2047) (Lemire multiply-shift on 32-bit engines: the
 * range 4096 divides 2^32 so there is never a rejection and the value is the top 12 bits) and
 * uniform_real_distribution<float>(-1,1) (generate_canonical<float,24>: one draw, float(x)/2^32).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint32_t mt[624]; int idx; } mt19937;

static void mt_seed(mt19937* g, uint32_t seed)
{
    g->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t) i;
    g->idx = 624;
}

static uint32_t mt_next(mt19937* g)
{
    if (g->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

void orc_sdrbench_gen_s16(int16_t* buf, int n_scalars)
{
    mt19937 g;
    mt_seed(&g, 5489u);
    for (int i = 0; i < n_scalars - 1; i++)
        buf[i] = (int16_t) ((int) (((uint64_t) mt_next(&g) * 4096u) >> 32) - 2048);
    if (n_scalars > 0) buf[n_scalars - 1] = 0; /* never written by sdrbench; oracle convention: 0 */
}

void orc_sdrbench_gen_f32(float* buf, int n_scalars)
{
    mt19937 g;
    mt_seed(&g, 5489u);
    for (int i = 0; i < n_scalars - 1; i++) {
        float r = (float) mt_next(&g) / 4294967296.0f;
        if (r >= 1.0f) r = nextafterf(1.0f, 0.0f);
        buf[i] = r * 2.0f + -1.0f;
    }
    if (n_scalars > 0) buf[n_scalars - 1] = 0.0f;
}

/* ------------------------------------------------------------------------------------------------
 * Half-band coefficient sets: hbfiltertraits.cpp:85-98 (order 48) and :136-153,173-190 (order 64),
 * integer = trunc(c * 4096) (hbShift = 12), float = the same design values rounded to float.
 * ---------------------------------------------------------------------------------------------- */
static const int32_t H64[16] = { -1, 2, -5, 8, -12, 17, -25, 35, -47, 64, -86, 117, -164, 244, -424, 1300 };
static const int32_t H48[12] = { -4, 7, -12, 19, -31, 48, -71, 103, -152, 236, -419, 1299 };
static const float H64F[16] = {
    -0x1.e7e85ep-12f, 0x1.75519cp-11f, -0x1.428736p-10f, 0x1.026daap-9f, -0x1.888716p-9f, 0x1.1e6afcp-8f,
    -0x1.956518p-8f, 0x1.18583ep-7f, -0x1.7d69a8p-7f, 0x1.00fc96p-6f, -0x1.59d186p-6f, 0x1.d5f9f8p-6f,
    -0x1.48769ap-5f, 0x1.e93d42p-5f, -0x1.a8bf74p-4f, 0x1.451f18p-2f };

/* rotation of stage input n by (sigma*j)^((n+1) mod 4): inthalfbandfiltereo.h:626-641 (Inf, sigma=+1),
 * :660-675 (Sup, sigma=-1), :158-206 (LowerHalf = +1), :357-405 (UpperHalf = -1). 0 = centred. */
static inline void rot_i32(int sigma, uint32_t n, int32_t* re, int32_t* im)
{
    uint32_t x = (uint32_t) *re, y = (uint32_t) *im;
    if (sigma == 0) return;
    switch (n & 3u) {
    case 0: if (sigma > 0) { *re = (int32_t) (0u - y); *im = (int32_t) x; } else { *re = (int32_t) y; *im = (int32_t) (0u - x); } break;
    case 1: *re = (int32_t) (0u - x); *im = (int32_t) (0u - y); break;
    case 2: if (sigma > 0) { *re = (int32_t) y; *im = (int32_t) (0u - x); } else { *re = (int32_t) (0u - y); *im = (int32_t) x; } break;
    default: break;
    }
}

/* ------------------------------------------------------------------------------------------------
 * HB64 integer stage: y[k] = ( sum_i h[i]*(x[2k+1-2i] + x[2k+1-62+2i]) + (x[2k+1-31] << 11) ) >> 11
 * int32 two's-complement wrap, arithmetic shift.  inthalfbandfiltereo.h:565-573,769-790,832-870.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int32_t x[2][64]; uint32_t n; } hb64i;

static void hb64i_reset(hb64i* f) { memset(f, 0, sizeof(*f)); }

/* push one (already rotated) sample; returns 1 and writes the output when the pushed index is odd */
static int hb64i_push(hb64i* f, int32_t re, int32_t im, int32_t* ore, int32_t* oim)
{
    uint32_t n = f->n++;
    f->x[0][n & 63u] = re;
    f->x[1][n & 63u] = im;
    if (!(n & 1u)) return 0;
    for (int c = 0; c < 2; c++) {
        const int32_t* x = f->x[c];
        uint32_t acc = 0;
        for (uint32_t i = 0; i < 16; i++)
            acc += (uint32_t) H64[i] * ((uint32_t) x[(n - 2 * i) & 63u] + (uint32_t) x[(n - 62 + 2 * i) & 63u]);
        acc += (uint32_t) x[(n - 31) & 63u] << 11;
        int32_t y = (int32_t) acc >> 11; /* gcc: arithmetic shift of negative values, as on the reference's targets */
        if (c == 0) *ore = y; else *oim = y;
    }
    return 1;
}

/* decimation_shifts<16,InputBits>: decimators.h:79-95 (16), :115-131 (12), :151-167 (8); index = log2 */
static const int PRE8[7]   = { 8, 7, 6, 5, 4, 3, 2 },  POST8[7]  = { 0, 0, 0, 0, 0, 0, 0 };
static const int PRE12[7]  = { 4, 3, 2, 1, 0, 0, 0 },  POST12[7] = { 0, 0, 0, 0, 0, 1, 2 };
static const int PRE16[7]  = { 0, 0, 0, 0, 0, 0, 0 },  POST16[7] = { 0, 1, 2, 3, 4, 5, 6 };

typedef struct { int bits; hb64i st[6]; } decim_ii;

void* orc_decim_ii_create(int input_bits)
{
    if (input_bits != 8 && input_bits != 12 && input_bits != 16) return 0;
    decim_ii* d = (decim_ii*) calloc(1, sizeof(decim_ii));
    d->bits = input_bits;
    for (int s = 0; s < 6; s++) hb64i_reset(&d->st[s]);
    return d;
}
void orc_decim_ii_destroy(void* h) { free(h); }

/* stage rotation signs: decimators.h:463-2584 (inf/sup), :2610-3886 (cen) */
static void stage_sigmas(int log2, int mode, int* sg)
{
    for (int s = 0; s < log2; s++) sg[s] = 0;
    if (mode == ORC_MODE_CEN || log2 == 0) return;
    int first = (mode == ORC_MODE_INF) ? +1 : -1;
    sg[0] = first;
    if (log2 == 2) sg[1] = -first;
    else for (int s = 1; s < log2 - 1; s++) sg[s] = -first;
}

/* recursive push through stages s..L-1; emits final outputs.  cnt[s] = stage-local index within THIS call:
 * the reference's rotation pattern is hard-wired per 4-sample group of a block (inthalfbandfiltereo.h:626-692),
 * so its phase restarts with every call while the filter history carries over. */
static void cascade_push(decim_ii* d, const int* sg, uint32_t* cnt, int L, int s, int32_t re, int32_t im, int post, int16_t** out)
{
    if (s == L) {
        *(*out)++ = (int16_t) (re >> post);
        *(*out)++ = (int16_t) (im >> post);
        return;
    }
    hb64i* f = &d->st[s];
    rot_i32(sg[s], cnt[s]++, &re, &im);
    int32_t yr, yi;
    if (hb64i_push(f, re, im, &yr, &yi)) cascade_push(d, sg, cnt, L, s + 1, yr, yi, post, out);
}

int orc_decim_ii_run(void* h, int log2, int mode, const int16_t* buf, int len, int16_t* out)
{
    decim_ii* d = (decim_ii*) h;
    if (log2 < 0 || log2 > 6 || mode < 0 || mode > 2) return -1;
    const int* PRE = d->bits == 8 ? PRE8 : d->bits == 12 ? PRE12 : PRE16;
    const int* POST = d->bits == 8 ? POST8 : d->bits == 12 ? POST12 : POST16;
    int16_t* o = out;
    if (log2 == 0) { /* decimators.h:344-356 */
        for (int pos = 0; pos < len - 1; pos += 2) {
            *o++ = (int16_t) ((int32_t) ((uint32_t) (int32_t) buf[pos] << PRE[0]));
            *o++ = (int16_t) ((int32_t) ((uint32_t) (int32_t) buf[pos + 1] << PRE[0]));
        }
        return (int) ((o - out) / 2);
    }
    int N = 1 << log2;
    /* block of int16 scalars consumed per loop iteration; the trailing partial block is dropped */
    int blk = (mode == ORC_MODE_CEN) ? (N >= 8 ? 2 * N : (N == 4 ? 16 : 8)) : 4 * N;
    int sg[6];
    stage_sigmas(log2, mode, sg);
    int nscal = (len / blk) * blk;
    uint32_t cnt[6] = { 0, 0, 0, 0, 0, 0 };
    for (int pos = 0; pos < nscal; pos += 2) {
        int32_t re = (int32_t) ((uint32_t) (int32_t) buf[pos] << PRE[log2]);
        int32_t im = (int32_t) ((uint32_t) (int32_t) buf[pos + 1] << PRE[log2]);
        cascade_push(d, sg, cnt, log2, 0, re, im, POST[log2], &o);
    }
    return (int) ((o - out) / 2);
}

/* 8-bit device samples: Decimators<qint32,qint8,16,8> (hackrfinputthread.h:57; decimators.h entry points with T = qint8)
 * and DecimatorsU<qint32,quint8,16,8,Shift> (decimatorsu.h:218-231,241-249: every scalar enters as (buf[i] - Shift) << pre,
 * otherwise the text of decimators.h).  Both are the <16,8> integer cascade on the widened scalars. */
typedef struct { void* ii; int is_unsigned, shift; int16_t* tmp; int cap; } decim_x8;

void* orc_decim_x8_create(int is_unsigned, int shift)
{
    decim_x8* d = (decim_x8*) calloc(1, sizeof(decim_x8));
    d->ii = orc_decim_ii_create(8);
    d->is_unsigned = is_unsigned; d->shift = is_unsigned ? shift : 0;
    return d;
}
void orc_decim_x8_destroy(void* h)
{
    decim_x8* d = (decim_x8*) h;
    if (!d) return;
    orc_decim_ii_destroy(d->ii);
    free(d->tmp);
    free(d);
}
int orc_decim_x8_run(void* h, int log2, int mode, const uint8_t* buf, int len, int16_t* out)
{
    decim_x8* d = (decim_x8*) h;
    if (len > d->cap) { free(d->tmp); d->tmp = (int16_t*) malloc((size_t) len * sizeof(int16_t)); d->cap = len; }
    for (int i = 0; i < len; i++)
        d->tmp[i] = d->is_unsigned ? (int16_t) ((int) buf[i] - d->shift) : (int16_t) (int8_t) buf[i];
    return orc_decim_ii_run(d->ii, log2, mode, d->tmp, len, out);
}

/* DSPDeviceSourceEngine::iqCorrections, DC branch (dspdevicesourceengine.cpp:175-183,254-261) with
 * MovingAverageUtil<int32_t,int64_t,1024> (util/movingaverage.h: fill-up then roll; operator T() = total / N). */
/* MovingAverageUtil<float,double,128> and <double,double,128> (util/movingaverage.h): note `m_total += sample - oldest`
 * subtracts in T (float for the power averages) before the double accumulation. */
typedef struct { float s[128]; int num; unsigned idx; double total; } mavg_f;
typedef struct { double s[128]; int num; unsigned idx; double total; } mavg_d;
static void mavg_f_push(mavg_f* m, float v)
{
    if (m->num < 128) { m->s[m->num++] = v; m->total += v; }
    else { float d = v - m->s[m->idx]; m->total += d; m->s[m->idx] = v; m->idx = (m->idx + 1) % 128; }
}
static void mavg_d_push(mavg_d* m, double v)
{
    if (m->num < 128) { m->s[m->num++] = v; m->total += v; }
    else { m->total += v - m->s[m->idx]; m->s[m->idx] = v; m->idx = (m->idx + 1) % 128; }
}

typedef struct {
    int32_t s[2][1024]; int num[2]; unsigned idx[2]; int64_t total[2];
    mavg_f avgII, avgIQ, avgII2, avgQQ2;
    mavg_d avgPhi, avgAmp;
} iqcorr;

void* orc_iqcorr_create(void) { return calloc(1, sizeof(iqcorr)); }
void  orc_iqcorr_destroy(void* h) { free(h); }
static int32_t mavg_push(iqcorr* q, int c, int32_t v)
{
    if (q->num[c] < 1024) { q->s[c][q->num[c]++] = v; q->total[c] += v; }
    else { q->total[c] += v - q->s[c][q->idx[c]]; q->s[c][q->idx[c]] = v; q->idx[c] = (q->idx[c] + 1) % 1024; }
    return (int32_t) (q->total[c] / 1024);
}
void orc_iqcorr_dc(void* h, int16_t* iq, int n)
{
    iqcorr* q = (iqcorr*) h;
    for (int i = 0; i < n; i++) {
        int32_t bi = mavg_push(q, 0, iq[2 * i]), bq = mavg_push(q, 1, iq[2 * i + 1]);
        iq[2 * i] = (int16_t) (iq[2 * i] - bi);
        iq[2 * i + 1] = (int16_t) (iq[2 * i + 1] - bq);
    }
}
/* imbalance branch, floating-point flavour (IMBALANCE_INT undefined): dspdevicesourceengine.cpp:219-252.
 * float -> qint16 stores are C conversions (truncation toward zero); values stay in range for in-range inputs. */
void orc_iqcorr_imbalance(void* h, int16_t* iq, int n)
{
    iqcorr* q = (iqcorr*) h;
    for (int i = 0; i < n; i++) {
        int32_t bi = mavg_push(q, 0, iq[2 * i]), bq = mavg_push(q, 1, iq[2 * i + 1]);
        float xi = (float) (iq[2 * i] - bi) / 32768.0f;
        float xq = (float) (iq[2 * i + 1] - bq) / 32768.0f;
        mavg_f_push(&q->avgII, xi * xi);
        mavg_f_push(&q->avgIQ, xi * xq);
        if (q->avgII.total / 128 != 0) mavg_d_push(&q->avgPhi, (q->avgIQ.total / 128) / (q->avgII.total / 128));
        float yi = xi;
        float yq = (float) ((double) xq - (q->avgPhi.total / 128) * (double) xi);
        mavg_f_push(&q->avgII2, yi * yi);
        mavg_f_push(&q->avgQQ2, yq * yq);
        if (q->avgQQ2.total / 128 != 0) mavg_d_push(&q->avgAmp, sqrt((q->avgII2.total / 128) / (q->avgQQ2.total / 128)));
        float zi = yi;
        float zq = (float) ((q->avgAmp.total / 128) * (double) yq);
        iq[2 * i] = (int16_t) (zi * 32768.0f);
        iq[2 * i + 1] = (int16_t) (zq * 32768.0f);
    }
}

/* ------------------------------------------------------------------------------------------------
 * HB64 float stage: y = ((..(0 + hF0*(a0+b0)) + hF1*(a1+b1)) ..) + 0.5f*centre, float throughout.
 * inthalfbandfiltereof.h:63-71,141-188.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float x[2][64]; uint32_t n; } hb64f;

static int hb64f_push(hb64f* f, float re, float im, float* ore, float* oim)
{
    uint32_t n = f->n++;
    f->x[0][n & 63u] = re;
    f->x[1][n & 63u] = im;
    if (!(n & 1u)) return 0;
    for (int c = 0; c < 2; c++) {
        const float* x = f->x[c];
        float acc = 0.0f;
        for (uint32_t i = 0; i < 16; i++)
            acc += (x[(n - 2 * i) & 63u] + x[(n - 62 + 2 * i) & 63u]) * H64F[i];
        acc += x[(n - 31) & 63u] * 0.5f;
        if (c == 0) *ore = acc; else *oim = acc;
    }
    return 1;
}

typedef struct { int in_fmt, out_fmt, bits; float scale_in; hb64f st[6]; } decim_f;

void* orc_decim_f_create(int in_fmt, int out_fmt, int input_bits)
{
    decim_f* d = (decim_f*) calloc(1, sizeof(decim_f));
    d->in_fmt = in_fmt;
    d->out_fmt = out_fmt;
    d->bits = input_bits;
    /* decimatorsif.cpp:19-21 */
    d->scale_in = input_bits == 8 ? 1.0f / 128.0f : input_bits == 12 ? 1.0f / 2048.0f : 1.0f / 32768.0f;
    return d;
}
void orc_decim_f_destroy(void* h) { free(h); }

static inline float f_in(const decim_f* d, const void* buf, int i)
{
    return d->in_fmt == ORC_FMT_F32 ? ((const float*) buf)[i] : (float) ((const int16_t*) buf)[i];
}

/* FI: (FixReal)(v * 32768.0) double product, truncating conversion (decimatorsfi.cpp:28-29,48-49,1168-1169;
 *     decimate1 uses the float constant SDR_RX_SCALEF: same value since *32768 is exact).
 * FF: unchanged (decimatorsff.cpp).  IF: * scaleIn (decimatorsif.h:92-93,158-159...; a power of two, exact
 *     wherever it is applied). */
static inline void f_out(const decim_f* d, void* out, int k, float re, float im)
{
    if (d->in_fmt == ORC_FMT_I16) { re *= d->scale_in; im *= d->scale_in; }
    if (d->out_fmt == ORC_FMT_I16) {
        ((int16_t*) out)[2 * k] = (int16_t) (int32_t) ((double) re * 32768.0);
        ((int16_t*) out)[2 * k + 1] = (int16_t) (int32_t) ((double) im * 32768.0);
    } else {
        ((float*) out)[2 * k] = re;
        ((float*) out)[2 * k + 1] = im;
    }
}

static void cascade_push_f(decim_f* d, int L, int s, float re, float im, void* out, int* k)
{
    if (s == L) { f_out(d, out, (*k)++, re, im); return; }
    float yr, yi;
    if (hb64f_push(&d->st[s], re, im, &yr, &yi)) cascade_push_f(d, L, s + 1, yr, yi, out, k);
}

int orc_decim_f_run(void* h, int log2, int mode, const void* buf, int len, void* out)
{
    decim_f* d = (decim_f*) h;
    if (log2 < 0 || log2 > 6 || mode < 0 || mode > 2) return -1;
    int k = 0;
    if (log2 == 0) { /* decimatorsfi.cpp:19-31 */
        for (int pos = 0; pos < len - 1; pos += 2) f_out(d, out, k++, f_in(d, buf, pos), f_in(d, buf, pos + 1));
        return k;
    }
    int N = 1 << log2;
    if (mode == ORC_MODE_CEN) { /* decimatorsfi.cpp:33-53,369-1172: 2N scalars -> 1 output */
        int blk = 2 * N, nscal = (len / blk) * blk;
        for (int pos = 0; pos < nscal; pos += 2) cascade_push_f(d, log2, 0, f_in(d, buf, pos), f_in(d, buf, pos + 1), out, &k);
        return k;
    }
#define B(i) f_in(d, buf, pos + (i))
    if (N == 2) { /* unfiltered: decimatorsfi.cpp:55-93 */
        for (int pos = 0; pos < len - 7; pos += 8) {
            if (mode == ORC_MODE_INF) {
                f_out(d, out, k++, B(0) - B(3), B(1) + B(2));
                f_out(d, out, k++, B(7) - B(4), -B(5) - B(6));
            } else {
                f_out(d, out, k++, B(1) - B(2), -B(0) - B(3));
                f_out(d, out, k++, B(6) - B(5), B(4) + B(7));
            }
        }
        return k;
    }
    /* N >= 4: unfiltered /4 rotate-and-add front-end then log2-2 half-band stages starting at m_decimator2
     * (decimatorsfi.cpp:95-367); 2N scalars -> 1 output */
    int blk = 2 * N, nscal = (len / blk) * blk;
    for (int pos = 0; pos < nscal; pos += 8) {
        float xr, yi;
        if (mode == ORC_MODE_INF) {
            xr = B(0) - B(3) + B(7) - B(4);
            yi = B(1) - B(5) + B(2) - B(6);
        } else {
            xr = B(1) - B(2) - B(5) + B(6);
            /* decimatorsfi.cpp:124,162 (N=4,8) vs :215,270,339 (N>=16): different association as written */
            yi = (N >= 16) ? B(4) + B(7) - B(0) - B(3) : -B(0) - B(3) + B(4) + B(7);
        }
        cascade_push_f(d, log2 - 2, 0, xr, yi, out, &k);
    }
#undef B
    return k;
}

/* ------------------------------------------------------------------------------------------------
 * DownChannelizer: half-band tree path selection in float32 (downchannelizer.cpp:165-177,250-287) and the
 * HB48 int16 stage (inthalfbandfiltereo.h:37-63,158-206,357-405,751-767,792-830):
 *   x'[n] = wrap16(rot(x[n]));  y[k] = wrap16(( sum_i h48[i]*(x'[2k+1-2i]+x'[2k+1-46+2i]) + (x'[2k+1-23]<<11) ) >> 11)
 * final: trunc_toward_zero(y / 2^S) (downchannelizer.cpp:78-83).  Phase and history persist across feeds.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int16_t x[2][64]; uint32_t n; int mode; } hb48;

static int hb48_push(hb48* f, int16_t re, int16_t im, int16_t* ore, int16_t* oim)
{
    uint32_t n = f->n++;
    int32_t r = re, q = im, a = r, b = q;
    int sigma = f->mode == 1 ? +1 : f->mode == 2 ? -1 : 0;
    if (sigma) {
        switch (n & 3u) {
        case 0: if (sigma > 0) { a = -q; b = r; } else { a = q; b = -r; } break;
        case 1: a = -r; b = -q; break;
        case 2: if (sigma > 0) { a = q; b = -r; } else { a = -q; b = r; } break;
        default: break;
        }
    }
    f->x[0][n & 63u] = (int16_t) a; /* (FixReal) cast: -(-32768) wraps back to -32768 */
    f->x[1][n & 63u] = (int16_t) b;
    if (!(n & 1u)) return 0;
    for (int c = 0; c < 2; c++) {
        const int16_t* x = f->x[c];
        int32_t acc = 0;
        for (uint32_t i = 0; i < 12; i++)
            acc += H48[i] * ((int32_t) x[(n - 2 * i) & 63u] + (int32_t) x[(n - 46 + 2 * i) & 63u]);
        acc += (int32_t) x[(n - 23) & 63u] << 11;
        int16_t y = (int16_t) (acc >> 11);
        if (c == 0) *ore = y; else *oim = y;
    }
    return 1;
}

typedef struct { int nstages; hb48 st[32]; } chan_t;

void* orc_chan_create(void) { return calloc(1, sizeof(chan_t)); }
void orc_chan_destroy(void* h) { free(h); }

static int contains(float ss, float se, float cs, float ce)
{
    if (se <= ss) return 0;
    if (ce <= cs) return 0;
    return (ss <= cs) && (se >= ce);
}

/* downchannelizer.cpp:250-287; float/double mixing as written there (SURVEY App. A restatement) */
static float create_chain(chan_t* c, float ss, float se, float cs, float ce)
{
    for (;;) {
        float bw = se - ss;
        float rot = bw / 4;
        float lower_end = (float) ((double) ss + (double) bw / 2.0);
        float upper_start = se - bw / 2.0f;
        if (c->nstages < 32 && contains(ss, lower_end, cs, ce)) {
            c->st[c->nstages++].mode = 1;
            se = lower_end;
            continue;
        }
        if (c->nstages < 32 && contains(upper_start, se, cs, ce)) {
            c->st[c->nstages++].mode = 2;
            ss = upper_start;
            continue;
        }
        if (c->nstages < 32 && contains(ss + rot, se - rot, cs, ce)) {
            c->st[c->nstages++].mode = 0;
            float ns = ss + rot, ne = se - rot;
            ss = ns; se = ne;
            continue;
        }
        return (float) (((double) (float) (ce - cs) / 2.0 + (double) cs) - ((double) (float) (se - ss) / 2.0 + (double) ss));
    }
}

int orc_chan_configure(void* h, int input_rate, int requested_rate, int center_offset,
                       int* out_rate, int* residual_offset, int* modes, int modes_cap)
{
    chan_t* c = (chan_t*) h;
    memset(c, 0, sizeof(*c));
    if (input_rate == 0) return 0;
    /* downchannelizer.cpp:169-171: integer divides first, then int -> Real */
    float ofs = create_chain(c, (float) (input_rate / -2), (float) (input_rate / 2),
                             (float) (center_offset - requested_rate / 2), (float) (center_offset + requested_rate / 2));
    if (out_rate) *out_rate = input_rate / (1 << c->nstages);
    if (residual_offset) *residual_offset = (int) ofs;
    for (int i = 0; i < c->nstages && i < modes_cap; i++) if (modes) modes[i] = c->st[i].mode;
    return c->nstages;
}

int orc_chan_feed(void* h, const int16_t* iq, int n, int16_t* out, int cap)
{
    chan_t* c = (chan_t*) h;
    int m = 0;
    if (c->nstages == 0) { /* downchannelizer.cpp:57-60: forwarded unchanged */
        if (n > cap) return -1;
        memcpy(out, iq, (size_t) n * 4);
        return n;
    }
    for (int i = 0; i < n; i++) {
        int16_t re = iq[2 * i], im = iq[2 * i + 1];
        int s = 0;
        for (; s < c->nstages; s++) {
            int16_t yr, yi;
            if (!hb48_push(&c->st[s], re, im, &yr, &yi)) break;
            re = yr; im = yi;
        }
        if (s == c->nstages) {
            if (m >= cap) return -1;
            out[2 * m] = (int16_t) ((int32_t) re / (1 << c->nstages));
            out[2 * m + 1] = (int16_t) ((int32_t) im / (1 << c->nstages));
            m++;
        }
    }
    return m;
}

/* ------------------------------------------------------------------------------------------------
 * NCO (nco.cpp:30-64, nco.h:43-50) and Interpolator (interpolator.cpp:21-129, interpolator.h:23-36,107-113,183-194)
 * ---------------------------------------------------------------------------------------------- */
#define ORC_PI 3.14159265358979323846

void orc_nco_table(float* t)
{
    for (int i = 0; i < 4096; i++) t[i] = (float) cos((2.0 * ORC_PI * i) / 4096);
}

int orc_nco_increment(float freq, float rate) { return (int) ((freq * 4096) / rate); }

int orc_interp_ntaps(int phase_steps, double taps_per_phase)
{
    int ntaps = (int) (taps_per_phase * phase_steps);
    if (ntaps % 2) ntaps++;
    return ntaps; /* per phase: (ntaps*phaseSteps)/phaseSteps */
}

/* taps[phase * ntaps + i]; 'rate' is the Interpolator::create sampleRate argument */
void orc_interp_taps(int phase_steps, double rate, double cutoff, double taps_per_phase, float* out)
{
    int per_phase = orc_interp_ntaps(phase_steps, taps_per_phase);
    int ntaps = per_phase * phase_steps;
    double fs = phase_steps * rate;
    float* taps = (float*) calloc((size_t) ntaps, sizeof(float));
    float* window = (float*) calloc((size_t) ntaps, sizeof(float));
    for (int n = 0; n < ntaps; n++) window[n] = (float) (0.54 - 0.46 * cos((2 * ORC_PI * n) / (ntaps - 1)));
    int M = (ntaps - 1) / 2;
    double fwT0 = 2 * ORC_PI * cutoff / fs;
    for (int n = -M; n <= M; n++) {
        if (n == 0) taps[n + M] = (float) (fwT0 / ORC_PI * window[n + M]);
        else taps[n + M] = (float) (sin(n * fwT0) / (n * ORC_PI) * window[n + M]);
    }
    double max = taps[0 + M];
    for (int n = 1; n <= M; n++) max += 2.0 * taps[n + M];
    double gain = 1.0 / max;
    for (int i = 0; i < ntaps; i++) taps[i] = (float) (taps[i] * gain);
    for (int p = 0; p < phase_steps; p++) {
        float sum = 0;
        for (int i = 0; i < per_phase; i++) { out[p * per_phase + i] = taps[i * phase_steps + p]; sum += out[p * per_phase + i]; }
        for (int i = 0; i < per_phase; i++) out[p * per_phase + i] /= sum;
    }
    free(taps);
    free(window);
}

typedef struct {
    float table[4096];
    int inc, phase;
    int ntaps, phase_steps, ptr;
    float* taps;
    float* ring; /* complex, ntaps entries */
    float distance, remain;
} frontend_t;

void* orc_frontend_create(float nco_freq, float nco_rate, int phase_steps, double interp_rate, double cutoff,
                          double taps_per_phase, float distance)
{
    frontend_t* f = (frontend_t*) calloc(1, sizeof(frontend_t));
    orc_nco_table(f->table);
    f->inc = orc_nco_increment(nco_freq, nco_rate);
    f->phase_steps = phase_steps;
    f->ntaps = orc_interp_ntaps(phase_steps, taps_per_phase);
    f->taps = (float*) calloc((size_t) f->ntaps * phase_steps, sizeof(float));
    orc_interp_taps(phase_steps, interp_rate, cutoff, taps_per_phase, f->taps);
    f->ring = (float*) calloc((size_t) f->ntaps * 2, sizeof(float));
    f->distance = distance;
    return f;
}

void orc_frontend_destroy(void* h)
{
    frontend_t* f = (frontend_t*) h;
    if (!f) return;
    free(f->taps);
    free(f->ring);
    free(f);
}

int orc_frontend_feed(void* h, const int16_t* iq, int n, float* out, int cap, int32_t* idx, int32_t* phase)
{
    frontend_t* f = (frontend_t*) h;
    int m = 0;
    for (int i = 0; i < n; i++) {
        /* NCO::nextIQ: phase advanced before lookup (nco.h:43-50, nco.cpp:60-64) */
        f->phase += f->inc;
        while (f->phase >= 4096) f->phase -= 4096;
        while (f->phase < 0) f->phase += 4096;
        float u = f->table[f->phase], v = -f->table[(f->phase + 1024) % 4096];
        float x = (float) iq[2 * i], y = (float) iq[2 * i + 1];
        float cr = x * u - y * v, ci = x * v + y * u; /* Complex *= (nfmdemod.cpp:153) */
        /* Interpolator::decimate (interpolator.h:23-36) */
        f->ptr--;
        if (f->ptr < 0) f->ptr = f->ntaps - 1;
        f->ring[2 * f->ptr] = cr;
        f->ring[2 * f->ptr + 1] = ci;
        f->remain = (float) ((double) f->remain - 1.0);
        if (f->remain >= 1.0f) continue;
        int ph = (int) floorf(f->remain * (float) f->phase_steps);
        if (ph < 0) ph = 0;
        const float* t = f->taps + (size_t) ph * f->ntaps;
        float ra = 0, ia = 0;
        int s = f->ptr;
        for (int k = 0; k < f->ntaps; k++) { /* scalar order (interpolator.h:183-194) */
            ra += t[k] * f->ring[2 * s];
            ia += t[k] * f->ring[2 * s + 1];
            s = (s + 1) % f->ntaps;
        }
        if (m >= cap) return -1;
        out[2 * m] = ra;
        out[2 * m + 1] = ia;
        if (idx) idx[m] = i;
        if (phase) phase[m] = ph;
        m++;
        f->remain += f->distance; /* nfmdemod.cpp:315 */
    }
    return m;
}

/* Interpolator::decimate / interpolate / resample on complex float input in their callers' loops (no NCO):
 * mode 0 decimate   interpolator.h:23-36 + nfmdemod.cpp:150-155,315
 * mode 1 interpolate interpolator.h:39-52 + plugins/channeltx/modnfm/nfmmod.cpp:126-133 (one call per output, fetch on consumed)
 * mode 2 resample   interpolator.h:55-76, per input "do { if (resample) { emit; remain += distance } } while (!consumed)" */
static void frontend_push(frontend_t* f, float re, float im)
{
    f->ptr--;
    if (f->ptr < 0) f->ptr = f->ntaps - 1;
    f->ring[2 * f->ptr] = re;
    f->ring[2 * f->ptr + 1] = im;
}
static void frontend_dot(frontend_t* f, float* out)
{
    int ph = (int) floorf(f->remain * (float) f->phase_steps);
    if (ph < 0) ph = 0;
    const float* t = f->taps + (size_t) ph * f->ntaps;
    float ra = 0, ia = 0;
    int s = f->ptr;
    for (int k = 0; k < f->ntaps; k++) {
        ra += t[k] * f->ring[2 * s];
        ia += t[k] * f->ring[2 * s + 1];
        s = (s + 1) % f->ntaps;
    }
    out[0] = ra; out[1] = ia;
}
int orc_interp_run(void* h, int mode, const float* cin, int n, float* out, int cap)
{
    frontend_t* f = (frontend_t*) h;
    int m = 0;
    if (mode == 0) {
        for (int i = 0; i < n; i++) {
            frontend_push(f, cin[2 * i], cin[2 * i + 1]);
            f->remain = (float) ((double) f->remain - 1.0);
            if (f->remain >= 1.0f) continue;
            if (m >= cap) return -1;
            frontend_dot(f, out + 2 * m); m++;
            f->remain += f->distance;
        }
    } else if (mode == 1) {
        int i = 0;
        for (;;) {
            if (f->remain >= 1.0f) {
                if (i >= n) break;
                frontend_push(f, cin[2 * i], cin[2 * i + 1]);
                f->remain = (float) ((double) f->remain - 1.0);
                i++;
            }
            if (m >= cap) return -1;
            frontend_dot(f, out + 2 * m); m++;
            f->remain += f->distance;
        }
    } else {
        for (int i = 0; i < n; i++) {
            int consumed = 0;
            do {
                int ok = 1;
                while (f->remain >= 1.0f) {
                    if (!consumed) {
                        frontend_push(f, cin[2 * i], cin[2 * i + 1]);
                        f->remain = (float) ((double) f->remain - 1.0);
                        consumed = 1;
                    } else { ok = 0; break; }
                }
                if (ok) {
                    if (m >= cap) return -1;
                    frontend_dot(f, out + 2 * m); m++;
                    f->remain += f->distance;
                }
            } while (!consumed);
        }
    }
    return m;
}
float orc_frontend_remain(void* h) { return ((frontend_t*) h)->remain; }
/* NCO::nextIQ n times (nco.h:40-53, nco.cpp:30-64) */
void orc_nco_block(float freq, float rate, int n, float* out)
{
    static float table[4096];
    static int init = 0;
    if (!init) { orc_nco_table(table); init = 1; }
    int inc = orc_nco_increment(freq, rate), phase = 0;
    for (int i = 0; i < n; i++) {
        phase += inc;
        while (phase >= 4096) phase -= 4096;
        while (phase < 0) phase += 4096;
        out[2 * i] = table[phase];
        out[2 * i + 1] = -table[(phase + 1024) % 4096];
    }
}

/* ------------------------------------------------------------------------------------------------
 * FFT window (fftwindow.h:52-84, fftwindow.cpp:20-52): Real(n), Real(i) arguments, double evaluation, float result
 * ---------------------------------------------------------------------------------------------- */
void orc_fft_window(int function, int n, float* w)
{
    float fn = (float) n;
    for (int k = 0; k < n; k++) {
        float i = (float) k;
        double v;
        switch (function) {
        case 0: v = (2.0 / (fn - 1.0)) * ((fn - 1.0) / 2.0 - fabs(i - (fn - 1.0) / 2.0)) * 2.0; break;                    /* Bartlett */
        case 1: v = (0.35875 - 0.48829 * cos((2.0 * ORC_PI * i) / fn) + 0.14128 * cos((4.0 * ORC_PI * i) / fn)
                     - 0.01168 * cos((6.0 * ORC_PI * i) / fn)) * 2.79; break;                                               /* BlackmanHarris */
        case 2: v = 1.0 - 1.93 * cos((2.0 * ORC_PI * i) / fn) + 1.29 * cos((4.0 * ORC_PI * i) / fn)
                    - 0.388 * cos((6.0 * ORC_PI * i) / fn) + 0.03222 * cos((8.0 * ORC_PI * i) / fn); break;                 /* Flattop */
        case 3: v = (0.54 - 0.46 * cos((2.0 * ORC_PI * i) / fn)) * 1.855; break;                                            /* Hamming */
        case 4: v = (0.5 - 0.5 * cos((2.0 * ORC_PI * i) / fn)) * 2.0; break;                                                /* Hanning */
        default: v = 1.0; break;                                                                                            /* Rectangle */
        }
        w[k] = (float) v;
    }
}

/* ------------------------------------------------------------------------------------------------
 * KissFFT forward transform, float, radix 4 then 2 (kissfft.h:44-80 twiddles/factoring, :127-170 kf_work,
 * :205-238 kf_bfly2/kf_bfly4).  Power-of-two sizes only (SpectrumVis uses 64..4096).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int n; int nstages; int radix[16], rem[16]; float* tw; } kiss_t;

static void kiss_plan(kiss_t* k, int n)
{
    k->n = n;
    k->tw = (float*) malloc(sizeof(float) * 2 * (size_t) n);
    float phinc = -2 * acosf(-1.0f) / n;
    for (int i = 0; i < n; i++) {
        float complex e = cexpf(I * (i * phinc));
        k->tw[2 * i] = crealf(e);
        k->tw[2 * i + 1] = cimagf(e);
    }
    int m = n, p = 4;
    k->nstages = 0;
    do {
        while (m % p) {
            switch (p) { case 4: p = 2; break; case 2: p = 3; break; default: p += 2; break; }
            if (p * p > m) p = m;
        }
        m /= p;
        k->radix[k->nstages] = p;
        k->rem[k->nstages] = m;
        k->nstages++;
    } while (m > 1);
}

static void kiss_work(const kiss_t* k, int stage, float* Fout, const float* f, size_t fstride)
{
    int p = k->radix[stage], m = k->rem[stage];
    if (m == 1) {
        for (int q = 0; q < p; q++) { Fout[2 * q] = f[2 * q * fstride]; Fout[2 * q + 1] = f[2 * q * fstride + 1]; }
    } else {
        for (int q = 0; q < p; q++) kiss_work(k, stage + 1, Fout + 2 * (size_t) q * m, f + 2 * q * fstride, fstride * p);
    }
    const float* tw = k->tw;
    if (p == 2) {
        for (int u = 0; u < m; u++) {
            float ar = Fout[2 * (m + u)], ai = Fout[2 * (m + u) + 1];
            float wr = tw[2 * u * fstride], wi = tw[2 * u * fstride + 1];
            float tr = ar * wr - ai * wi, ti = ar * wi + ai * wr;
            Fout[2 * (m + u)] = Fout[2 * u] - tr;
            Fout[2 * (m + u) + 1] = Fout[2 * u + 1] - ti;
            Fout[2 * u] += tr;
            Fout[2 * u + 1] += ti;
        }
    } else { /* radix 4, forward */
        for (int u = 0; u < m; u++) {
            float s0r, s0i, s1r, s1i, s2r, s2i;
            {
                float ar = Fout[2 * (u + m)], ai = Fout[2 * (u + m) + 1], wr = tw[2 * u * fstride], wi = tw[2 * u * fstride + 1];
                s0r = ar * wr - ai * wi; s0i = ar * wi + ai * wr;
            }
            {
                float ar = Fout[2 * (u + 2 * m)], ai = Fout[2 * (u + 2 * m) + 1], wr = tw[2 * u * fstride * 2], wi = tw[2 * u * fstride * 2 + 1];
                s1r = ar * wr - ai * wi; s1i = ar * wi + ai * wr;
            }
            {
                float ar = Fout[2 * (u + 3 * m)], ai = Fout[2 * (u + 3 * m) + 1], wr = tw[2 * u * fstride * 3], wi = tw[2 * u * fstride * 3 + 1];
                s2r = ar * wr - ai * wi; s2i = ar * wi + ai * wr;
            }
            float s5r = Fout[2 * u] - s1r, s5i = Fout[2 * u + 1] - s1i;
            Fout[2 * u] += s1r; Fout[2 * u + 1] += s1i;
            float s3r = s0r + s2r, s3i = s0i + s2i;
            float dr = s0r - s2r, di = s0i - s2i;
            float s4r = di, s4i = -dr; /* multiply by -j (forward) */
            Fout[2 * (u + 2 * m)] = Fout[2 * u] - s3r; Fout[2 * (u + 2 * m) + 1] = Fout[2 * u + 1] - s3i;
            Fout[2 * u] += s3r; Fout[2 * u + 1] += s3i;
            Fout[2 * (u + m)] = s5r + s4r; Fout[2 * (u + m) + 1] = s5i + s4i;
            Fout[2 * (u + 3 * m)] = s5r - s4r; Fout[2 * (u + 3 * m) + 1] = s5i - s4i;
        }
    }
}

void orc_kissfft_forward(int n, const float* in, float* out)
{
    kiss_t k;
    kiss_plan(&k, n);
    kiss_work(&k, 0, out, in, 1);
    free(k.tw);
}

/* ------------------------------------------------------------------------------------------------
 * SpectrumVis (sdrgui/dsp/spectrumvis.cpp:77-254,283-327) with MovingAverage2D<double>
 * (util/movingaverage2d.h:40-109) and FixedAverage2D<double> (util/fixedaverage2d.h:36-113).
 * Overlap 0 only (the reference's overlap arithmetic makes no progress otherwise).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    kiss_t plan; int planned;
    int n, avg_nb, avg_mode, linear, fill;
    float scalef, ofs, div, mult;
    float* win; float* buf; float* tmp; float* fout; float* power;
    double* mov_data; double* mov_sum; unsigned mov_idx;
    double* fix_sum; unsigned fix_idx;
} spectrum_t;

void* orc_spectrum_create(float scalef)
{
    spectrum_t* s = (spectrum_t*) calloc(1, sizeof(spectrum_t));
    s->scalef = scalef;
    s->mult = 10.0f / log2f(10.0f);
    orc_spectrum_configure(s, 1024, 0, 0, 0, 1, 0);
    return s;
}

static void spectrum_free(spectrum_t* s)
{
    if (s->planned) free(s->plan.tw);
    free(s->win); free(s->buf); free(s->tmp); free(s->fout); free(s->power);
    free(s->mov_data); free(s->mov_sum); free(s->fix_sum);
}

void orc_spectrum_destroy(void* h)
{
    if (!h) return;
    spectrum_free((spectrum_t*) h);
    free(h);
}

void orc_spectrum_configure(void* h, int n, int overlap_pct, unsigned avg_nb, int avg_mode, int window, int linear)
{
    spectrum_t* s = (spectrum_t*) h;
    float scalef = s->scalef, mult = s->mult;
    (void) overlap_pct;
    spectrum_free(s);
    memset(s, 0, sizeof(*s));
    s->scalef = scalef; s->mult = mult;
    if (n > 4096) n = 4096; else if (n < 64) n = 64;
    s->n = n; s->avg_nb = (int) avg_nb; s->avg_mode = avg_mode; s->linear = linear;
    kiss_plan(&s->plan, n); s->planned = 1;
    s->win = (float*) malloc(sizeof(float) * (size_t) n);
    orc_fft_window(window, n, s->win);
    s->buf = (float*) calloc((size_t) n * 2, sizeof(float));
    s->tmp = (float*) calloc((size_t) n * 2, sizeof(float));
    s->fout = (float*) calloc((size_t) n * 2, sizeof(float));
    s->power = (float*) calloc((size_t) n, sizeof(float));
    s->mov_data = (double*) calloc((size_t) n * (avg_nb ? avg_nb : 1), sizeof(double));
    s->mov_sum = (double*) calloc((size_t) n, sizeof(double));
    s->fix_sum = (double*) calloc((size_t) n, sizeof(double));
    s->ofs = 20.0f * log10f(1.0f / n);
    s->div = (float) (n * n);
}

static inline float spec_val(const spectrum_t* s, float v)
{
    return s->linear ? v / s->div : s->mult * log2f(v) + s->ofs;
}

int orc_spectrum_feed(void* h, const int16_t* iq, int cnt, int positive_only, float* frames, int cap_frames)
{
    spectrum_t* s = (spectrum_t*) h;
    int n = s->n, half = n / 2, nframes = 0, pos = 0;
    while (pos < cnt) {
        int todo = cnt - pos, need = n - s->fill;
        if (todo < need) {
            for (; pos < cnt; pos++, s->fill++) {
                s->buf[2 * s->fill] = iq[2 * pos] / s->scalef;
                s->buf[2 * s->fill + 1] = iq[2 * pos + 1] / s->scalef;
            }
            break;
        }
        for (int i = 0; i < need; i++, pos++, s->fill++) {
            s->buf[2 * s->fill] = iq[2 * pos] / s->scalef;
            s->buf[2 * s->fill + 1] = iq[2 * pos + 1] / s->scalef;
        }
        for (int i = 0; i < n; i++) { s->tmp[2 * i] = s->buf[2 * i] * s->win[i]; s->tmp[2 * i + 1] = s->buf[2 * i + 1] * s->win[i]; }
        kiss_work(&s->plan, 0, s->fout, s->tmp, 1);
        int emit_frame = 0;
        /* bin order of the reference loops: for i<half: bin i+half first, then bin i (only matters for nothing:
         * per-bin state is independent) */
        for (int b = 0; b < n; b++) {
            if (positive_only && b >= half) continue;
            float re = s->fout[2 * b], im = s->fout[2 * b + 1];
            float v = re * re + im * im;
            int ready = 1;
            float res;
            if (s->avg_mode == 1) {
                double a;
                if (s->avg_nb <= 1) a = v;
                else {
                    double first = s->mov_data[(size_t) s->mov_idx * n + b];
                    s->mov_sum[b] += ((double) v - first);
                    s->mov_data[(size_t) s->mov_idx * n + b] = v;
                    a = s->mov_sum[b] / s->avg_nb;
                }
                res = spec_val(s, (float) a);
            } else if (s->avg_mode == 2) {
                double avg;
                if (s->avg_nb <= 1) avg = v;
                else {
                    s->fix_sum[b] += v;
                    if ((int) s->fix_idx == s->avg_nb - 1) avg = s->fix_sum[b] / s->avg_nb;
                    else { ready = 0; avg = 0; }
                }
                /* spectrumvis.cpp:198,213,222: linear mode divides the LAST frame's v, not the average */
                res = s->linear ? v / s->div : s->mult * log2f((float) avg) + s->ofs;
            } else {
                res = spec_val(s, v);
            }
            if (!ready) continue;
            if (positive_only) { s->power[2 * b] = res; s->power[2 * b + 1] = res; }
            else s->power[b < half ? b + half : b - half] = res;
        }
        if (s->avg_mode == 1) {
            emit_frame = 1;
            s->mov_idx = (s->avg_nb > 0 && (int) s->mov_idx == s->avg_nb - 1) ? 0 : s->mov_idx + 1;
            if (s->avg_nb <= 1) s->mov_idx = 0;
        } else if (s->avg_mode == 2) {
            if (s->avg_nb <= 1) emit_frame = 1;
            else if ((int) s->fix_idx == s->avg_nb - 1) { s->fix_idx = 0; memset(s->fix_sum, 0, sizeof(double) * (size_t) n); emit_frame = 1; }
            else s->fix_idx++;
        } else emit_frame = 1;
        if (emit_frame) {
            if (nframes >= cap_frames) return -1;
            memcpy(frames + (size_t) nframes * n, s->power, sizeof(float) * (size_t) n);
            nframes++;
        }
        s->fill = 0;
    }
    return nframes;
}

/* ======================================================================================================
 * Tx mirror: interpolating half-bands, Interpolators<>, UpChannelizer
 * ====================================================================================================== */
/* hbfiltertraits.cpp: (int32_t) (c * (1 << hbShift)); hbShift 12 for orders 16/32/64, 16 for order 96 */
static const int32_t HI16[4] = { -21, 95, -311, 1260 };
static const int32_t HI32[8] = { -7, 15, -33, 65, -117, 207, -401, 1294 };
static const int32_t HI96[24] = { -1, 3, -6, 11, -19, 31, -47, 70, -99, 139, -189, 254, -335, 436, -563, 722, -923, 1181, -1525, 2004,
                                  -2730, 3990, -6842, 20823 };

/* IntHalfbandFilterEO1<order> as the interpolating entry points use it: m_samples ring of order/2 complex int32, m_ptr,
 * m_state (inthalfbandfiltereo1.h:625-631,823-838) */
typedef struct { int32_t ring[48][2]; int ptr, state, K, shift; const int32_t* h; } hbup;

static void hbup_init(hbup* f, int order)
{
    memset(f, 0, sizeof(*f));
    f->K = order / 2;
    f->h = order == 16 ? HI16 : order == 32 ? HI32 : order == 64 ? H64 : HI96;
    f->shift = order == 96 ? 16 : 12;
}

static void hbup_insert(hbup* f, int32_t re, int32_t im)      /* "insert sample into ring double buffer" + "advance pointer" */
{
    f->ring[f->ptr][0] = re; f->ring[f->ptr][1] = im;
    f->ptr = (f->ptr < f->K - 1) ? f->ptr + 1 : 0;
}

static void hbup_mid(const hbup* f, int32_t* re, int32_t* im)  /* m_samples[m_ptr + hbOrder/4 - 1] */
{
    const int q = (f->ptr + f->K / 2 - 1) % f->K;
    *re = f->ring[q][0]; *im = f->ring[q][1];
}

static void hbup_fir(const hbup* f, int32_t* re, int32_t* im)  /* doInterpolateFIR, inthalfbandfiltereo1.h:777-815 */
{
    uint32_t ia = 0, qa = 0;
    int a = f->ptr, b = f->ptr + f->K - 1;
    for (int i = 0; i < f->K / 2; i++) {
        ia += (uint32_t) (f->ring[a % f->K][0] + f->ring[b % f->K][0]) * (uint32_t) f->h[i];
        qa += (uint32_t) (f->ring[a % f->K][1] + f->ring[b % f->K][1]) * (uint32_t) f->h[i];
        a++; b--;
    }
    *re = (int32_t) ia >> (f->shift - 1);
    *im = (int32_t) qa >> (f->shift - 1);
}

/* ---- Interpolators<T, 16, OutputBits> (interpolators.h:104-617) ---- */
typedef struct { hbup st[6]; int bits; } interps_t;

void* orc_interps_create(int output_bits)
{
    static const int order[6] = { 64, 32, 16, 16, 16, 16 };       /* INTERPOLATORS_HB_FILTER_ORDER_FIRST / SECOND / NEXT */
    if (output_bits != 8 && output_bits != 12 && output_bits != 16) return 0;
    interps_t* d = (interps_t*) calloc(1, sizeof(interps_t));
    for (int s = 0; s < 6; s++) hbup_init(&d->st[s], order[s]);
    d->bits = output_bits;
    return d;
}
void orc_interps_destroy(void* h) { free(h); }

/* one stage input through myInterpolate (inthalfbandfiltereo1.h:601-622) and on through the later stages; every stage sees its
 * inputs in time order, which is all its ring depends on */
static void interps_push(interps_t* d, int L, int s, int32_t re, int32_t im, int32_t* blk, int* k)
{
    if (s == L) { blk[2 * *k] = re; blk[2 * *k + 1] = im; ++*k; return; }
    hbup* f = &d->st[s];
    int32_t r1, i1, r2, i2;
    hbup_insert(f, re, im);
    hbup_mid(f, &r1, &i1);
    hbup_fir(f, &r2, &i2);
    interps_push(d, L, s + 1, r1, i1, blk, k);
    interps_push(d, L, s + 1, r2, i2, blk, k);
}

/* buf: int16 (12/16 output bits) or int8 (8); len counts output scalars.  Returns the samples consumed. */
int orc_interps_run(void* h, int log2, const int16_t* iq, void* buf, int len)
{
    interps_t* d = (interps_t*) h;
    if (log2 < 0 || log2 > 6) return -1;
    const int pre = log2 < 3 ? log2 : 3;                        /* interpolation_shifts<16,*>: pre2 1, pre4 2, pre8.. 3 */
    const int post = pre + (16 - d->bits);                      /* post = pre + (SdrBits - OutputBits) */
    const int blk = 2 << log2;
    int32_t tmp[128];
    int n = 0;
    for (int pos = 0; pos + blk <= len; pos += blk, n++) {
        int k = 0;
        if (log2 == 0) { tmp[0] = iq[2 * n]; tmp[1] = iq[2 * n + 1]; }
        else interps_push(d, log2, 0, (int32_t) iq[2 * n] << pre, (int32_t) iq[2 * n + 1] << pre, tmp, &k);
        /* interpolate64_cen stores only buf[pos+0 .. pos+109] (the store list of its loop body ends there) */
        const int stores = (log2 == 6) ? 110 : blk;
        for (int j = 0; j < stores; j++) {
            if (d->bits == 8) ((int8_t*) buf)[pos + j] = (int8_t) (tmp[j] >> post);
            else ((int16_t*) buf)[pos + j] = (int16_t) (tmp[j] >> post);
        }
    }
    return n;
}

/* ---- UpChannelizer (upchannelizer.cpp:51-104,175-209,252-327) ---- */
typedef struct { hbup st[32]; int mode[32]; int16_t stage_sample[32][2]; int16_t sample_in[2]; int nstages; } upchan_t;

void* orc_upchan_create(void) { return calloc(1, sizeof(upchan_t)); }
void orc_upchan_destroy(void* h) { free(h); }

int orc_upchan_configure(void* h, int output_rate, int requested_rate, int center_offset, int* in_rate, int* residual_offset, int* modes, int modes_cap)
{
    upchan_t* u = (upchan_t*) h;
    /* applyConfiguration: freeFilterChain + createFilterChain: the same selection as DownChannelizer's, run on a scratch chan_t */
    chan_t c;
    memset(&c, 0, sizeof(c));
    if (output_rate == 0) return 0;
    float ofs = create_chain(&c, (float) (output_rate / -2), (float) (output_rate / 2),
                             (float) (center_offset - requested_rate / 2), (float) (center_offset + requested_rate / 2));
    int16_t keep_in[2] = { u->sample_in[0], u->sample_in[1] };  /* m_sampleIn is not reset by applyConfiguration */
    memset(u, 0, sizeof(*u));
    u->sample_in[0] = keep_in[0]; u->sample_in[1] = keep_in[1];
    u->nstages = c.nstages;
    for (int i = 0; i < c.nstages; i++) { u->mode[i] = c.st[i].mode; hbup_init(&u->st[i], 96); }
    if (in_rate) *in_rate = output_rate / (1 << c.nstages);
    if (residual_offset) *residual_offset = (int) ofs;
    for (int i = 0; i < c.nstages && i < modes_cap; i++) if (modes) modes[i] = c.st[i].mode;
    return c.nstages;
}

/* workInterpolateCenter / LowerHalf / UpperHalf (inthalfbandfiltereo1.h:98-127,291-355,490-554): returns 1 when `in` was consumed */
static int hbup_work(hbup* f, int mode, const int16_t* in, int16_t* out)
{
    int32_t re, im;
    const int st = f->state;
    if ((st & 1) == 0) {
        hbup_mid(f, &re, &im);
        if (mode == 0) { out[0] = (int16_t) re; out[1] = (int16_t) im; f->state = 1; return 0; }
        /* lower half: state 0 (imag, -real), state 2 (-imag, real); upper half the opposite */
        const int flip = (mode == 1) ? (st == 2) : (st == 0);
        if (!flip) { out[0] = (int16_t) im; out[1] = (int16_t) -re; }
        else       { out[0] = (int16_t) -im; out[1] = (int16_t) re; }
        f->state = st + 1;
        return 0;
    }
    hbup_fir(f, &re, &im);
    int16_t sr = (int16_t) re, si = (int16_t) im;               /* doInterpolateFIR(Sample*): setReal / setImag narrow to int16 */
    if (mode != 0 && st == 1) { sr = (int16_t) -sr; si = (int16_t) -si; }
    out[0] = sr; out[1] = si;
    hbup_insert(f, in[0], in[1]);
    f->state = (mode == 0) ? 0 : ((st + 1) & 3);
    return 1;
}

/* n_out calls of UpChannelizer::pull (upchannelizer.cpp:51-104); the modulator hands out iq[0..n_in) then zeros.
 * Returns how many modulator samples were pulled. */
int orc_upchan_pull(void* h, const int16_t* iq, int n_in, int16_t* out, int n_out)
{
    upchan_t* u = (upchan_t*) h;
    int pos = 0;
    for (int k = 0; k < n_out; k++) {
        if (u->nstages == 0) {
            if (pos < n_in) { out[2 * k] = iq[2 * pos]; out[2 * k + 1] = iq[2 * pos + 1]; } else { out[2 * k] = 0; out[2 * k + 1] = 0; }
            pos++;
            continue;
        }
        for (int s = 0; s < u->nstages; s++) {
            if (s == u->nstages - 1) {
                if (hbup_work(&u->st[s], u->mode[s], u->sample_in, u->stage_sample[s])) {
                    if (pos < n_in) { u->sample_in[0] = iq[2 * pos]; u->sample_in[1] = iq[2 * pos + 1]; } else { u->sample_in[0] = 0; u->sample_in[1] = 0; }
                    pos++;
                }
            } else if (!hbup_work(&u->st[s], u->mode[s], u->stage_sample[s + 1], u->stage_sample[s])) break;
        }
        out[2 * k] = u->stage_sample[0][0];
        out[2 * k + 1] = u->stage_sample[0][1];
    }
    return pos;
}

void orc_hb_interp_coeffs(int order, int32_t* out)
{
    const int32_t* h = order == 16 ? HI16 : order == 32 ? HI32 : order == 64 ? H64 : HI96;
    for (int i = 0; i < order / 4; i++) out[i] = h[i];
}

/* ======================================================================================================
 * demodulator back-ends after Interpolator::decimate, and the .sdriq header
 * ====================================================================================================== */
/* PhaseDiscriminators (sdrbase/dsp/phasediscri.h:26-198) */
typedef struct { float m1[2], m2[2], scaling, prev_arg; } discri_t;

void* orc_discri_create(float fm_scaling)
{
    discri_t* d = (discri_t*) calloc(1, sizeof(discri_t));
    d->scaling = fm_scaling;
    return d;
}
void orc_discri_destroy(void* h) { free(h); }

/* phasediscri.h:162-194: |error| < 0.005 */
static float atan2_approximation2(float y, float x)
{
    const float PI_FLOAT = 3.14159265f, PIBY2_FLOAT = 1.5707963f;
    if (x == 0.0f) {
        if (y > 0.0f) return PIBY2_FLOAT;
        if (y == 0.0f) return 0.0f;
        return -PIBY2_FLOAT;
    }
    float atan;
    float z = y / x;
    if (fabsf(z) < 1.0f) {
        atan = z / (1.0f + 0.28f * z * z);
        if (x < 0.0f) {
            if (y < 0.0f) return atan - PI_FLOAT;
            return atan + PI_FLOAT;
        }
    } else {
        atan = PIBY2_FLOAT - z / (z * z + 0.28f);
        if (y < 0.0f) return atan - PI_FLOAT;
    }
    return atan;
}

/* kind 0: phaseDiscriminator (:48-53), 1: phaseDiscriminatorDelta (:59-77; aux0 = magsq, aux1 = fmDev), 2: phaseDiscriminator2
 * (:84-96), 3: AMDemod::processOneSample's magnitude (plugins/channelrx/demodam/amdemod.cpp:154-156,241; aux0 = magsq) */
void orc_discri_run(void* h, int kind, const float* in, int n, float* out, float* aux0, float* aux1)
{
    discri_t* d = (discri_t*) h;
    const double PI = 3.14159265358979323846;
    for (int i = 0; i < n; i++) {
        const float re = in[2 * i], im = in[2 * i + 1];
        if (kind == 0) {
            /* d = conj(m1) * sample */
            const float dr = d->m1[0] * re + d->m1[1] * im;
            const float di = d->m1[0] * im - d->m1[1] * re;
            d->m1[0] = re; d->m1[1] = im;
            out[i] = (float) (((double) atan2f(di, dr) / PI) * (double) d->scaling);
        } else if (kind == 1) {
            const float magsq = re * re + im * im;
            const float cur = atan2_approximation2(im, re);
            float dev = (float) ((double) (cur - d->prev_arg) / PI);
            d->prev_arg = cur;
            if (dev < -1.0f) dev += 2.0f; else if (dev > 1.0f) dev -= 2.0f;
            if (aux0) aux0[i] = magsq;
            if (aux1) aux1[i] = dev;
            out[i] = dev * d->scaling;
        } else if (kind == 2) {
            const float ip = re - d->m2[0], qp = im - d->m2[1];
            const float h1 = d->m1[0] * qp, h2 = d->m1[1] * ip;
            d->m2[0] = d->m1[0]; d->m2[1] = d->m1[1];
            d->m1[0] = re; d->m1[1] = im;
            out[i] = (h1 - h2) * d->scaling;
        } else {
            const float r = re / 32768.0f, q = im / 32768.0f;
            const float magsq = r * r + q * q;
            if (aux0) aux0[i] = magsq;
            out[i] = sqrtf(magsq);
        }
    }
}

/* FileRecord::writeHeader / readHeader (sdrbase/dsp/filerecord.cpp:129-148): qint32 rate, quint64 centre frequency,
 * time_t (8 bytes) start, quint32 sample size; 24 bytes, native little-endian, no padding */
void orc_sdriq_header(int32_t rate, uint64_t center, int64_t ts, uint32_t sample_size, uint8_t* out24)
{
    memcpy(out24, &rate, 4); memcpy(out24 + 4, &center, 8); memcpy(out24 + 12, &ts, 8); memcpy(out24 + 20, &sample_size, 4);
}

/* ---- fftfilt (sdrbase/dsp/fftfilt.cpp:49-360): Fldigi's overlap-add FFT filter.  The reference's g_fft<float> (gfft.h) is
 * replaced by a plain radix-2 transform in double: agreement with the reference is to float32 rounding (~2e-7 of the block
 * maximum), which is also what its own two builds differ by. ---- */
typedef struct { int flen, flen2, inptr; double (*filter)[2]; double (*data)[2]; double (*ovl)[2]; } fftfilt_t;

static void fft_d(double (*a)[2], int n, int inverse)
{
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { double t0 = a[i][0], t1 = a[i][1]; a[i][0] = a[j][0]; a[i][1] = a[j][1]; a[j][0] = t0; a[j][1] = t1; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const double ang = (inverse ? 2.0 : -2.0) * 3.14159265358979323846 / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; k++) {
                const double wr = cos(ang * k), wi = sin(ang * k);
                double* u = a[i + k]; double* v = a[i + k + len / 2];
                const double vr = v[0] * wr - v[1] * wi, vi = v[0] * wi + v[1] * wr;
                v[0] = u[0] - vr; v[1] = u[1] - vi; u[0] += vr; u[1] += vi;
            }
    }
    if (inverse) for (int i = 0; i < n; i++) { a[i][0] /= n; a[i][1] /= n; }        /* InverseComplexFFT comes out scaled by 1/N */
}

static float ff_fsinc(float fc, int i, int len)      /* fftfilt.h:52-57 */
{
    int len2 = len / 2;
    return (i == len2) ? (float) (2.0 * fc) : (float) (sin(2 * 3.14159265358979323846 * fc * (i - len2)) / (3.14159265358979323846 * (i - len2)));
}
static float ff_blackman(int i, int len)             /* fftfilt.h:59-64 */
{
    return (float) (0.42 - 0.50 * cos(2.0 * 3.14159265358979323846 * i / len) + 0.08 * cos(4.0 * 3.14159265358979323846 * i / len));
}

void orc_fftfilt_set(void* h, int kind, float f1, float f2)
{
    fftfilt_t* f = (fftfilt_t*) h;
    memset(f->filter, 0, (size_t) f->flen * sizeof(*f->filter));
    for (int i = 0; i < f->flen2; i++) {
        float v = 0;
        if (kind == 0) {                               /* create_filter, fftfilt.cpp:107-145 */
            if (f2 != 0) v += ff_fsinc(f2, i, f->flen2);
            if (f1 != 0) v -= ff_fsinc(f1, i, f->flen2);
        } else v = ff_fsinc(f2, i, f->flen2);          /* create_dsb_filter, :148-166 */
        f->filter[i][0] = v;
    }
    if (kind == 0 && f1 != 0 && f2 < f1) f->filter[f->flen2 / 2][0] = (float) (f->filter[f->flen2 / 2][0] + 1);
    for (int i = 0; i < f->flen2; i++) f->filter[i][0] = (float) ((float) f->filter[i][0] * ff_blackman(i, f->flen2));
    fft_d(f->filter, f->flen, 0);
    float scale = 0;
    for (int i = 0; i < f->flen2; i++) { float mag = (float) hypot(f->filter[i][0], f->filter[i][1]); if (mag > scale) scale = mag; }
    if (scale != 0) for (int i = 0; i < f->flen; i++) { f->filter[i][0] /= scale; f->filter[i][1] /= scale; }
}

void* orc_fftfilt_create(int kind, float f1, float f2, int len)
{
    fftfilt_t* f = (fftfilt_t*) calloc(1, sizeof(fftfilt_t));
    f->flen = len; f->flen2 = len >> 1;
    f->filter = calloc((size_t) len, sizeof(*f->filter)); f->data = calloc((size_t) len, sizeof(*f->data)); f->ovl = calloc((size_t) len / 2, sizeof(*f->ovl));
    orc_fftfilt_set(f, kind, f1, f2);
    return f;
}
void orc_fftfilt_destroy(void* h) { fftfilt_t* f = (fftfilt_t*) h; if (f) { free(f->filter); free(f->data); free(f->ovl); free(f); } }
void orc_fftfilt_filter(void* h, float* out_c64) { fftfilt_t* f = (fftfilt_t*) h; for (int i = 0; i < f->flen; i++) { out_c64[2 * i] = (float) f->filter[i][0]; out_c64[2 * i + 1] = (float) f->filter[i][1]; } }

/* op 0 runFilt (:261-282), 1 runSSB(usb = flag & 1, getDC = flag & 2) (:285-325), 2 runDSB(getDC = flag & 2) (:328-357), per input sample */
int orc_fftfilt_run(void* h, int op, int flag, const float* in, int n, float* out, int cap)
{
    fftfilt_t* f = (fftfilt_t*) h;
    const int N = f->flen, N2 = f->flen2;
    int m = 0;
    for (int s = 0; s < n; s++) {
        f->data[f->inptr][0] = in[2 * s]; f->data[f->inptr][1] = in[2 * s + 1];
        if (++f->inptr < N2) continue;
        f->inptr = 0;
        fft_d(f->data, N, 0);
        for (int i = 0; i < N; i++) {
            double mr = f->filter[i][0], mi = f->filter[i][1];
            if (op == 1) {
                const int usb = flag & 1, dc = flag & 2;
                if (i == 0) { if (!dc) { mr = 0; mi = 0; } }
                else if (i == N2) { mr = 1; mi = 0; }                                  /* the loops run i = 1 .. flen2-1: bin flen2 is untouched */
                else if ((i < N2) != (usb != 0)) { mr = 0; mi = 0; }
            } else if (op == 2 && i == 0 && !(flag & 2)) { mr = 0; mi = 0; }
            const double r = f->data[i][0] * mr - f->data[i][1] * mi, q = f->data[i][0] * mi + f->data[i][1] * mr;
            f->data[i][0] = r; f->data[i][1] = q;
        }
        fft_d(f->data, N, 1);
        if (m + N2 > cap) return -1;
        for (int i = 0; i < N2; i++) {
            out[2 * (m + i)] = (float) (f->ovl[i][0] + f->data[i][0]); out[2 * (m + i) + 1] = (float) (f->ovl[i][1] + f->data[i][1]);
            f->ovl[i][0] = f->data[i + N2][0]; f->ovl[i][1] = f->data[i + N2][1];
        }
        m += N2;
        memset(f->data, 0, (size_t) N * sizeof(*f->data));
    }
    return m;
}
