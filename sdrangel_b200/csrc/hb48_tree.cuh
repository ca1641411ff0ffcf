// hb48_tree.cuh — K3: DownChannelizer bank as a shared-prefix half-band tree (order 48, int16 stage semantics).
//
// Replaces, for a BANK of channels fed from one baseband (paths relative to the reference tree):
//   DownChannelizer::feed                                  sdrbase/dsp/downchannelizer.cpp:50-91
//   IntHalfbandFilterEO<qint32,qint32,48>::workDecimate{Center,LowerHalf,UpperHalf}(Sample*)
//                                                          sdrbase/dsp/inthalfbandfiltereo.h:37-63,158-206,357-405
//   storeSample(FixReal,FixReal) / doFIR(Sample*)          sdrbase/dsp/inthalfbandfiltereo.h:751-767,792-830
//
// Arithmetic per stage: stage input n is rotated by (sigma*j)^((n+1)&3) (sigma = +1 lower half,
// -1 upper half, none for centre) and stored as int16 (wrap); every second input emits
//   y[k] = wrap16( ( sum_{i<12} h48[i]*(x'[2k+1-2i] + x'[2k+1-46+2i]) + (x'[2k+1-23] << 11) ) >> 11 )
// and the channel output is trunc_toward_zero(y_last / 2^S) (downchannelizer.cpp:78-83).
//
// B200 design: the reference runs one independent chain per channel; here all channels of the bank share the tree of
// distinct path prefixes, so a node is computed once however many channels pass through it.  The tree is evaluated level by
// level on time chunks small enough that a level's outputs (int16 IQ) are still L2-resident when the next level reads
// them.  One launch per level: a warp owns (family = one parent node + its <=3 children, time slice), de-interleaves the
// parent's samples once into shared memory (component x parity arrays, as in K1) and produces every child from the same
// register windows.  Rotation is folded into the coefficient signs (zero instructions); the one value for which that is not
// exact (-32768, whose int16 negation wraps) is detected per batch and handled by an explicit slow path.
#pragma once
#include "hb64_cascade.cuh"

namespace b200dsp {

constexpr int TAIL_WORDS = 68;      // per parent: 64 history samples + 1 pending + pad, packed int16 IQ words

struct LevelParams {
    const uint32_t* in_base;  long long in_stride;    // parent-level buffers (packed int16 IQ), B[i] = parent sample C_before + i
    uint32_t*       out_base; long long out_stride;   // child-level buffers
    const uint32_t* tail_in;                          // [parents][TAIL_WORDS]
    uint32_t*       tail_out;                         // next call's tails (ping-pong): written by each family's last slice
    const int*      fam;                              // [n_fam][4]: parent index, child index for mode 0 (C), 1 (L), 2 (U); -1 = absent
    int n_fam;
    int n_in;          // samples consumed per parent this call (even), B[0 .. n_in)
    int in_limit;      // samples that may be read from B (>= n_in)
    int pend;          // 1: B[0] is not valid, the sample lives in tail[64]
    int wo;            // child output write offset (the child's own pending count, 0/1)
    int flip;          // (C_before / 2) & 1 : parity of the absolute pair index of B[0]
    int slices, bps;   // slices per family, batches per slice
    int opq_zero, opq_one, opq_mone;
};

__device__ __forceinline__ constexpr int hb48_h(int i)
{
    constexpr int h[12] = { -4, 7, -12, 19, -31, 48, -71, 103, -152, 236, -419, 1299 };   // hbfiltertraits.cpp:85-98
    return h[i];
}

// 12 consecutive outputs of one component.  w[t] = XO[12j-24+t] (36 words), c[t] = XE[12j-12+t] (16 words, centre source).
// Output r: XO[k-i] = w[24+r-i], XO[k-23+i] = w[1+r+i], XE[k-11] = c[1+r].
template<bool ROT, bool FLIP>
__device__ __forceinline__ void hb48_item(const int32_t (&w)[36], const int32_t (&c)[16], int csgn, const IntOpaque& q, int32_t (&y)[HB_R])
{
#pragma unroll
    for (int r = 0; r < HB_R; ++r) {
        uint32_t acc;
        if (!ROT) {
            acc = (uint32_t) c[1 + r] << 11;
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const uint32_t a = (uint32_t) w[24 + r - i], b = (uint32_t) w[1 + r + i];
                const uint32_t s = (i < HB_XH) ? mad_fma(a, q.one, b) : add_alu(a, b, q.zero);
                acc += (uint32_t) hb48_h(i) * s;
            }
        } else {
            const int sk = ((r & 1) ? 1 : -1) * (FLIP ? -1 : 1);
            acc = (uint32_t) c[1 + r] * (uint32_t) (csgn * sk * 2048);
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const int g = ((i & 1) ? -sk : sk) * hb48_h(i);
                const uint32_t a = (uint32_t) w[24 + r - i], b = (uint32_t) w[1 + r + i];
                const uint32_t d = (i < HB_XH) ? mad_fma(b, q.mone, a) : sub_alu(a, b, q.zero);
                acc += (uint32_t) g * d;
            }
        }
        y[r] = (int32_t) acc >> 11;
    }
}

// Lower-half and upper-half children of one parent share everything but the centre term: the odd-phase samples a
// half-band stage taps sit at odd indices n, where the stage rotation (+-j)^((n+1)&3) is the real factor (-1)^((n+1)/2) for
// BOTH directions; only the centre sample (even n) is multiplied by +-j.  So with F = sum_i g_i (a_i - b_i) (the rotated
// tap sum, identical for both children) and Cc = the lower-half child's centre term,
//     y_lower = (F + Cc) >> 11,   y_upper = (F - Cc) >> 11        (all modulo 2^32, exactly what hb48_item<true> computes)
// and a two-child family costs 12 subtractions + 13 multiply-adds + 2 additions per output instead of 2 x (12 + 13).
template<bool FLIP>
__device__ __forceinline__ void hb48_pair_terms(const int32_t (&w)[36], const int32_t (&co)[16], int comp, const IntOpaque& q,
                                                uint32_t (&F)[HB_R], uint32_t (&Cc)[HB_R])
{
    const int csgn = comp ? 1 : -1;            // the lower-half child (sigma = +1): hb48_item's csgn = comp ? sigma : -sigma
#pragma unroll
    for (int r = 0; r < HB_R; ++r) {
        const int sk = ((r & 1) ? 1 : -1) * (FLIP ? -1 : 1);
        Cc[r] = (uint32_t) co[1 + r] * (uint32_t) (csgn * sk * 2048);
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int g = ((i & 1) ? -sk : sk) * hb48_h(i);
            const uint32_t a = (uint32_t) w[24 + r - i], b = (uint32_t) w[1 + r + i];
            const uint32_t d = (i < HB_XH) ? mad_fma(b, q.mone, a) : sub_alu(a, b, q.zero);
            acc += (uint32_t) g * d;
        }
        F[r] = acc;
    }
}

__device__ __forceinline__ int32_t wrap16(int32_t v) { return sext_lo16(v); }

// Explicit rotation of the register windows (slow path, exact for every int16 value incl. -32768):
//   odd-phase  x'[m] = (-1)^(m+1) u[m]                       (same component)
//   even-phase x'_re[m] = -sigma (-1)^m u_im[m],  x'_im[m] = sigma (-1)^m u_re[m]   (other component)
// m = absolute pair index = relative + flip.
__device__ __forceinline__ void hb48_rotate_exact(const int32_t (&w)[36], const int32_t (&co)[16], int comp, int sigma, int flip,
                                                  int32_t (&wr)[36], int32_t (&cr)[16])
{
#pragma unroll
    for (int t = 0; t < 36; ++t) {            // relative m = 12j - 24 + t  -> parity of t (12j, 24 even)
        const bool neg = (((t + flip) & 1) == 0);
        wr[t] = neg ? wrap16(-w[t]) : w[t];
    }
    const int s0 = comp ? sigma : -sigma;     // sign for even absolute m
#pragma unroll
    for (int t = 0; t < 16; ++t) {            // relative m = 12j - 12 + t
        const int sg = (((t + flip) & 1) == 0) ? s0 : -s0;
        cr[t] = (sg < 0) ? wrap16(-co[t]) : co[t];
    }
}

__device__ __forceinline__ void hb48_load_windows(const int32_t* __restrict__ X, int comp, int j, bool need_other,
                                                  int32_t (&w)[36], int32_t (&c)[16], int32_t (&co)[16])
{
    const int32_t* xo = X + (comp * 2 + 1) * HB_ARR + HB_R * j + 8;
    const int32_t* xe = X + (comp * 2 + 0) * HB_ARR + HB_R * j + 20;
    const int32_t* xeo = X + ((comp ^ 1) * 2 + 0) * HB_ARR + HB_R * j + 20;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int4 v = *reinterpret_cast<const int4*>(xo + 4 * q);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int4 v = *reinterpret_cast<const int4*>(xe + 4 * q);
        c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
    }
    if (need_other) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int4 v = *reinterpret_cast<const int4*>(xeo + 4 * q);
            co[4 * q] = v.x; co[4 * q + 1] = v.y; co[4 * q + 2] = v.z; co[4 * q + 3] = v.w;
        }
    }
}

// odd-phase window of the own component + the centre window of the own (rot = false) or the other (rot = true) component
__device__ __forceinline__ void hb48_load_windows2(const int32_t* __restrict__ X, int comp, int j, bool rot, int32_t (&w)[36], int32_t (&c)[16])
{
    const int32_t* xo = X + (comp * 2 + 1) * HB_ARR + HB_R * j + 8;
    const int32_t* xe = X + ((rot ? (comp ^ 1) : comp) * 2 + 0) * HB_ARR + HB_R * j + 20;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int4 v = *reinterpret_cast<const int4*>(xo + 4 * q);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int4 v = *reinterpret_cast<const int4*>(xe + 4 * q);
        c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
    }
}

// Lane (comp 0, j) holds re of outputs 12j..12j+11, lane (comp 1, j) holds im.  One shuffle per output hands the im values
// to the comp-0 lanes, which pack (PRMT) and store all 12 IQ words (3 x 128-bit, 768 contiguous bytes per warp).
__device__ __forceinline__ void hb48_store_child(uint32_t* out, const int32_t (&y)[HB_R], int comp, int j, int kbase, int wo, int n_out)
{
    uint32_t wds[HB_R];
#pragma unroll
    for (int t = 0; t < HB_R; ++t) {
        const int32_t im = __shfl_down_sync(0xffffffffu, y[t], 16);
        asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(wds[t]) : "r"(y[t]), "r"(im));      // (re & 0xffff) | (im << 16): int16 wrap = the packing
    }
    const int k0 = kbase + HB_R * j;
    if (comp != 0 || k0 >= n_out) return;
    uint32_t* o = out + wo + k0;
    if (k0 + HB_R <= n_out && wo == 0) {
#pragma unroll
        for (int t = 0; t < 3; ++t) *reinterpret_cast<uint4*>(o + 4 * t) = make_uint4(wds[4 * t], wds[4 * t + 1], wds[4 * t + 2], wds[4 * t + 3]);
    } else {
#pragma unroll
        for (int t = 0; t < HB_R; ++t) if (k0 + t < n_out) o[t] = wds[t];
    }
}

__device__ __forceinline__ uint32_t has_m32768(uint32_t v)
{
    const uint32_t t = v ^ 0x80008000u;
    return (t - 0x00010001u) & ~t & 0x80008000u;
}

// rare path (a -32768 in the batch): explicit rotation with int16 wrap, then the unrotated item.  Not inlined so that its
// 52 extra live registers do not inflate the common path.
__device__ __noinline__ void hb48_slow_child(const int32_t* X, int comp, int j, int sigma, int flip, const IntOpaque& opq, int32_t (&y)[HB_R])
{
    int32_t wv[36], cv[16], co[16], wr[36], cr[16];
    hb48_load_windows(X, comp, j, true, wv, cv, co);
    hb48_rotate_exact(wv, co, comp, sigma, flip, wr, cr);
    hb48_item<false, false>(wr, cr, 0, opq, y);
}

// one work item: warp `w` = (family, time slice) of this level
__device__ __forceinline__ void hb48_level_warp(const LevelParams& p, int w, int32_t* X, int lane)
{
    const int f = w / p.slices, slice = w - f * p.slices;
    const int4 fam = reinterpret_cast<const int4*>(p.fam)[f];      // parent, child C, child L, child U
    const int comp = lane >> 4, j = lane & 15;
    const IntOpaque opq = { p.opq_zero, p.opq_one, p.opq_mone };
    const uint32_t* B = p.in_base + (long long) fam.x * p.in_stride;
    const uint32_t* tail = p.tail_in + (long long) fam.x * TAIL_WORDS;
    const int nb = (p.n_in + HB_IN - 1) / HB_IN;
    const int q0 = slice * p.bps;
    int q1 = q0 + p.bps;
    if (q1 > nb) q1 = nb;
    if (slice == p.slices - 1) {
        // carry: new_tail[t] = parent sample (C_after - 64 + t), t = 0..64 (64 history + the possibly pending sample)
        uint32_t* tout = p.tail_out + (long long) fam.x * TAIL_WORDS;
        for (int t = lane; t <= 64; t += 32) {
            const int i = p.n_in - 64 + t;
            uint32_t v;
            if (i < p.pend) v = tail[i + 64];
            else            v = (i < p.in_limit) ? B[i] : 0u;
            tout[t] = v;
        }
    }
    if (q0 >= q1) return;
    const int n_out = p.n_in >> 1;
    const bool rotated = (fam.z >= 0) || (fam.w >= 0);

    // history: the 64 samples before this slice (from the carried tail for the first slice)
    uint32_t hist_bad;
    {
        const uint32_t* src = (q0 == 0) ? tail : (B + (long long) q0 * HB_IN - 64);
        const uint32_t s0 = src[2 * lane], s1 = src[2 * lane + 1];
        hist_bad = has_m32768(s0) | has_m32768(s1);
        X[0 * HB_ARR + lane] = sext_lo16((int32_t) s0);
        X[1 * HB_ARR + lane] = sext_lo16((int32_t) s1);
        X[2 * HB_ARR + lane] = sext_hi16((int32_t) s0);
        X[3 * HB_ARR + lane] = sext_hi16((int32_t) s1);
    }

    bool prev_bad = rotated && __any_sync(0xffffffffu, hist_bad != 0);
    int4 pre[3];
    auto fetch = [&](int q) {
        const int base = q * HB_IN + 4 * lane;
        const uint32_t* src = B + base;
        if ((q + 1) * HB_IN <= p.in_limit) {
#pragma unroll
            for (int t = 0; t < 3; ++t) pre[t] = ldg_nc_v4(src + 128 * t);
        } else {
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                int v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = (base + 128 * t + e < p.in_limit) ? (int) src[128 * t + e] : 0;
                pre[t] = make_int4(v[0], v[1], v[2], v[3]);
            }
        }
        if (q == 0 && p.pend && lane == 0) pre[0].x = (int) tail[64];
    };
    fetch(q0);
    for (int q = q0; q < q1; ++q) {
        uint32_t bad = 0;
        if (rotated) {
#pragma unroll
            for (int t = 0; t < 3; ++t)
                bad |= has_m32768((uint32_t) pre[t].x) | has_m32768((uint32_t) pre[t].y) | has_m32768((uint32_t) pre[t].z) | has_m32768((uint32_t) pre[t].w);
        }
        CascadeParams dummy;   // LoaderI16<false>::store does not read params
        LoaderI16<false>::store(dummy, X, lane, pre);
        if (q + 1 < q1) fetch(q + 1);
        // the history region holds the previous batch's tail: keep the flag for one more batch
        const bool bad_now = rotated && __any_sync(0xffffffffu, bad != 0);
        const bool slow = bad_now || prev_bad;
        prev_bad = bad_now;
        __syncwarp();
        int32_t wv[36], cv[16], co[16], y[HB_R];
        hb48_load_windows(X, comp, j, rotated, wv, cv, co);
        const int4 tl = hb64_tail_load<int32_t>(X, lane);
        const int kbase = q * HB_BATCH;
        if (fam.y >= 0) {
            hb48_item<false, false>(wv, cv, 0, opq, y);
            hb48_store_child(p.out_base + (long long) fam.y * p.out_stride, y, comp, j, kbase, p.wo, n_out);
        }
        if (fam.z >= 0 && fam.w >= 0 && !slow) {
            // both rotated children: one shared tap sum, the centre term added for the lower half, subtracted for the upper
            uint32_t F[HB_R], Cc[HB_R];
            if (p.flip) hb48_pair_terms<true>(wv, co, comp, opq, F, Cc);
            else        hb48_pair_terms<false>(wv, co, comp, opq, F, Cc);
#pragma unroll
            for (int r = 0; r < HB_R; ++r) y[r] = (int32_t) add_alu(F[r], Cc[r], opq.zero) >> 11;
            hb48_store_child(p.out_base + (long long) fam.z * p.out_stride, y, comp, j, kbase, p.wo, n_out);
#pragma unroll
            for (int r = 0; r < HB_R; ++r) y[r] = (int32_t) sub_alu(F[r], Cc[r], opq.zero) >> 11;
            hb48_store_child(p.out_base + (long long) fam.w * p.out_stride, y, comp, j, kbase, p.wo, n_out);
        } else {
#pragma unroll
            for (int m = 1; m <= 2; ++m) {          // lower-half (+j) and upper-half (-j) children
                const int child = (m == 1) ? fam.z : fam.w;
                if (child < 0) continue;
                const int sigma = (m == 1) ? 1 : -1;
                if (!slow) {
                    if (p.flip) hb48_item<true, true>(wv, co, comp ? sigma : -sigma, opq, y);
                    else        hb48_item<true, false>(wv, co, comp ? sigma : -sigma, opq, y);
                } else {
                    hb48_slow_child(X, comp, j, sigma, p.flip, opq, y);
                }
                hb48_store_child(p.out_base + (long long) child * p.out_stride, y, comp, j, kbase, p.wo, n_out);
            }
        }
        __syncwarp();
        hb64_tail_store<int32_t>(X, lane, tl);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256, 2) hb48_level_kernel(const LevelParams p)
{
    extern __shared__ __align__(16) unsigned char hb48_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * (blockDim.x >> 5) + wib;
    if (w >= p.n_fam * p.slices) return;
    hb48_level_warp(p, w, reinterpret_cast<int32_t*>(hb48_smem) + (size_t) wib * HB_STAGE_WORDS, lane);
}

} // namespace b200dsp
