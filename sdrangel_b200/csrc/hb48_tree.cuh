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
    // work-queue mode (queue != null): resident warps pull (family, slice) items from *queue; blocks that land on an SM
    // whose bit is set in rsv[] exit at once, leaving those SMs to a concurrent collective (NCCL's ring CTAs need a whole
    // SM's registers and otherwise only get placed between kernels)
    int* queue;
    uint32_t rsv[5];
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

// work-queue form: one resident wave; blocks on reserved SMs exit, the other warps pull items until none is left
__global__ void __launch_bounds__(256, 2) hb48_level_queue_kernel(const LevelParams p)
{
    extern __shared__ __align__(16) unsigned char hb48_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int32_t* X = reinterpret_cast<int32_t*>(hb48_smem) + (size_t) wib * HB_STAGE_WORDS;
    const int total = p.n_fam * p.slices;
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    if (smid < 160u && ((p.rsv[smid >> 5] >> (smid & 31)) & 1u)) return;
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(p.queue, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= total) break;
        hb48_level_warp(p, w, X, lane);
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------------------------
// Two tree levels per launch ("pair kernel"): a warp owns (root = a node of level d-1 with children, time slice) and
// produces the root's children (level d) into warp-private shared-memory buffers and, from those, the grandchildren
// (level d+1) to HBM.  Level d never round-trips HBM (only children that are some channel's leaf are also written out),
// and the grandchildren stage needs no global loader, no unpack and no -32768 scan of packed words.  Used when the call
// is aligned at both levels (no pending sample, even output counts, whole batch pairs); anything else takes the
// one-level kernel above, so arbitrary feed lengths stay exact.
// Slices other than the first recompute one batch pair of warm-up for the children's history (FIR: exact).
// ---------------------------------------------------------------------------------------------------------
struct PairParams {
    const uint32_t* in_base;  long long in_stride;     // level d-1 buffers (roots)
    uint32_t*       mid_base; long long mid_stride;    // level d buffers (children): written only for leaf children
    uint32_t*       out_base; long long out_stride;    // level d+1 buffers (grandchildren)
    const uint32_t* root_tail_in;  uint32_t* root_tail_out;     // [roots][TAIL_WORDS]   (level d-1 tails)
    const uint32_t* child_tail_in; uint32_t* child_tail_out;    // [children][TAIL_WORDS] (level d tails)
    const int*      fam;                               // [n_fam][16]: root, child[3], child_is_leaf[3], grandchild[3][3]
    int n_fam;
    int n_in;          // root samples consumed this call: a multiple of 768
    int slices, pps;   // slices per family, batch pairs per slice
    int opq_zero, opq_one, opq_mone;
};

__device__ __forceinline__ void hb48_hist_fill(int32_t* X, const uint32_t* src, int lane, uint32_t& bad)
{
    const uint32_t s0 = src[2 * lane], s1 = src[2 * lane + 1];
    bad = has_m32768(s0) | has_m32768(s1);
    X[0 * HB_ARR + lane] = sext_lo16((int32_t) s0);
    X[1 * HB_ARR + lane] = sext_lo16((int32_t) s1);
    X[2 * HB_ARR + lane] = sext_hi16((int32_t) s0);
    X[3 * HB_ARR + lane] = sext_hi16((int32_t) s1);
}

__global__ void __launch_bounds__(256, 2) hb48_pair_kernel(const PairParams p)
{
    extern __shared__ __align__(16) unsigned char hb48_smem2[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * (blockDim.x >> 5) + wib;
    if (w >= p.n_fam * p.slices) return;
    const int f = w / p.slices, slice = w - f * p.slices;
    const int nwarps = blockDim.x >> 5;
    int32_t* X0 = reinterpret_cast<int32_t*>(hb48_smem2) + (size_t) wib * 4 * HB_STAGE_WORDS;     // root samples
    int32_t* XC = X0 + HB_STAGE_WORDS;                                                           // 3 child buffers
    int* fam = reinterpret_cast<int*>(hb48_smem2 + (size_t) nwarps * 4 * HB_STAGE_BYTES) + 16 * wib;   // the family, in shared memory
    if (lane < 16) fam[lane] = p.fam[16 * f + lane];
    __syncwarp();
    const int root = fam[0];
    const int comp = lane >> 4, j = lane & 15;
    const IntOpaque opq = { p.opq_zero, p.opq_one, p.opq_mone };
    const uint32_t* B = p.in_base + (long long) root * p.in_stride;
    const uint32_t* rtail = p.root_tail_in + (long long) root * TAIL_WORDS;
    const int npairs = p.n_in / (2 * HB_IN);
    const int p0 = slice * p.pps;
    int p1 = p0 + p.pps;
    if (p1 > npairs) p1 = npairs;
    const bool last = (slice == p.slices - 1);
    if (last) {      // root tail carry (aligned: no pending sample)
        uint32_t* tout = p.root_tail_out + (long long) root * TAIL_WORDS;
        for (int t = lane; t <= 64; t += 32) {
            const int i = p.n_in - 64 + t;
            tout[t] = (i < 0) ? rtail[i + 64] : ((t < 64) ? B[i] : 0u);
        }
    }
    if (p0 >= p1) return;
    const int pw = (slice == 0) ? p0 : p0 - 1;          // first processed pair (one warm-up pair for later slices)
    const int n_child = p.n_in >> 1, n_grand = p.n_in >> 2;
    const bool any_rot_child = (fam[2] >= 0) || (fam[3] >= 0);

    // histories
    uint32_t hb;
    hb48_hist_fill(X0, (pw == 0) ? rtail : (B + (long long) pw * 2 * HB_IN - 64), lane, hb);
    bool prev_bad = any_rot_child && __any_sync(0xffffffffu, hb != 0);
    unsigned cprev_mask = 0;                            // bit c: child c's previous batch held a -32768
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
        if (fam[1 + c] < 0) continue;
        int32_t* Xc = XC + c * HB_STAGE_WORDS;
        if (slice == 0) {
            uint32_t cb;
            hb48_hist_fill(Xc, p.child_tail_in + (long long) fam[1 + c] * TAIL_WORDS, lane, cb);
            if (__any_sync(0xffffffffu, cb != 0)) cprev_mask |= 1u << c;
        } else {
#pragma unroll
            for (int a = 0; a < 4; ++a) Xc[a * HB_ARR + lane] = 0;
        }
    }

    int4 pre[3];
    auto fetch = [&](int q) {
        const uint32_t* src = B + (long long) q * HB_IN + 4 * lane;
#pragma unroll
        for (int t = 0; t < 3; ++t) pre[t] = ldg_nc_v4(src + 128 * t);
    };
    fetch(2 * pw);
#pragma unroll 1
    for (int pr = pw; pr < p1; ++pr) {
        const bool emit = (pr >= p0);
        unsigned cbad_mask = 0;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int q = 2 * pr + half;
            uint32_t bad = 0;
            if (any_rot_child) {
#pragma unroll
                for (int t = 0; t < 3; ++t)
                    bad |= has_m32768((uint32_t) pre[t].x) | has_m32768((uint32_t) pre[t].y) | has_m32768((uint32_t) pre[t].z) | has_m32768((uint32_t) pre[t].w);
            }
            CascadeParams dummy;
            LoaderI16<false>::store(dummy, X0, lane, pre);
            if (q + 1 < 2 * p1) fetch(q + 1);
            const bool bad_now = any_rot_child && __any_sync(0xffffffffu, bad != 0);
            const bool slow = bad_now || prev_bad;
            prev_bad = bad_now;
            __syncwarp();
            {
                const int4 tl = hb64_tail_load<int32_t>(X0, lane);
#pragma unroll 1
                for (int c = 0; c < 3; ++c) {       // children: centre, lower half (+j), upper half (-j); the rotated two share code
                    const int child = fam[1 + c];
                    if (child < 0) continue;
                    int32_t wv[36], cx[16], y[HB_R];
                    hb48_load_windows2(X0, comp, j, c != 0, wv, cx);
                    if (c == 0) {
                        hb48_item<false, false>(wv, cx, 0, opq, y);
                    } else {
                        const int sigma = (c == 1) ? 1 : -1;
                        if (!slow) hb48_item<true, false>(wv, cx, comp ? sigma : -sigma, opq, y);
                        else       hb48_slow_child(X0, comp, j, sigma, 0, opq, y);
                    }
                    uint32_t yb = 0;
#pragma unroll
                    for (int r = 0; r < HB_R; ++r) { y[r] = wrap16(y[r]); yb |= (y[r] == -32768) ? 1u : 0u; }
                    if (__any_sync(0xffffffffu, yb != 0)) cbad_mask |= 1u << c;
                    hb64_store_next<int32_t>(XC + c * HB_STAGE_WORDS, comp, j, half * (HB_BATCH / 2), y);
                    if (fam[4 + c] && emit)       // the child is some channel's leaf: it also goes to its level buffer
                        hb48_store_child(p.mid_base + (long long) child * p.mid_stride, y, comp, j, q * HB_BATCH, 0, n_child);
                }
                __syncwarp();
                hb64_tail_store<int32_t>(X0, lane, tl);
            }
            __syncwarp();
        }
        // grandchildren: each child buffer now holds a full batch (384 child samples)
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
            if (fam[1 + c] < 0) continue;
            const int* gk = fam + 7 + 3 * c;
            int32_t* Xc = XC + c * HB_STAGE_WORDS;
            const bool has_g = (gk[0] >= 0) || (gk[1] >= 0) || (gk[2] >= 0);
            const bool slowc = ((cbad_mask | cprev_mask) >> c) & 1u;
            if (!has_g) continue;
            const int4 tl = hb64_tail_load<int32_t>(Xc, lane);
#pragma unroll 1
            for (int m = 0; m < 3; ++m) {
                const int g = gk[m];
                if (g < 0) continue;
                int32_t wv[36], cx[16], y[HB_R];
                hb48_load_windows2(Xc, comp, j, m != 0, wv, cx);
                if (m == 0) {
                    hb48_item<false, false>(wv, cx, 0, opq, y);
                } else {
                    const int sigma = (m == 1) ? 1 : -1;
                    if (!slowc) hb48_item<true, false>(wv, cx, comp ? sigma : -sigma, opq, y);
                    else        hb48_slow_child(Xc, comp, j, sigma, 0, opq, y);
                }
                if (emit) hb48_store_child(p.out_base + (long long) g * p.out_stride, y, comp, j, pr * HB_BATCH, 0, n_grand);
            }
            __syncwarp();
            hb64_tail_store<int32_t>(Xc, lane, tl);
            __syncwarp();
        }
        cprev_mask = cbad_mask;
    }
    if (last) {      // children's tails for the next call: the 64 newest child samples sit in the history regions
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
            if (fam[1 + c] < 0) continue;
            const int32_t* Xc = XC + c * HB_STAGE_WORDS;
            uint32_t* tout = p.child_tail_out + (long long) fam[1 + c] * TAIL_WORDS;
            const bool has_g = (fam[7 + 3 * c] >= 0) || (fam[8 + 3 * c] >= 0) || (fam[9 + 3 * c] >= 0);
            // without grandchildren the buffer was never rotated into the history region: the newest samples are at [192,224)
            const int off = has_g ? 0 : HB_BATCH;
            const uint32_t e = ((uint32_t) Xc[0 * HB_ARR + off + lane] & 0xffffu) | ((uint32_t) Xc[2 * HB_ARR + off + lane] << 16);
            const uint32_t o = ((uint32_t) Xc[1 * HB_ARR + off + lane] & 0xffffu) | ((uint32_t) Xc[3 * HB_ARR + off + lane] << 16);
            tout[2 * lane] = e; tout[2 * lane + 1] = o;
            if (lane == 0) tout[64] = 0u;
        }
    }
}

} // namespace b200dsp

