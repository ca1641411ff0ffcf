// hb48_fused.cuh — K3 (fused form): several levels of the DownChannelizer half-band tree per launch, intermediates in
// shared memory.
//
// Same arithmetic and the same reference functions as hb48_tree.cuh (paths relative to the reference tree):
//   DownChannelizer::feed                                              sdrbase/dsp/downchannelizer.cpp:50-91
//   IntHalfbandFilterEO<qint32,qint32,48>::workDecimate{Center,LowerHalf,UpperHalf}   inthalfbandfiltereo.h:37-63,158-206,357-405
//   storeSample / doFIR(Sample*)                                       inthalfbandfiltereo.h:751-767,792-830
//
// B200 design.  The tree is cut into GROUPS: a node of depth b (the group root) and its descendants down to depth b+k,
// k <= 4.  One launch evaluates all groups of one depth range.  A CTA (8 warps) walks a contiguous range of (group, time
// tile) pairs; for a tile of T root samples it
//   1. unpacks the root's int16 IQ words (128-bit coalesced loads) into level 0 of its shared-memory pyramid
//      ([node][component][parity][32 history + T/2^(j+1)] int32, the K1 layout, so a lane reads its 36-word register window
//      with 128-bit conflict-free shared loads),
//   2. for j = 0..k-1: the warps take the (family, 384-sample batch) items of level j; an item produces 192 outputs of
//      every child of the family -- lower-half and upper-half children from ONE shared tap sum -- and writes them either
//      to level j+1 of the pyramid (int32, already wrapped to int16) or, for nodes some other launch or a channel needs, as
//      packed int16 IQ to HBM.  One block barrier per level.
//   3. slides each level's newest 64 samples to the front of its arrays (the next tile's filter history).
// So a level inside a group never touches HBM: the 11-level tree of the 1024-channel plan moves 3 level arrays instead of
// 11.  State: the first tile of a group starts from the carried per-node tails (the same ones hb48_level_kernel uses, so
// both kernels can serve one stream alternately); a CTA whose range starts in the middle of a group's stream first runs
// the previous tile with zero history and its stores off (FIR: finite memory, exact).  The last tile writes the tails.
// Coefficients are scaled by 32, so the stage output, wrapped to int16, is simply the high half of the accumulator.
// Used when a pass starts aligned at every level (no pending sample, even pair index) and its length is a multiple of
// 2^depth; any other pass takes hb48_level_kernel.
#pragma once
#include "hb48_tree.cuh"

namespace b200dsp {

constexpr int FZ_MAXK = 4;
#ifndef FZ_WARPS_DEF
#define FZ_WARPS_DEF 8
#endif
constexpr int FZ_WARPS = FZ_WARPS_DEF;           // warps per CTA: one (family, batch) item per warp per level of a minimal tile
constexpr int FZ_THREADS = FZ_WARPS * 32;
constexpr int FZ_CTAS_PER_SM = 16 / FZ_WARPS;    // 128 registers per thread: 16 warps per SM

struct FusedFam {            // a parent at group level j and its children at level j+1; offsets precomputed for the launch's tile size
    int pbase, pslot, pindex;    // parent: word offset of its slot within pyramid level j, slot number, index within its tree level (tails)
    int cbase[3];                // children C, L, U: word offset of the slot within pyramid level j+1 (-1: not a parent inside this group)
    int cbit[3];                 // 1 << slot of that child (0 when cbase < 0)
    int cindex[3];               // children written to HBM: >= 0 index within their tree level (raw stage output for another launch
                                 // or the finalize kernel); <= -2: channel -2 - cindex ends here and takes trunc(y / 2^shift) straight
                                 // into its output buffer (downchannelizer.cpp:78-83); -1: not written
};
struct LeafChan { const uint32_t* src; uint32_t* dst; int depth; int shift; int direct; int pad; };     // static per channel (until a reallocation)
struct FusedGroup {
    int root_index, k;
    int fam_begin[FZ_MAXK + 1];     // families of level j: [fam_begin[j], fam_begin[j+1]) of the launch's family table
    int nslots[FZ_MAXK];
    int pad[5];
};
static_assert(sizeof(FusedGroup) == 64, "FusedGroup is loaded as 16 words");
static_assert(sizeof(FusedFam) == 48, "FusedFam is loaded as three int4");

struct FusedParams {
    const uint32_t* in_base;  long long in_stride;                  // depth-b node streams (packed int16 IQ)
    uint32_t*       out_base[FZ_MAXK]; long long out_stride[FZ_MAXK];   // level buffers of depths b+1 .. b+k
    const uint32_t* tail_in[FZ_MAXK];  uint32_t* tail_out[FZ_MAXK];     // carried tails of depths b .. b+k-1
    const FusedGroup* groups; const FusedFam* fams;
    const LeafChan* leaf; long long leaf_count[FZ_MAXK];            // channel table; outputs of depth b+j+1 channels already produced in this feed
    int n_groups, T, tpr;        // tile size in root samples (384 * 2^n), tiles per root stream this pass
    int l2items;                 // log2(T / 384)
    int n_root;                  // root samples per group this pass (a multiple of 2^k)
    int lvl_off[FZ_MAXK + 1];    // word offset of each pyramid level in shared memory
    int opq_zero, opq_one, opq_mone;
};

// y'[r] = (c[1+r] << 16) + 32 * sum_i h_i (w[24+r-i] + w[1+r+i])       (centre child: no rotation)
__device__ __forceinline__ void fz_centre(const int32_t (&w)[36], const int32_t (&c)[16], const IntOpaque& q, uint32_t (&y)[HB_R])
{
#pragma unroll
    for (int r = 0; r < HB_R; ++r) {
        uint32_t acc = (uint32_t) c[1 + r] << 16;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const uint32_t a = (uint32_t) w[24 + r - i], b = (uint32_t) w[1 + r + i];
            const uint32_t s = (i < HB_XH) ? mad_fma(a, q.one, b) : add_alu(a, b, q.zero);
            acc += (uint32_t) (32 * hb48_h(i)) * s;
        }
        y[r] = acc;
    }
}

// F'[r] = 32 * (the rotated tap sum both rotated children share, see hb48_pair_terms), pair index parity 0
__device__ __forceinline__ void fz_rot_sum(const int32_t (&w)[36], const IntOpaque& q, uint32_t (&F)[HB_R])
{
#pragma unroll
    for (int r = 0; r < HB_R; ++r) {
        const int sk = (r & 1) ? 1 : -1;
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int g = ((i & 1) ? -sk : sk) * 32 * hb48_h(i);
            const uint32_t a = (uint32_t) w[24 + r - i], b = (uint32_t) w[1 + r + i];
            const uint32_t d = (i < HB_XH) ? mad_fma(b, q.mone, a) : sub_alu(a, b, q.zero);
            acc += (uint32_t) g * d;
        }
        F[r] = acc;
    }
}

__device__ __forceinline__ void fz_load_w(const int32_t* xo, int32_t (&w)[36])
{
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int4 v = *reinterpret_cast<const int4*>(xo + 4 * q);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
}
__device__ __forceinline__ void fz_load_c(const int32_t* xe, int32_t (&c)[16])
{
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int4 v = *reinterpret_cast<const int4*>(xe + 4 * q);
        c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
    }
}

// rare path: a -32768 somewhere in the parent's tile or the one before (its int16 negation wraps): explicit rotation.
// P = the parent's slot, off = array position of the lane's window origin.
__device__ __noinline__ void fz_slow_child(const int32_t* P, int A, int off, int comp, int sigma, const IntOpaque& opq, uint32_t (&y)[HB_R])
{
    int32_t wv[36], co[16], wr[36], cr[16];
    fz_load_w(P + (comp * 2 + 1) * A + off + 8, wv);
    fz_load_c(P + ((comp ^ 1) * 2) * A + off + 20, co);
    hb48_rotate_exact(wv, co, comp, sigma, 0, wr, cr);
    fz_centre(wr, cr, opq, y);
}

// 12 outputs of one component -> the child's pyramid arrays (even k -> even array at xe, odd k -> odd array); returns the minimum
__device__ __forceinline__ int32_t fz_store_smem(int32_t* xe, int A, const uint32_t (&y)[HB_R])
{
    int32_t v[HB_R];
#pragma unroll
    for (int r = 0; r < HB_R; ++r) v[r] = sext_hi16((int32_t) y[r]);
    int32_t* xo = xe + A;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        *reinterpret_cast<int2*>(xe + 2 * q) = make_int2(v[4 * q], v[4 * q + 2]);
        *reinterpret_cast<int2*>(xo + 2 * q) = make_int2(v[4 * q + 1], v[4 * q + 3]);
    }
    int32_t m = min(min(v[0], v[1]), v[2]);
#pragma unroll
    for (int r = 3; r + 1 < HB_R; r += 2) m = min(min(m, v[r]), v[r + 1]);
    return min(m, v[HB_R - 1]);
}

__device__ __forceinline__ void fz_store12(uint32_t* o, const uint32_t (&wds)[HB_R], int k0, int n_valid)
{
    // (a channel's output buffer is written at its running count within the feed: any word alignment)
    if (k0 + HB_R <= n_valid && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
        for (int t = 0; t < 3; ++t) *reinterpret_cast<uint4*>(o + 4 * t) = make_uint4(wds[4 * t], wds[4 * t + 1], wds[4 * t + 2], wds[4 * t + 3]);
    } else {
#pragma unroll
        for (int t = 0; t < HB_R; ++t) if (k0 + t < n_valid) o[t] = wds[t];
    }
}

struct FzDst { uint32_t* ptr; int shift; };      // shift < 0: the raw int16 stage output; else the channel output trunc(y / 2^shift)

__device__ __forceinline__ FzDst fz_dst(const FusedParams& p, int j, int ci, long long tile_out)
{
    FzDst d;
    if (ci >= 0) { d.ptr = p.out_base[j] + (long long) ci * p.out_stride[j] + tile_out; d.shift = -1; }
    else {
        const LeafChan* lc = p.leaf + (-2 - ci);
        d.ptr = lc->dst + p.leaf_count[j] + tile_out; d.shift = lc->shift;
    }
    return d;
}

// The channel output is the C++ truncating division of the int16 stage output by 2^shift (downchannelizer.cpp:78-83).  On the
// accumulator (value in the high half): add 2^shift - 1 at bit 16 when negative, then shift by 16 + shift.  Each lane does
// this for its own component BEFORE the exchange; the words are then packed from the low halves.
__device__ __forceinline__ void fz_finalize12(uint32_t (&y)[HB_R], int shift)
{
    const int nb = -(((1 << shift) - 1) << 16);
#pragma unroll
    for (int t = 0; t < HB_R; ++t) {
        const int v = (int) y[t];
        y[t] = (uint32_t) (((v >> 31) * nb + v) >> (16 + shift));
    }
}
__device__ __forceinline__ uint32_t fz_pack(uint32_t re, uint32_t im, bool low)
{
    uint32_t w;
    if (low) asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(w) : "r"(re), "r"(im));       // finalised values: low halves
    else     asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(w) : "r"(re), "r"(im));       // raw accumulators: the int16-wrapped stage outputs are the high halves
    return w;
}

// packed int16 IQ to HBM: lane (comp 0, lj) packs its re with the im of lane (comp 1, lj) and stores 12 words
__device__ __forceinline__ void fz_store_global(const FzDst& d, uint32_t (&y)[HB_R], int comp, int k0, int n_valid)
{
    uint32_t wds[HB_R];
    const bool fin = d.shift >= 0;
    if (fin) fz_finalize12(y, d.shift);
#pragma unroll
    for (int t = 0; t < HB_R; ++t) {
        const uint32_t im = __shfl_down_sync(0xffffffffu, y[t], 16);
        wds[t] = fz_pack(y[t], im, fin);
    }
    if (comp != 0 || k0 >= n_valid) return;
    fz_store12(d.ptr + k0, wds, k0, n_valid);
}

// both rotated children at once.  yA is the lane's own child (lower half on the comp-0 lanes, upper half on the comp-1
// lanes), yB the other one: each lane sends yB to its partner lane and stores its own child's 12 words.  (Raw or
// finalised is the same for both children -- the caller checks -- but the two channels' shifts may differ.)
__device__ __forceinline__ void fz_store_global_pair(const FzDst& dA, uint32_t (&yA)[HB_R], uint32_t (&yB)[HB_R], int comp, int k0, int n_valid)
{
    uint32_t wds[HB_R];
    const bool fin = dA.shift >= 0;
    if (fin) {
        fz_finalize12(yA, dA.shift);
        fz_finalize12(yB, __shfl_xor_sync(0xffffffffu, dA.shift, 16));       // yB belongs to the partner lane's channel
    }
#pragma unroll
    for (int t = 0; t < HB_R; ++t) {
        const uint32_t got = __shfl_xor_sync(0xffffffffu, yB[t], 16);
        // comp 0: (re = own, im = partner's); comp 1: (re = partner's, im = own)
        wds[t] = fz_pack(comp ? got : yA[t], comp ? yA[t] : got, fin);
    }
    if (k0 >= n_valid) return;
    fz_store12(dA.ptr + k0, wds, k0, n_valid);
}

constexpr int FZ_MAX_FAMS = 40;      // per group: at most 1 + 3 + 9 + 27 parents
struct FzLevel { int off, A, fb, nfam; };   // pyramid level j of the current group: word offset, words per array, first family, families
struct FzShared {
    FusedGroup grp;
    __align__(16) FzLevel lvl[FZ_MAXK + 1];
    unsigned bad[2][FZ_MAXK + 1];   // [0] this tile, [1] the tile before: bit s = slot s of that level holds a -32768
    __align__(16) FusedFam fams[FZ_MAX_FAMS];     // the group's families (fam_begin rebased to 0)
};

struct FzPhase {                     // per-lane constants of one level of one tile
    const int32_t* Lp; int32_t* Lc;  // pyramid level j (parents) and j+1 (children)
    int A, A1;                       // words per array at level j / j+1
    int woff, coff_own, coff_oth;    // lane offsets of the register windows within a parent slot
    int soff;                        // lane offset of the stores within a child slot (own component's even array)
    unsigned badmask;                // slots of level j with a -32768 in this tile or the one before
    long long tile_out; int nv_child;
};

__device__ __forceinline__ FusedFam fz_get_fam(const FzShared& sh, int f)
{
    const int4* s = reinterpret_cast<const int4*>(&sh.fams[f]);
    const int4 a = s[0], b = s[1], c = s[2];
    FusedFam r;
    r.pbase = a.x; r.pslot = a.y; r.pindex = a.z; r.cbase[0] = a.w; r.cbase[1] = b.x; r.cbase[2] = b.y;
    r.cbit[0] = b.z; r.cbit[1] = b.w; r.cbit[2] = c.x; r.cindex[0] = c.y; r.cindex[1] = c.z; r.cindex[2] = c.w;
    return r;
}

// one item: family `fam` of group level j, batch q (parent samples [384 q, 384 q + 384) of this tile)
__device__ __forceinline__ void fz_item(const FusedParams& p, FzShared& sh, const FzPhase& ph, const FusedFam& fam, int j, int q, int lane, bool emit)
{
    const int comp = lane >> 4, lj = lane & 15;
    const IntOpaque opq = { p.opq_zero, p.opq_one, p.opq_mone };
    const int A1 = ph.A1;
    const int32_t* P = ph.Lp + fam.pbase + HB_BATCH * q;
    int32_t* Sc = ph.Lc + ph.soff + (HB_BATCH / 2) * q;
    const int k0 = HB_BATCH * q + HB_R * lj;                   // first output of this lane within the tile
    int32_t wv[36];
    fz_load_w(P + ph.woff, wv);
    unsigned newbad = 0;
    if (fam.cbase[0] >= 0 || fam.cindex[0] != -1) {
        int32_t cv[16];
        uint32_t y[HB_R];
        fz_load_c(P + ph.coff_own, cv);
        fz_centre(wv, cv, opq, y);
        if (fam.cbase[0] >= 0) {
            const int32_t m = fz_store_smem(Sc + fam.cbase[0], A1, y);
            if (__any_sync(0xffffffffu, m == -32768)) newbad |= (unsigned) fam.cbit[0];
        }
        if (fam.cindex[0] != -1 && emit) fz_store_global(fz_dst(p, j, fam.cindex[0], ph.tile_out), y, comp, k0, ph.nv_child);
    }
    const bool hasL = fam.cbase[1] >= 0 || fam.cindex[1] != -1, hasU = fam.cbase[2] >= 0 || fam.cindex[2] != -1;
    if (hasL || hasU) {
        const bool slow = ((ph.badmask >> fam.pslot) & 1u) != 0;
        // the pair path needs both children treated alike (both or neither kept in the pyramid / written to HBM): its
        // stores pick the child by half-warp, so anything else would diverge inside a warp
        const bool alike = ((fam.cbase[1] >= 0) == (fam.cbase[2] >= 0)) && ((fam.cindex[1] >= 0) == (fam.cindex[2] >= 0)) &&
                           ((fam.cindex[1] == -1) == (fam.cindex[2] == -1));
        if (hasL && hasU && alike && !slow) {
            // Both rotated children from one tap sum F.  With yL = F + co*m, yU = F - co*m and m = +-65536 by (component,
            // output parity), the lane's OWN child (lower half on comp-0 lanes, upper half on comp-1 lanes) is
            // yA = F + co*mA with mA = (r odd ? -65536 : 65536) on every lane, the other child yB = F - co*mA: immediates.
            int32_t co[16];
            uint32_t yA[HB_R], yB[HB_R];
            fz_load_c(P + ph.coff_oth, co);
            fz_rot_sum(wv, opq, yB);
#pragma unroll
            for (int r = 0; r < HB_R; ++r) {
                const int mA = (r & 1) ? -65536 : 65536;
                const uint32_t F = yB[r];
                yA[r] = (uint32_t) co[1 + r] * (uint32_t) mA + F;
                yB[r] = (uint32_t) co[1 + r] * (uint32_t) (-mA) + F;
            }
            if (fam.cbase[1] >= 0) {
                const int baseA = comp ? fam.cbase[2] : fam.cbase[1], baseB = comp ? fam.cbase[1] : fam.cbase[2];
                const int32_t ma = fz_store_smem(Sc + baseA, A1, yA);
                const int32_t mb = fz_store_smem(Sc + baseB, A1, yB);
                const unsigned bal = __ballot_sync(0xffffffffu, ma == -32768), bbl = __ballot_sync(0xffffffffu, mb == -32768);
                if ((bal & 0xffffu) | (bbl >> 16)) newbad |= (unsigned) fam.cbit[1];
                if ((bal >> 16) | (bbl & 0xffffu)) newbad |= (unsigned) fam.cbit[2];
            }
            if (emit && fam.cindex[1] != -1)
                fz_store_global_pair(fz_dst(p, j, comp ? fam.cindex[2] : fam.cindex[1], ph.tile_out), yA, yB, comp, k0, ph.nv_child);
        } else {
#pragma unroll 1
            for (int m = 1; m <= 2; ++m) {
                // (selected, not indexed: a runtime index would move the family descriptor to local memory)
                const int cb = (m == 1) ? fam.cbase[1] : fam.cbase[2], ci = (m == 1) ? fam.cindex[1] : fam.cindex[2];
                const int bit = (m == 1) ? fam.cbit[1] : fam.cbit[2];
                if (cb < 0 && ci == -1) continue;
                const int sigma = (m == 1) ? 1 : -1;
                uint32_t y[HB_R];
                if (!slow) {
                    int32_t co[16];
                    fz_load_c(P + ph.coff_oth, co);
                    fz_rot_sum(wv, opq, y);
                    const int cs = ((comp != 0) == (m == 1)) ? 65536 : -65536;
#pragma unroll
                    for (int r = 0; r < HB_R; ++r) y[r] += (uint32_t) co[1 + r] * (uint32_t) ((r & 1) ? cs : -cs);
                } else {
                    // (a temporary: handing y itself to the non-inlined function would pin it to local memory)
                    uint32_t ys[HB_R];
                    fz_slow_child(ph.Lp + fam.pbase, ph.A, HB_BATCH * q + HB_R * lj, comp, sigma, opq, ys);
#pragma unroll
                    for (int r = 0; r < HB_R; ++r) y[r] = ys[r];
                }
                if (cb >= 0) {
                    const int32_t mn = fz_store_smem(Sc + cb, A1, y);
                    if (__any_sync(0xffffffffu, mn == -32768)) newbad |= (unsigned) bit;
                }
                if (ci != -1 && emit) fz_store_global(fz_dst(p, j, ci, ph.tile_out), y, comp, k0, ph.nv_child);
            }
        }
    }
    if (newbad && lane == 0) atomicOr(&sh.bad[0][j + 1], newbad);
}

__device__ __forceinline__ void fz_slide_tails(int32_t* Lp, int A, int nslots, int tid)
{
    for (int idx = tid; idx < nslots * 32; idx += FZ_THREADS) {
        int32_t* arr = Lp + (idx >> 3) * A;           // (slot * 4 + array) * A
        *reinterpret_cast<int4*>(arr + 4 * (idx & 7)) = *reinterpret_cast<const int4*>(arr + (A - HB_HIST) + 4 * (idx & 7));
    }
}

// one tile of one group.  emit = false: warm-up (nothing leaves the CTA).
__device__ __forceinline__ void fz_tile(const FusedParams& p, int32_t* S, FzShared& sh, int t, bool emit)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int comp = lane >> 4, lj = lane & 15;
    const FusedGroup& g = sh.grp;
    const int k = g.k;
    const long long t0 = (long long) t * p.T;
    const int nv = (p.n_root - t0 < p.T) ? (int) (p.n_root - t0) : p.T;
    const bool last = (t == p.tpr - 1), full = (nv == p.T);
    // ---- root tile: packed int16 IQ -> level 0 arrays (all loads of a thread in flight together)
    {
        const uint32_t* B = p.in_base + (long long) g.root_index * p.in_stride + t0;
        const int A0 = HB_HIST + (p.T >> 1);
        int32_t* L0 = S + p.lvl_off[0] + HB_HIST;
        uint32_t bad = 0;
        for (int v0 = tid; v0 < (p.T >> 2); v0 += 3 * FZ_THREADS) {
            int4 x[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int v = v0 + u * FZ_THREADS;
                if (full || 4 * v + 4 <= nv) x[u] = ldg_nc_v4(B + 4 * v);
                else {
                    x[u].x = (4 * v < nv) ? (int) B[4 * v] : 0;         x[u].y = (4 * v + 1 < nv) ? (int) B[4 * v + 1] : 0;
                    x[u].z = (4 * v + 2 < nv) ? (int) B[4 * v + 2] : 0; x[u].w = (4 * v + 3 < nv) ? (int) B[4 * v + 3] : 0;
                }
            }
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int v = v0 + u * FZ_THREADS;
                bad |= has_m32768((uint32_t) x[u].x) | has_m32768((uint32_t) x[u].y) | has_m32768((uint32_t) x[u].z) | has_m32768((uint32_t) x[u].w);
                int32_t* a = L0 + 2 * v;
                *reinterpret_cast<int2*>(a)          = make_int2(sext_lo16(x[u].x), sext_lo16(x[u].z));
                *reinterpret_cast<int2*>(a + A0)     = make_int2(sext_lo16(x[u].y), sext_lo16(x[u].w));
                *reinterpret_cast<int2*>(a + 2 * A0) = make_int2(sext_hi16(x[u].x), sext_hi16(x[u].z));
                *reinterpret_cast<int2*>(a + 3 * A0) = make_int2(sext_hi16(x[u].y), sext_hi16(x[u].w));
            }
        }
        if (__any_sync(0xffffffffu, bad != 0) && lane == 0) atomicOr(&sh.bad[0][0], 1u);
        // the next tile of this stream into L2 while this one is computed (one 128-byte line per thread)
        if (!last && tid < (p.T >> 5)) asm volatile("prefetch.global.L2 [%0];" :: "l"(B + p.T + 32 * tid));
    }
    __syncthreads();
    for (int j = 0; j < k; ++j) {
        FzPhase ph;
        const int4 lv = *reinterpret_cast<const int4*>(&sh.lvl[j]), lv1 = *reinterpret_cast<const int4*>(&sh.lvl[j + 1]);
        ph.badmask = sh.bad[0][j] | sh.bad[1][j];
        ph.A = lv.y; ph.A1 = lv1.y;
        if (j > 0) {
            // level j-1 is consumed: its newest 64 samples become the next tile's history; its flags move on
            if (!last) fz_slide_tails(S + sh.lvl[j - 1].off, sh.lvl[j - 1].A, g.nslots[j - 1], tid);
            if (tid == 0) { sh.bad[1][j - 1] = sh.bad[0][j - 1]; sh.bad[0][j - 1] = 0; }
        }
        const int nvj = nv >> j;
        const int fb = lv.z, nfam = lv.w;
        ph.Lp = S + lv.x;
        ph.Lc = S + lv1.x;
        ph.woff = (comp * 2 + 1) * ph.A + HB_R * lj + 8;
        ph.coff_own = (comp * 2) * ph.A + HB_R * lj + 20;
        ph.coff_oth = ((comp ^ 1) * 2) * ph.A + HB_R * lj + 20;
        ph.soff = (comp * 2) * ph.A1 + HB_HIST + 6 * lj;
        ph.tile_out = t0 >> (j + 1); ph.nv_child = nvj >> 1;
        // whole tile: 2^(l2items - j) batches per family (shift and mask); a stream's ragged last tile divides
        const int sft = p.l2items - j, msk = (1 << sft) - 1;
        const int ipf = full ? (1 << sft) : (nvj + HB_IN - 1) / HB_IN;
        for (int it = warp; it < nfam * ipf; it += FZ_WARPS) {
            int f, q;
            if (full) { f = it >> sft; q = it & msk; }
            else      { f = it / ipf; q = it - f * ipf; }
            const FusedFam fam = fz_get_fam(sh, fb + f);
            fz_item(p, sh, ph, fam, j, q, lane, emit);
        }
        __syncthreads();
    }
    if (!last) {
        fz_slide_tails(S + p.lvl_off[k - 1], HB_HIST + (p.T >> k), g.nslots[k - 1], tid);
    } else if (emit) {
        // the group's stream ends here: every parent's newest 64 samples (+ no pending sample) are the carried tails
        for (int j = 0; j < k; ++j) {
            const int A = HB_HIST + (p.T >> (j + 1));
            const int nvj = nv >> j;
            for (int f = g.fam_begin[j] - g.fam_begin[0] + warp; f < g.fam_begin[j + 1] - g.fam_begin[0]; f += FZ_WARPS) {
                const FusedFam fam = fz_get_fam(sh, f);
                const int32_t* P = S + p.lvl_off[j] + fam.pbase;
                uint32_t* tout = p.tail_out[j] + (long long) fam.pindex * TAIL_WORDS;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = nvj - 64 + lane + 32 * h;            // tile-relative sample index, may be negative (history)
                    const int pos = HB_HIST + (i >> 1), par = i & 1;
                    const uint32_t re = (uint32_t) P[par * A + pos], im = (uint32_t) P[(2 + par) * A + pos];
                    tout[lane + 32 * h] = (re & 0xffffu) | (im << 16);
                }
                if (lane == 0) tout[64] = 0u;
            }
        }
    }
    if (tid == 0) { sh.bad[1][k - 1] = sh.bad[0][k - 1]; sh.bad[0][k - 1] = 0; }
    __syncthreads();
}

__global__ void __launch_bounds__(FZ_THREADS, FZ_CTAS_PER_SM) hb48_fused_kernel(const FusedParams p)
{
    extern __shared__ __align__(16) unsigned char fz_smem[];
    __shared__ FzShared sh;
    int32_t* S = reinterpret_cast<int32_t*>(fz_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long tot = (long long) p.n_groups * p.tpr;
    const long long c0 = tot * blockIdx.x / gridDim.x, c1 = tot * (blockIdx.x + 1) / gridDim.x;
    int cur = -1;
    for (long long ft = c0; ft < c1; ++ft) {
        const int gi = (int) (ft / p.tpr), t = (int) (ft - (long long) gi * p.tpr);
        int warm = 0;
        if (gi != cur) {
            cur = gi;
            if (tid < 16) reinterpret_cast<int*>(&sh.grp)[tid] = reinterpret_cast<const int*>(p.groups + gi)[tid];
            if (tid < 2 * (FZ_MAXK + 1)) (&sh.bad[0][0])[tid] = 0u;
            __syncthreads();
            const FusedGroup& g = sh.grp;
            {
                const int nf = g.fam_begin[g.k] - g.fam_begin[0];
                const int* src = reinterpret_cast<const int*>(p.fams + g.fam_begin[0]);
                for (int i = tid; i < nf * 12 && i < FZ_MAX_FAMS * 12; i += FZ_THREADS) reinterpret_cast<int*>(sh.fams)[i] = src[i];
                if (tid <= FZ_MAXK) {
                    FzLevel L;
                    L.off = p.lvl_off[tid]; L.A = HB_HIST + (p.T >> (tid + 1));
                    L.fb = g.fam_begin[tid < g.k ? tid : g.k] - g.fam_begin[0];
                    L.nfam = (tid < g.k) ? g.fam_begin[tid + 1] - g.fam_begin[tid] : 0;
                    sh.lvl[tid] = L;
                }
            }
            __syncthreads();
            if (t == 0) {
                // stream start of this pass: the carried tails are the history
                for (int j = 0; j < g.k; ++j) {
                    const int A = HB_HIST + (p.T >> (j + 1));
                    for (int f = g.fam_begin[j] - g.fam_begin[0] + warp; f < g.fam_begin[j + 1] - g.fam_begin[0]; f += FZ_WARPS) {
                        const FusedFam fam = fz_get_fam(sh, f);
                        int32_t* X = S + p.lvl_off[j] + fam.pbase;
                        const uint32_t* src = p.tail_in[j] + (long long) fam.pindex * TAIL_WORDS;
                        const uint32_t s0 = src[2 * lane], s1 = src[2 * lane + 1];
                        const uint32_t hb = has_m32768(s0) | has_m32768(s1);
                        X[0 * A + lane] = sext_lo16((int32_t) s0);
                        X[1 * A + lane] = sext_lo16((int32_t) s1);
                        X[2 * A + lane] = sext_hi16((int32_t) s0);
                        X[3 * A + lane] = sext_hi16((int32_t) s1);
                        if (__any_sync(0xffffffffu, hb != 0) && lane == 0) atomicOr(&sh.bad[1][j], 1u << fam.pslot);
                    }
                }
                __syncthreads();
            } else {
                for (int j = 0; j < g.k; ++j) {
                    const int A = HB_HIST + (p.T >> (j + 1));
                    int32_t* Lp = S + p.lvl_off[j];
                    for (int idx = tid; idx < g.nslots[j] * 128; idx += FZ_THREADS) Lp[(idx >> 5) * A + (idx & 31)] = 0;
                }
                __syncthreads();
                warm = 1;
            }
        }
        // (one call site, so the tile code exists once: the warm-up pass is the previous tile with its stores off)
        for (int pass = 1 - warm; pass < 2; ++pass) fz_tile(p, S, sh, t - 1 + pass, pass != 0);
    }
}

} // namespace b200dsp
