// hb48_chain.cuh — K3 (chain form): a run of single-child levels of the DownChannelizer tree as one warp-private cascade.
//
// Same arithmetic and reference functions as hb48_tree.cuh (inthalfbandfiltereo.h:37-63,158-206,357-405,751-830;
// downchannelizer.cpp:50-91).  When the channels of a bank all lie in one half of the band -- a rank's frequency block of a
// bank sharded over N GPUs -- the top log2(N) levels of its tree are a chain: one node per level, every baseband sample
// passes through it, nothing is shared.  The fused pyramid kernel (hb48_fused.cuh) leaves most warps idle there (a level
// of a chain has half the items of the level above).  Here, as in K1 (hb64_cascade.cuh): the pass is cut into time slices,
// ONE WARP owns a slice and runs all L chain stages on it with the stage hand-off buffers in its own shared memory (no block
// barrier); stage s runs once every 2^(s-1) batches of 384 baseband samples, so every instruction has 32 busy lanes.  A
// slice other than the first starts one superphase early with zero history and drops those outputs (FIR: exact); the first
// slice starts from the carried tails, the last one writes them (the same tails the other two tree kernels use).
// The chain's last level goes to its level buffer as packed int16 IQ, where the fused kernel picks it up.
#pragma once
#include "hb48_tree.cuh"

namespace b200dsp {

constexpr int CH_MAXL = 5;

struct ChainParams {
    const uint32_t* in;                 // the pass's baseband (packed int16 IQ), in[i] = root sample C_before + i, 16-byte aligned
    uint32_t*       out;                // the depth-L node's level buffer row
    const uint32_t* tail_in[CH_MAXL];   // the chain node of depth s = 0 .. L-1 (s = 0: the root): its carried tail [TAIL_WORDS]
    uint32_t*       tail_out[CH_MAXL];
    long long       n0;                 // baseband samples of this pass: a multiple of 2^L, the stream aligned at every chain level
    int             L;
    int             slice_sp, n_slices; // superphases (384 * 2^(L-1) baseband samples) per slice
    int             opq_zero, opq_one, opq_mone;
    signed char     rot[8];             // rot[s], s = 1..L: 0 centre, +1 lower half (+j), -1 upper half (-j)
};

// buffer b holds level-b samples [B0 - 64, B0 + 384), B0 = done - 384; write the 64 that end at n_b as the level's carried tail
__device__ __noinline__ void hb48_chain_save_tail(const int32_t* Xs, uint32_t* tout, long long done, long long n_b, int lane)
{
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int t = lane + 32 * h;
        const int rel = (int) (n_b - 64 + t - (done - HB_IN));          // [-64, 384)
        const int32_t* a = Xs + (rel & 1) * HB_ARR + HB_HIST + (rel >> 1);
        tout[t] = ((uint32_t) a[0] & 0xffffu) | ((uint32_t) a[2 * HB_ARR] << 16);
    }
    if (lane == 0) tout[64] = 0u;
}

__global__ void __launch_bounds__(256, 2) hb48_chain_kernel(const ChainParams p)
{
    extern __shared__ __align__(16) unsigned char hb48_chain_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * (blockDim.x >> 5) + wib;
    if (w >= p.n_slices) return;
    const int L = p.L;
    int32_t* X = reinterpret_cast<int32_t*>(hb48_chain_smem) + (size_t) wib * L * HB_STAGE_WORDS;
    const int comp = lane >> 4, j = lane & 15;
    const IntOpaque opq = { p.opq_zero, p.opq_one, p.opq_mone };

    const long long U = (long long) HB_IN << (L - 1);
    const long long sp_total = (p.n0 + U - 1) / U;
    const long long a0 = (long long) w * p.slice_sp;
    long long a1 = a0 + p.slice_sp;
    if (a1 > sp_total) a1 = sp_total;
    const bool first = (w == 0), last = (w == p.n_slices - 1);
    const long long sp_begin = first ? 0 : a0 - 1;
    const long long n_out = p.n0 >> L;

    unsigned badprev = 0, badcur = 0;       // bit b: buffer b's previous / current batch holds a -32768 (its int16 negation wraps)
    for (int s = 0; s < L; ++s) {
        int32_t* Xs = X + s * HB_STAGE_WORDS;
        if (first) {
            const uint32_t* src = p.tail_in[s];
            const uint32_t s0 = src[2 * lane], s1 = src[2 * lane + 1];
            Xs[0 * HB_ARR + lane] = sext_lo16((int32_t) s0);
            Xs[1 * HB_ARR + lane] = sext_lo16((int32_t) s1);
            Xs[2 * HB_ARR + lane] = sext_hi16((int32_t) s0);
            Xs[3 * HB_ARR + lane] = sext_hi16((int32_t) s1);
            if (__any_sync(0xffffffffu, (has_m32768(s0) | has_m32768(s1)) != 0)) badprev |= 1u << s;
        } else {
#pragma unroll
            for (int a = 0; a < 4; ++a) Xs[a * HB_ARR + lane] = 0;
        }
    }

    CascadeParams lp;                       // the int16 loader of K1 reads only these two
    lp.in = p.in; lp.n0 = p.n0;
    const int nph = (int) ((a1 - sp_begin) << (L - 1));
    long long pos = sp_begin * U;
    int4 pre[3];
    LoaderI16<false>::fetch(lp, pos, lane, pre);
    for (int ph = 0; ph < nph; ++ph) {
        {
            uint32_t bad = 0;
#pragma unroll
            for (int t = 0; t < 3; ++t)
                bad |= has_m32768((uint32_t) pre[t].x) | has_m32768((uint32_t) pre[t].y) | has_m32768((uint32_t) pre[t].z) | has_m32768((uint32_t) pre[t].w);
            if (__any_sync(0xffffffffu, bad != 0)) badcur |= 1u;
        }
        LoaderI16<false>::store(lp, X, lane, pre);
        pos += HB_IN;
        if (ph + 1 < nph) LoaderI16<false>::fetch(lp, pos, lane, pre);
        __syncwarp();
        for (int s = 1; s <= L; ++s) {
            if (((ph + 1) & ((1 << (s - 1)) - 1)) != 0) break;
            int32_t* Xin = X + (s - 1) * HB_STAGE_WORDS;
            const int sigma = (int) p.rot[s];
            const unsigned bit = 1u << (s - 1);
            const bool slow = sigma != 0 && ((badcur | badprev) & bit) != 0;
            badprev = (badprev & ~bit) | (badcur & bit);
            badcur &= ~bit;
            int32_t wv[36], cv[16], y[HB_R];
            hb48_load_windows2(Xin, comp, j, sigma != 0, wv, cv);
            const int4 tl = hb64_tail_load<int32_t>(Xin, lane);
            if (last) {
                const long long done = pos >> (s - 1), n_b = p.n0 >> (s - 1);
                if (n_b > done - HB_IN && n_b <= done) hb48_chain_save_tail(Xin, p.tail_out[s - 1], done, n_b, lane);
            }
            if (sigma == 0)   hb48_item<false, false>(wv, cv, 0, opq, y);
            else if (!slow)   hb48_item<true, false>(wv, cv, comp ? sigma : -sigma, opq, y);
            else {
                // (a temporary: handing y itself to the non-inlined function would pin it to local memory on the hot path too)
                int32_t ys[HB_R];
                hb48_slow_child(Xin, comp, j, sigma, 0, opq, ys);
#pragma unroll
                for (int r = 0; r < HB_R; ++r) y[r] = ys[r];
            }
            __syncwarp();
            hb64_tail_store<int32_t>(Xin, lane, tl);
            if (s < L) {
                uint32_t yb = 0;
#pragma unroll
                for (int r = 0; r < HB_R; ++r) { y[r] = wrap16(y[r]); yb |= (y[r] == -32768) ? 1u : 0u; }
                if (__any_sync(0xffffffffu, yb != 0)) badcur |= 1u << s;
                const int fill = ((((ph + 1) >> (s - 1)) - 1) & 1) * (HB_BATCH / 2);
                hb64_store_next<int32_t>(X + s * HB_STAGE_WORDS, comp, j, fill, y);
            } else {
                const long long sp = sp_begin + (((ph + 1) >> (L - 1)) - 1);          // the superphase just completed
                if (sp >= a0) hb48_store_child(p.out + sp * HB_BATCH, y, comp, j, 0, 0, (int) ((n_out - sp * HB_BATCH < HB_BATCH) ? n_out - sp * HB_BATCH : HB_BATCH));
            }
            __syncwarp();
        }
    }
}

} // namespace b200dsp
