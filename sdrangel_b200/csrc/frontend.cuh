// frontend.cuh — K4: the channel plugin front-end, NCO mix + polyphase Interpolator::decimate, for every channel of a bank.
//
// Replaces (paths relative to the reference tree), per channel:
//   Complex c(re, im); c *= m_nco.nextIQ();                          plugins/channelrx/demodnfm/nfmdemod.cpp:152-153
//   m_interpolator.decimate(&m_interpolatorDistanceRemain, c, &ci)   sdrbase/dsp/interpolator.h:23-36,107-113,183-194
//   m_interpolatorDistanceRemain += m_interpolatorDistance           plugins/channelrx/demodnfm/nfmdemod.cpp:315
//   NCO::nextPhase/nextIQ (phase advanced BEFORE the lookup)         sdrbase/dsp/nco.h:43-50, nco.cpp:60-64
//
// B200 design: the reference's float32 "distance" recurrence decides which inputs emit an output and at which of the 16
// phases; it depends only on the ratio, not on the data, so a schedule kernel replays it exactly (one thread per channel,
// one iteration per OUTPUT) on a side stream while the tree kernels run.  Then one CTA per channel mixes the channel's new
// samples with the table NCO (the phase of sample i is phase0 + (i+1)*inc, so the mix is data-parallel) and every thread
// computes outputs as 72-tap dot products  y = sum_k taps[phase][k] * z[idx-k]  with the taps in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200dsp {

constexpr int FE_MAX_TAPS = 128;          // taps per phase supported by the kernel's history area (reference default: 72)
constexpr int FE_TILE = 1024;             // channel samples per CTA tile
constexpr int FE_THREADS = 128;
constexpr int FE_NOUT = 4;                // consecutive outputs per thread (they share the shared-memory reads of z)
constexpr int FE_PAD = 16;                // zero taps on both sides of every phase row: out-of-range tap indices contribute 0
constexpr int FE_HIST_WORDS = FE_MAX_TAPS + 4;   // per ping-pong half: FE_MAX_TAPS samples + the NCO phase

struct FrontendChan {            // one per channel with a front-end, device array (static between reallocations)
    const uint32_t* in;          // the channel's channelizer output buffer of this feed (packed int16 IQ)
    uint32_t*       hist;        // [2][FE_HIST_WORDS] carried state (ping-pong): newest channel samples (packed int16 IQ), NCO phase
    const float*    taps;        // [phase_steps][ntaps]
    float2*         out;         // outputs of this feed
    int*            sched;       // [cap] packed (idx << 8 | phase) of this pass's outputs
    int*            tile_start;  // [cap / FE_TILE + 2] first output of each input tile (written by the schedule kernel)
    int*            state;       // [4]: unused, float distance remain (bits), outputs of the last pass, outputs of this feed
    long long*      plan;        // [4] written by the schedule kernel: k0 (outputs listed in sched), i0, D, closed-form output count
    long long       A;           // ratio * 2^23 (exact integer)
    int             in_f32;      // 0: `in` is packed int16 IQ and the NCO mix applies; 1: `in` is complex64, no NCO (plain Interpolator)
    int             hist_stride; // words per ping-pong half of `hist`
    int             lattice;     // 1: every sum of the distance recurrence is exactly representable => closed-form schedule
    int             phshift;     // 23 - log2(phase_steps)
    int             depth;       // S: selects the per-depth pass counts
    int             inc;         // NCO phase increment
    int             ntaps, phase_steps;
    float           ratio;       // m_interpolatorDistance
    int             mode;        // caller loop the schedule replays: 0 Interpolator::decimate (Rx plugins), 1 ::interpolate (Tx pull loop),
                                 // 2 ::resample (do-while per input); modes 1-2 store idx + 1 (an output may precede the pass's first input)
    int             sched_cap;   // entries of `sched` (modes 1-2: outputs are not bounded by the input count)
    int             scan_kb;     // > 0: mode 0, non-lattice ratio: the replay runs as a warp-parallel scan; the distance is an integer
                                 // in units of 2^-scan_kb (the fine ulp of the binade below the power of two the sums r + ratio cross)
};

// per-pass scalars, identical for all nodes/channels of one depth (every stream of a depth has the same length)
struct PassInfo {
    int       n_new[32];         // samples produced at depth d in this pass
    int       wo[32];            // offset of the first new sample in the depth-d node buffers (0/1)
    long long out_count[32];     // channel outputs of depth-d channels already produced in this feed
    int       first_pass;
    int       parity;            // which half of the ping-pong history is current
    int       fused_pass;        // the tree ran as hb48_fused_kernel: channels marked direct already hold their outputs
};

// Schedule.  The reference's float32 recurrence (interpolator.h:23-36 + nfmdemod.cpp:315): per input d -= 1; if d < 1 emit
// at phase max(0, floor(d * steps)) and d += ratio.  The -1 steps are exact in float32, so one iteration per OUTPUT is
// bit-identical: j = max(1, floor d) inputs later, r = d - j, then d = fl(r + ratio).
//   * General ratios: one thread per channel replays that loop sequentially (exact, but a serial chain per channel).
//   * "Lattice" ratios: when ratio * 2^23 is a multiple of the coarsest ulp any sum r + ratio can have, every sum is exactly
//     representable, fl() never rounds, and the recurrence has the closed form  E_k = D + k*A  (units of 2^-23):
//     output k sits at input i0 + (E_k >> 23) - 1 with r = E_k & (2^23-1).  The thread then only runs the loop while d < 1
//     (stream start) and leaves the rest to frontend_kernel, which evaluates E_k per output: no serial chain at all
//     (1.25 = 60 kS/s -> 48 kS/s, the 1024-channel plan, is such a ratio).
// Depends only on counts, not on samples: runs on a high-priority side stream concurrently with the tree kernels.
// One WARP per channel: lane 0 runs the serial recurrence (the other lanes idle), then all lanes fill the tile table.
// ---- exact parallel replay of the float32 distance recurrence for non-lattice ratios ------------------------------------
// In units u = 2^-kb the distance before an emission is an integer D in [A, A + q), q = 2^kb = 1.0, A = ratio / u.  An
// emission takes j = D >> kb inputs, leaves R = D mod q, and the next distance is fl(R u + ratio): R + A exactly when that is
// below the power of two P the interval [ratio, ratio + 1) straddles (P / u = 2^24), else R + A rounded to a multiple of 2
// units, ties to even: an odd sum becomes the neighbouring multiple of 4.  So
//   * without the +-1 corrections R_k = (R_0 + k A) mod q in closed form, and "the sum crosses P" (R_k + A >= 2^24) too;
//   * whether a crossing step corrects, and by which sign, depends only on D mod 4, which follows a 4-state automaton
//     driven by the crossing flags: s' = (s + A) mod 4, and 0 after a correction -- a scan over function composition;
//   * the corrections accumulate in a small offset e_k (a prefix sum), R_k = closed form + e_k.
// The closed-form crossing flags are then checked against the corrected R_k; the rare chunk where an offset moves a value
// across a boundary is redone serially from its (exact) starting state.
__device__ __forceinline__ unsigned fe_fcomp(unsigned g, unsigned f)       // first g, then f; functions on {0..3}, 2 bits per entry
{
    unsigned h = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) h |= ((f >> (2 * ((g >> (2 * s)) & 3u))) & 3u) << (2 * s);
    return h;
}

// One super-step: FE_SW warps of a CTA take FE_SW consecutive chunks of 32 * FE_EPL candidate emissions (FE_EPL consecutive
// emissions per lane) from the exact state (Db, ib = last consumed input, nb = emissions so far).  Everything a chunk needs
// from the chunks before it is a composition / sum of per-chunk totals exchanged through shared memory: the automaton's
// composed step function (-> D mod 4 at the chunk's start), the corrections' sum (-> the offset e), the inputs taken (-> the
// input index).  A single warp's chunk is latency-bound (dependent integer chains, ~12 cycles per emission), so the warps of a
// CTA multiply the rate; closed-form residues are recomputed where needed instead of being kept in registers.
// On return sh.nv[w] = valid emissions of chunk w (a prefix of the super-step; 0 from the first inconsistent chunk sh.wf on),
// sh.D[w] / sh.i[w] = the exact state after chunk w's last valid emission.
constexpr int FE_EPL = 16;
constexpr int FE_CHUNK = 32 * FE_EPL;
constexpr int FE_SW = 8;                       // warps per channel

struct FeScanShared {
    unsigned fn[FE_SW];
    int csum[FE_SW], jsum[FE_SW], ok[FE_SW], nv[FE_SW], i[FE_SW];
    long long D[FE_SW];
    int wf;
};

__device__ __forceinline__ void fe_scan_super(const FrontendChan& c, int lane, int warp, int m, long long Db, int ib, int nb, int* sched, FeScanShared& sh)
{
    const int kb = c.scan_kb;
    const long long q = 1ll << kb, qm = q - 1, P = 1ll << 24, A = c.A;
    const long long Am = A & qm, R0 = Db & qm;
    const unsigned a4 = (unsigned) (A & 3);
    unsigned FL = 0, FH = 0;                   // the automaton's two step functions
#pragma unroll
    for (unsigned s2 = 0; s2 < 4; ++s2) {
        const unsigned t = (s2 + a4) & 3u;
        FL |= t << (2 * s2);
        FH |= ((t & 1u) ? 0u : t) << (2 * s2);
    }
    const long long kbase = (long long) FE_CHUNK * warp + (long long) FE_EPL * lane;
    auto Rz = [&](long long k) -> long long { return (R0 + k * Am) & qm; };      // closed-form R of emission k of the super-step
    // pass 1: crossing flags and the composed step function of this lane's emissions
    unsigned himask = 0, fn = 0xE4u;           // identity
#pragma unroll
    for (int t = 0; t < FE_EPL; ++t) {
        const bool hi = (Rz(kbase + t) + A >= P);
        himask |= (hi ? 1u : 0u) << t;
        fn = fe_fcomp(fn, hi ? FH : FL);
    }
    unsigned inc = fn;                         // inclusive scan of the step functions over the lanes
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc = fe_fcomp(o, inc);
    }
    unsigned exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = 0xE4u;
    if (lane == 31) sh.fn[warp] = inc;
    __syncthreads();
    unsigned g = 0xE4u;                        // the chunks before this one
    for (int w = 0; w < warp; ++w) g = fe_fcomp(g, sh.fn[w]);
    const unsigned s_chunk = (g >> (2 * (unsigned) (Db & 3))) & 3u;
    unsigned st = (exc >> (2 * s_chunk)) & 3u;                      // D mod 4 before this lane's first emission
    // pass 2: the corrections (+1 / -1 as two bit masks) and their sum
    unsigned cpos = 0, cneg = 0;
    int csum = 0;
#pragma unroll
    for (int t = 0; t < FE_EPL; ++t) {
        const unsigned tt = (st + a4) & 3u;
        const bool corr = ((himask >> t) & 1u) && (tt & 1u);
        if (corr) { if (tt == 1u) { cneg |= 1u << t; --csum; } else { cpos |= 1u << t; ++csum; } }
        st = corr ? 0u : tt;
    }
    int einc = csum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, einc, off);
        if (lane >= off) einc += o;
    }
    if (lane == 31) sh.csum[warp] = einc;
    __syncthreads();
    int e0 = einc - csum;                      // corrections before this lane's first emission
    for (int w = 0; w < warp; ++w) e0 += sh.csum[w];
    // pass 3: distances D_k, inputs taken j_k, true R_k; consistency of the closed-form flags
    bool ok = true;
    int jsum = 0, jj[FE_EPL];
    {
        int e = e0;
        long long before = (kbase > 0) ? Rz(kbase - 1) : 0;          // closed-form R of the emission before this lane's first
#pragma unroll
        for (int t = 0; t < FE_EPL; ++t) {
            const long long rz = Rz(kbase + t);
            const long long Dk = (kbase == 0 && t == 0) ? Db : before + A + e;    // (R0_{k-1} + A) + e_k
            const long long Rk = rz + e;
            ok = ok && Rk >= 0 && Rk < q && ((Rk + A >= P) == (((himask >> t) & 1u) != 0)) && ((Dk & qm) == Rk);
            jj[t] = (int) (Dk >> kb);
            jsum += jj[t];
            e += (int) ((cpos >> t) & 1u) - (int) ((cneg >> t) & 1u);
            before = rz;
        }
    }
    const bool ok_warp = __all_sync(0xffffffffu, ok);
    int jinc = jsum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, jinc, off);
        if (lane >= off) jinc += o;
    }
    if (lane == 31) sh.jsum[warp] = jinc;
    if (lane == 0) sh.ok[warp] = ok_warp ? 1 : 0;
    __syncthreads();
    int wf = FE_SW;
    for (int w = FE_SW - 1; w >= 0; --w) if (!sh.ok[w]) wf = w;
    // pass 4: the schedule entries; the valid emissions are a prefix, the lane that owns the last one publishes the state
    // after it: D = (closed-form R + A) + e_{k+1}
    int idx = ib + (jinc - jsum);
    for (int w = 0; w < warp; ++w) idx += sh.jsum[w];
    const float steps = (float) c.phase_steps, uf = 1.0f / (float) q;
    int cnt = 0, in_ = 0;
    long long Dn = 0;
    if (warp < wf) {
        int e = e0;
#pragma unroll
        for (int t = 0; t < FE_EPL; ++t) {
            idx += jj[t];
            const long long rz = Rz(kbase + t);
            const long long Rk = rz + e;
            e += (int) ((cpos >> t) & 1u) - (int) ((cneg >> t) & 1u);
            if (idx < m) {
                int ph = (int) floorf(__fmul_rn(__fmul_rn((float) Rk, uf), steps));      // (float) R u is exact
                ph = ph < 0 ? 0 : ph;
                sched[nb + (int) kbase + t] = (int) (((unsigned) idx << 8) | (unsigned) ph);
                ++cnt;
                Dn = rz + A + e; in_ = idx;
            }
        }
    }
    int tot = cnt;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
    const int owner = (tot > 0) ? (tot - 1) / FE_EPL : 0;
    Dn = __shfl_sync(0xffffffffu, Dn, owner);
    in_ = __shfl_sync(0xffffffffu, in_, owner);
    if (lane == 0) { sh.nv[warp] = tot; sh.D[warp] = Dn; sh.i[warp] = in_; }
    if (threadIdx.x == 0) sh.wf = wf;
    __syncthreads();
}

// One CTA (FE_SW warps) per channel.
__global__ void __launch_bounds__(32 * FE_SW) frontend_schedule_kernel(const FrontendChan* __restrict__ chans, int n_chans, const PassInfo pi)
{
    __shared__ FeScanShared sh;
    __shared__ long long sh_res[4];
    const int ch = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (ch >= n_chans) return;
    const FrontendChan c = chans[ch];
    const int m = pi.n_new[c.depth];
    int* __restrict__ sched = c.sched;
    int* __restrict__ tstart = c.tile_start;   // tstart[t] = first output whose input index is >= t * FE_TILE
    const int ntiles = (m + FE_TILE - 1) / FE_TILE;
    int n = 0;
    long long ncf = 0, i0 = 0, D = 0;
    if (c.scan_kb > 0 && c.mode == 0) {
        // non-lattice ratio: serial until the distance has been through one fl(r + ratio) (a few emissions at a stream's
        // start), then FE_SW * 32 * FE_EPL emissions per super-step.  Every thread carries the same (d, i, n).
        float d = __int_as_float(c.state[1]);
        const float steps = (float) c.phase_steps, ratio = c.ratio;
        int i = -1;
        bool done = false;
        auto serial_emit = [&]() -> bool {              // every thread computes, thread 0 writes: returns false when the input runs out
            const float fj = fmaxf(floorf(d), 1.0f);
            if (i + (int) fj >= m) return false;
            i += (int) fj;
            const float r = __fsub_rn(d, fj);
            int ph = (int) floorf(__fmul_rn(r, steps));
            ph = ph < 0 ? 0 : ph;
            if (tid == 0) sched[n] = (int) (((unsigned) i << 8) | (unsigned) ph);
            ++n;
            d = __fadd_rn(r, ratio);
            return true;
        };
        for (int k = 0; k < 2 && !done; ++k) done = !serial_emit();
        while (!done && !(d >= ratio && d < ratio + 1.0f)) done = !serial_emit();
        const float qf = (float) (1ll << c.scan_kb);
        while (!done) {
            const long long Db = (long long) (d * qf);          // exact: d is a multiple of 2^-kb below 2^25
            fe_scan_super(c, lane, warp, m, Db, i, n, sched, sh);
            const int wf = sh.wf;
            int total = 0, wl = -1;
            bool short_ = false;
            for (int w = 0; w < wf; ++w) {
                total += sh.nv[w];
                if (sh.nv[w] > 0) wl = w;
                if (sh.nv[w] < FE_CHUNK) short_ = true;
            }
            if (wl >= 0) { n += total; i = sh.i[wl]; d = (float) sh.D[wl] / qf; }
            __syncthreads();                                    // the exchange arrays are free for the next super-step
            if (short_) done = true;
            else if (wf < FE_SW) {
                for (int k = 0; k < FE_CHUNK && !done; ++k) done = !serial_emit();     // rare: an offset moved a value across a boundary
            }
        }
        d = __fadd_rn(d, -(float) (m - 1 - i));        // the remaining inputs of this pass each subtract 1.0f (exact)
        if (tid == 0) {
            c.plan[0] = n; c.plan[1] = 0; c.plan[2] = 0; c.plan[3] = 0;
            const int base = pi.first_pass ? 0 : c.state[3];
            c.state[1] = __float_as_int(d);
            c.state[2] = n;
            c.state[3] = base + n;
        }
    } else
    if (tid == 0) {
        float d = __int_as_float(c.state[1]);      // distance remain before the next input
        const float steps = (float) c.phase_steps, ratio = c.ratio;
        int i = -1;                                // index of the last consumed input
        // one emission: consumes j = max(1, floor(d)) inputs.  The loop-carried chain is floor -> max -> sub -> add.
#define FE_EMIT()                                                                              \
        {                                                                                      \
            const float fj = fmaxf(floorf(d), 1.0f);                                           \
            i += (int) fj;                                                                     \
            const float r = __fsub_rn(d, fj);                                                  \
            int ph = (int) floorf(__fmul_rn(r, steps));                                        \
            ph = ph < 0 ? 0 : ph;                                                              \
            sched[n++] = (int) (((unsigned) i << 8) | (unsigned) ph);   /* i < 2^24, ph < 256 */ \
            d = __fadd_rn(r, ratio);                                                           \
        }
        if (c.mode != 0) {
            // Interpolator::interpolate / resample in their callers' loops (interpolator.h:39-76; nfmmod.cpp:126-133): per OUTPUT
            // one iteration; `-= 1.0` is exact for d in [1, 2^24), `+= ratio` and the phase product are float32 operations
            int over = 0;
            if (c.mode == 1) {
                for (;;) {
                    if (d >= 1.0f) {
                        if (i + 1 >= m) break;
                        ++i;
                        d = __fsub_rn(d, 1.0f);
                    }
                    if (n >= c.sched_cap) { over = 1; break; }
                    int ph = (int) floorf(__fmul_rn(d, steps));
                    ph = ph < 0 ? 0 : ph;
                    sched[n++] = (int) (((unsigned) (i + 1) << 8) | (unsigned) ph);
                    d = __fadd_rn(d, ratio);
                }
            } else {
                for (int ii = 0; ii < m && !over; ++ii) {
                    bool consumed = false;
                    do {
                        bool ok = true;
                        while (d >= 1.0f) {
                            if (!consumed) { i = ii; d = __fsub_rn(d, 1.0f); consumed = true; }
                            else { ok = false; break; }
                        }
                        if (ok) {
                            if (n >= c.sched_cap) { over = 1; break; }
                            int ph = (int) floorf(__fmul_rn(d, steps));
                            ph = ph < 0 ? 0 : ph;
                            sched[n++] = (int) (((unsigned) ((consumed ? ii : ii - 1) + 1) << 8) | (unsigned) ph);
                            d = __fadd_rn(d, ratio);
                        }
                    } while (!consumed);
                }
            }
            c.plan[0] = n; c.plan[1] = 0; c.plan[2] = 0; c.plan[3] = 0;
            const int base = pi.first_pass ? 0 : c.state[3];
            c.state[0] = over;
            c.state[1] = __float_as_int(d);
            c.state[2] = n;
            c.state[3] = base + n;
        } else {
        if (!c.lattice) {
            // an emission consumes at most max(1, floor(d)) <= max(d0, ratio + 1) inputs: 8 at a time while that is safe
            const int per8 = 8 * ((int) fmaxf(ratio + 1.0f, d) + 1);
            while (m - 1 - i > per8 && d < ratio + 1.0f) {
#pragma unroll
                for (int k = 0; k < 8; ++k) FE_EMIT();
            }
        }
        for (;;) {
            if (c.lattice && d >= 1.0f) break;     // the closed form takes over
            const int j = (int) fmaxf(floorf(d), 1.0f);
            if (i + j >= m) break;
            FE_EMIT();
        }
#undef FE_EMIT
        if (c.lattice && d >= 1.0f) {
            // closed form from here: i0 = inputs consumed so far, D = d in units of 2^-23 (exact: d is on the lattice)
            const long long A = c.A;
            i0 = i + 1; D = (long long) (d * 8388608.0f);
            const long long lim = ((long long) (m - i0) + 1) << 23;            // E_k < lim  <=>  input index < m
            ncf = (lim - 1 - D >= 0) ? (lim - 1 - D) / A + 1 : 0;
            long long Dend;
            if (ncf > 0) {
                const long long El = D + (ncf - 1) * A;
                const long long idx_last = i0 + (El >> 23) - 1;
                Dend = (El & 0x7fffffll) + A - (((long long) m - 1 - idx_last) << 23);
            } else {
                Dend = D - (((long long) m - i0) << 23);
            }
            d = (float) Dend * (1.0f / 8388608.0f);                            // exact: a value the float recurrence would hold
        } else {
            d = __fadd_rn(d, -(float) (m - 1 - i));    // the remaining inputs of this pass each subtract 1.0f (exact)
        }
        c.plan[0] = n; c.plan[1] = i0; c.plan[2] = D; c.plan[3] = ncf;
        const int total = n + (int) ncf;
        const int base = pi.first_pass ? 0 : c.state[3];
        c.state[1] = __float_as_int(d);
        c.state[2] = total;
        c.state[3] = base + total;
        }
    }
    // (the serial forms ran on thread 0 only: hand its results to the CTA; the scan form left them in every thread)
    if (!(c.scan_kb > 0 && c.mode == 0)) {
        if (tid == 0) { sh_res[0] = n; sh_res[1] = ncf; sh_res[2] = i0; sh_res[3] = D; }
        __syncthreads();
        n = (int) sh_res[0]; ncf = sh_res[1]; i0 = sh_res[2]; D = sh_res[3];
    } else {
        __syncthreads();                       // thread 0's schedule entries (serial prologue / replay) are visible to the binary searches below
    }
    // tile table, one tile per thread at a time: listed outputs by binary search, closed-form outputs by division
    // (entry 1 is written even for a pass without new samples: block 0 of the front-end kernel always reads entries 0 and 1)
    for (int t = tid; t <= (ntiles > 1 ? ntiles : 1); t += blockDim.x) {
        const int T = t * FE_TILE;
        int lo = 0, hi = n;
        const int bias = c.mode ? 1 : 0;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int) ((unsigned) sched[mid] >> 8) - bias < T) lo = mid + 1; else hi = mid; }
        int o = (t == 0) ? 0 : lo;                 // (an output that precedes the pass's first input belongs to the first tile)
        if (lo == n && ncf > 0) {
            const long long need = (((long long) T - i0 + 1) << 23) - D;
            long long k = need <= 0 ? 0 : (need + c.A - 1) / c.A;
            if (k > ncf) k = ncf;
            o = n + (int) k;
        }
        tstart[t] = o;
    }
}

// grid = (tiles, channels).  A CTA mixes FE_TILE (+ ntaps-1 halo) channel samples with the table NCO into shared memory
// and computes every output whose newest input lies in the tile.
// LAT54 = every channel resamples by exactly 5/4 on the closed-form schedule with 72 taps per phase (ratio 1.25, the
// 60 kS/s -> 48 kS/s case of the 1024-channel plan): output o+4 then sits exactly 5 inputs after output o at the same phase,
// so a thread keeps one phase row of taps in registers and computes outputs o, o+4, o+8, o+12 in one sweep over the 87
// samples they span: 1 shared-memory load per 8 FFMA instead of 5 per 8 (same accumulation order: bit-identical results).
constexpr int FE_Z_EXTRA = 16;            // z entries past the tile that sweep may touch for outputs it does not store

template<bool LAT54>
__global__ void __launch_bounds__(FE_THREADS) frontend_kernel_t(const FrontendChan* __restrict__ chans, const float* __restrict__ nco_table, const PassInfo pi)
{
    extern __shared__ float fe_smem[];                     // taps [phase_steps*ntaps] then z [FE_MAX_TAPS + FE_TILE] float2
    FrontendChan c = chans[blockIdx.y];
    const int tid = threadIdx.x;
    const int m = pi.n_new[c.depth];
    const int t0 = blockIdx.x * FE_TILE;
    if (t0 >= m && !(blockIdx.x == 0)) return;
    c.in += pi.out_count[c.depth] * (c.in_f32 ? 2 : 1);
    const uint32_t* hin = c.hist + pi.parity * c.hist_stride;
    const int nt = c.ntaps, nts = (c.ntaps + 2 * FE_PAD) | 1;   // zero-padded rows, odd stride: the phases start in distinct banks
    const int ntp = nts * c.phase_steps;
    float* taps = fe_smem;
    float2* z = reinterpret_cast<float2*>(fe_smem + ((ntp + 3) & ~3));
    for (int ph = tid >> 5; ph < c.phase_steps; ph += FE_THREADS / 32)        // a warp per phase row: no division by the row length
        for (int kk = tid & 31; kk < nts; kk += 32) {
            const int k = kk - FE_PAD;
            taps[ph * nts + kk] = (k >= 0 && k < nt) ? c.taps[ph * nt + k] : 0.0f;
        }
    const unsigned phase0 = c.in_f32 ? 0u : hin[FE_MAX_TAPS];
    const int t1 = (t0 + FE_TILE < m) ? t0 + FE_TILE : m;
    // z[k] holds mixed sample (t0 - FE_MAX_TAPS + k), k in [0, FE_MAX_TAPS + t1 - t0)
    for (int k = tid; k < FE_MAX_TAPS + (t1 - t0); k += FE_THREADS) {
        const int i = t0 - FE_MAX_TAPS + k;
        if (c.in_f32) {
            z[k] = (i >= 0) ? reinterpret_cast<const float2*>(c.in)[i] : reinterpret_cast<const float2*>(hin)[FE_MAX_TAPS + i];
            continue;
        }
        const uint32_t w = (i >= 0) ? c.in[i] : hin[FE_MAX_TAPS + i];
        const float x = (float) (short) (w & 0xffffu), y = (float) ((int) w >> 16);
        const int p = (int) ((phase0 + (unsigned) (i + 1) * (unsigned) c.inc) & 4095u);   // phase advanced before the lookup; wrap == mod 4096
        const float u = nco_table[p], v = -nco_table[(p + 1024) & 4095];
        z[k] = make_float2(x * u - y * v, x * v + y * u);
    }
    __syncthreads();
    // outputs of this pass whose input index falls in [t0, t1)
    const int out_base = c.state[3] - c.state[2];
    const int o0 = c.tile_start[blockIdx.x], o1 = (t0 < m || c.mode != 0) ? c.tile_start[blockIdx.x + 1] : o0;    // Interpolator::decimate: no new samples, no outputs
    const int k0 = (int) c.plan[0];
    const long long pi0 = c.plan[1], D0 = c.plan[2];
    // LAT54: the generic loop only takes the outputs listed by the schedule kernel (stream start, before the closed form)
    const int o1g = LAT54 ? ((k0 < o0) ? o0 : (k0 < o1 ? k0 : o1)) : o1;
    // FE_NOUT consecutive outputs per thread: one pass over the z samples they share, newest first
    for (int ob = o0 + FE_NOUT * tid; ob < o1g; ob += FE_NOUT * FE_THREADS) {
        int idx[FE_NOUT];
        const float* trow[FE_NOUT];
#pragma unroll
        for (int q = 0; q < FE_NOUT; ++q) {
            const int o = (ob + q < o1g) ? ob + q : o1g - 1;        // clamp: duplicates are computed but not stored
            int ph;
            if (c.lattice && o >= k0) {                    // closed-form region (lattice ratios)
                const long long E = D0 + (long long) (o - k0) * c.A;
                idx[q] = (int) (pi0 + (E >> 23) - 1);
                ph = (int) ((E & 0x7fffffll) >> c.phshift);
            } else {
                const unsigned s = (unsigned) c.sched[o];
                idx[q] = (int) (s >> 8) - (c.mode ? 1 : 0); ph = (int) (s & 0xffu);
            }
            trow[q] = taps + ph * nts + FE_PAD;
        }
        float ra[FE_NOUT], ia[FE_NOUT];
#pragma unroll
        for (int q = 0; q < FE_NOUT; ++q) { ra[q] = 0.0f; ia[q] = 0.0f; }
        const int e_hi = idx[FE_NOUT - 1], e_lo = idx[0] - (nt - 1);
        if (e_hi - idx[0] <= FE_PAD) {
            // tap index of output q at z sample e is idx[q] - e; outside [0, nt) it lands in the zero padding
            const float* tq[FE_NOUT];
#pragma unroll
            for (int q = 0; q < FE_NOUT; ++q) tq[q] = trow[q] + (idx[q] - e_hi);
            const float2* zz = z + (e_hi - t0 + FE_MAX_TAPS);
            const int cnt = e_hi - e_lo + 1;
#pragma unroll 4
            for (int k = 0; k < cnt; ++k) {
                const float2 v = zz[-k];
#pragma unroll
                for (int q = 0; q < FE_NOUT; ++q) {
                    const float t = tq[q][k];
                    ra[q] = fmaf(t, v.x, ra[q]);
                    ia[q] = fmaf(t, v.y, ia[q]);
                }
            }
        } else {
            // widely spaced outputs (large decimation ratio): independent dot products
#pragma unroll
            for (int q = 0; q < FE_NOUT; ++q) {
                const float2* zz = z + (idx[q] - t0 + FE_MAX_TAPS);
                for (int k = 0; k < nt; ++k) {
                    const float2 v = zz[-k];
                    ra[q] = fmaf(trow[q][k], v.x, ra[q]);
                    ia[q] = fmaf(trow[q][k], v.y, ia[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < FE_NOUT; ++q) if (ob + q < o1g) c.out[out_base + ob + q] = make_float2(ra[q], ia[q]);
    }
    if (LAT54) {
        const int obl = (k0 > o0) ? k0 : o0;                           // first closed-form output of this tile
        const int ntask = (o1 > obl) ? ((o1 - obl + 15) / 16) * 4 : 0;  // task = (16 consecutive outputs, residue mod 4)
        for (int task = tid; task < ntask; task += FE_THREADS) {
            const int of = obl + (task & 3) + 16 * (task >> 2);        // outputs of, of+4, of+8, of+12
            if (of >= o1) continue;
            const long long E = D0 + (long long) (of - k0) * c.A;
            const int idx0 = (int) (pi0 + (E >> 23) - 1);
            const float* row = taps + (int) ((E & 0x7fffffll) >> c.phshift) * nts + FE_PAD;
            float t[72];
#pragma unroll
            for (int k = 0; k < 72; ++k) t[k] = row[k];
            float ra[4] = { 0.0f, 0.0f, 0.0f, 0.0f }, ia[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
            const float2* zz = z + (idx0 + 15 - t0 + FE_MAX_TAPS);    // newest sample of output of+12
#pragma unroll
            for (int sidx = 0; sidx < 87; ++sidx) {
                const float2 v = zz[-sidx];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = 5 * q - 15 + sidx;                   // tap of output q at this sample (compile-time)
                    if (k >= 0 && k < 72) { ra[q] = fmaf(t[k], v.x, ra[q]); ia[q] = fmaf(t[k], v.y, ia[q]); }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) if (of + 4 * q < o1) c.out[out_base + of + 4 * q] = make_float2(ra[q], ia[q]);
        }
    }
    // the CTA of the last tile carries the newest FE_MAX_TAPS channel samples and the NCO phase to the next pass
    if (t1 == m && (t0 < m || blockIdx.x == 0)) {
        uint32_t* hout = c.hist + (pi.parity ^ 1) * c.hist_stride;
        if (c.in_f32) {
            for (int k = tid; k < FE_MAX_TAPS; k += FE_THREADS) {
                const int i = m - FE_MAX_TAPS + k;
                reinterpret_cast<float2*>(hout)[k] = (i >= 0) ? reinterpret_cast<const float2*>(c.in)[i] : reinterpret_cast<const float2*>(hin)[FE_MAX_TAPS + i];
            }
        } else {
            for (int k = tid; k < FE_MAX_TAPS; k += FE_THREADS) {
                const int i = m - FE_MAX_TAPS + k;
                hout[k] = (i >= 0) ? c.in[i] : hin[FE_MAX_TAPS + i];
            }
            if (tid == 0) hout[FE_MAX_TAPS] = (phase0 + (unsigned) m * (unsigned) c.inc) & 4095u;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// frontend54_kernel: every channel a 5/4 closed-form resampler with 72 taps per phase (ratio 1.25: the 60 kS/s -> 48 kS/s
// channels of the 1024-channel plan).  Output k' of the closed form sits at E = D + k' A (units of 2^-23 inputs), A = 5 * 2^21:
// four outputs later it is exactly five inputs later at the same phase, so only four of the sixteen phases occur in a pass.
//   * grid = (blocks of 4608 outputs, channels); a block mixes the 5760 (+ 72) channel samples its outputs span with the
//     table NCO into shared memory once;
//   * warp w of the block owns phase residue w: its 72 taps live in registers for the whole block (loaded once);
//   * a lane computes NINE consecutive outputs of that phase (45 inputs apart from the next lane's: an odd stride, so
//     the 64-bit shared loads of a warp hit 32 distinct banks) in one sweep over the 112 samples they span: one shared
//     load per 11.6 FFMA, accumulation in ascending tap order like the generic kernel;
//   * results go through a small padded staging tile so that the global stores are contiguous.
// Outputs listed by the schedule kernel before the closed form takes over (the stream's first output) and the carried
// state (newest 128 channel samples + NCO phase) are block 0's.
// ---------------------------------------------------------------------------------------------------------
constexpr int F54_NOUT = 7;                             // consecutive same-phase outputs per lane (35 inputs apart from the next lane's: odd)
constexpr int F54_ROUNDS = 4;
constexpr int F54_THREADS = 256;                        // warp w: phase residue w & 3, rounds 2 (w >> 2) and 2 (w >> 2) + 1
constexpr int F54_G = 32 * F54_ROUNDS;                  // lanes x rounds: groups of 4 * F54_NOUT outputs per block
constexpr int F54_RND = 32 * 4 * F54_NOUT;              // outputs per round
constexpr int F54_OPB = F54_RND * F54_ROUNDS;           // 3584 outputs per block
constexpr int F54_IPB = 5 * F54_NOUT * F54_G;           // 4480 inputs
constexpr int F54_ZN = F54_IPB + 72 + 8;                // mixed samples a block may touch
constexpr int F54_SROW = 4 * F54_NOUT + 1;              // staging row of one lane group, padded to an odd number of float2
constexpr int F54_STAGE = 32 * F54_SROW;
constexpr int F54_SMEM = (F54_ZN + 2 * F54_STAGE) * (int) sizeof(float2) + 4096 * (int) sizeof(float);

__device__ __forceinline__ float2 fe_mix(uint32_t w, unsigned phase0, int i, int inc, const float* __restrict__ nco_table)
{
    const float x = (float) (short) (w & 0xffffu), y = (float) ((int) w >> 16);
    const int p = (int) ((phase0 + (unsigned) (i + 1) * (unsigned) inc) & 4095u);   // phase advanced before the lookup; wrap == mod 4096
    const float u = nco_table[p], v = -nco_table[(p + 1024) & 4095];
    return make_float2(x * u - y * v, x * v + y * u);
}

__global__ void __launch_bounds__(F54_THREADS, 2) frontend54_kernel(const FrontendChan* __restrict__ chans, const float* __restrict__ nco_table, const PassInfo pi)
{
    extern __shared__ float2 f54_smem[];
    float2* z = f54_smem;
    float2* stage = f54_smem + F54_ZN;
    float* nco = reinterpret_cast<float*>(f54_smem + F54_ZN + 2 * F54_STAGE);      // the NCO table: the mix gathers two entries per sample
    const FrontendChan c = chans[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, cc = (tid >> 5) & 3, quad = tid >> 7;
    const int m = pi.n_new[c.depth];
    const uint32_t* __restrict__ in = c.in + pi.out_count[c.depth];
    const uint32_t* __restrict__ hin = c.hist + pi.parity * c.hist_stride;
    const unsigned phase0 = hin[FE_MAX_TAPS];
    const int k0 = (int) c.plan[0], ncf = (int) c.plan[3];
    const long long i0 = c.plan[1], D0 = c.plan[2];
    const int out_base = c.state[3] - c.state[2];
    if (blockIdx.x == 0) {
        uint32_t* hout = c.hist + (pi.parity ^ 1) * c.hist_stride;
        if (tid < FE_MAX_TAPS) {
            const int i = m - FE_MAX_TAPS + tid;
            hout[tid] = (i >= 0) ? in[i] : hin[FE_MAX_TAPS + i];
            if (tid == 0) hout[FE_MAX_TAPS] = (phase0 + (unsigned) m * (unsigned) c.inc) & 4095u;
        }
        for (int o = tid; o < k0; o += F54_THREADS) {          // before the closed form: straight from global memory (a handful per stream)
            const unsigned s = (unsigned) c.sched[o];
            const int idx = (int) (s >> 8);
            const float* row = c.taps + (s & 0xffu) * c.ntaps;
            float ra = 0.0f, ia = 0.0f;
            for (int k = 0; k < c.ntaps; ++k) {
                const int i = idx - k;
                const float2 v = fe_mix((i >= 0) ? in[i] : hin[FE_MAX_TAPS + i], phase0, i, c.inc, nco_table);
                ra = fmaf(row[k], v.x, ra);
                ia = fmaf(row[k], v.y, ia);
            }
            c.out[out_base + o] = make_float2(ra, ia);
        }
    }
    const int kb = blockIdx.x * F54_OPB;
    if (kb >= ncf) return;
#pragma unroll
    for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(nco)[tid + F54_THREADS * q] = __ldg(reinterpret_cast<const float4*>(nco_table) + tid + F54_THREADS * q);
    const long long Eb = D0 + (long long) kb * c.A;
    const int zlo = (int) (i0 + (Eb >> 23) - 1) - 71;          // oldest sample any output of this block taps (>= -71: inside the history)
    // samples up to the newest one the block's last output taps (a channel's last block is usually partly empty)
    int zn = F54_ZN;
    {
        const int klast = ((ncf - kb < F54_OPB) ? ncf - kb : F54_OPB) - 1;
        const int need = (int) (i0 + ((Eb + (long long) klast * c.A) >> 23) - 1) - zlo + 1 + 8;
        if (need < zn) zn = need;
    }
    const float4* taprow;
    int idx_c;
    {
        const long long Ec = Eb + (long long) cc * c.A;
        idx_c = (int) (i0 + (Ec >> 23) - 1);
        taprow = reinterpret_cast<const float4*>(c.taps + (int) ((Ec & 0x7fffffll) >> c.phshift) * 72);      // 72 floats per phase: 16-byte aligned rows
    }
    __syncthreads();
    // mix: all loads of eight samples in flight; the common case (every index inside this pass) has no selects
    const bool inside = (zlo >= 0) && (zlo + zn + 8 * F54_THREADS <= m);
    for (int kk = tid; kk < zn; kk += F54_THREADS * 8) {
        uint32_t w[8];
        float u[8], v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = zlo + kk + F54_THREADS * q;
            if (inside) w[q] = __ldg(in + i);
            else {
                const int ii = (i < m) ? i : m - 1;
                w[q] = __ldg((ii >= 0) ? in + ii : hin + FE_MAX_TAPS + ii);
            }
            const int p = (int) ((phase0 + (unsigned) (i + 1) * (unsigned) c.inc) & 4095u);
            u[q] = nco[p]; v[q] = -nco[(p + 1024) & 4095];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k = kk + F54_THREADS * q, i = zlo + k;
            const float x = (float) (short) (w[q] & 0xffffu), y = (float) ((int) w[q] >> 16);
            if (k < F54_ZN) z[k] = (i < m) ? make_float2(x * u[q] - y * v[q], x * v[q] + y * u[q]) : make_float2(0.0f, 0.0f);
        }
    }
    float t[72];
#pragma unroll
    for (int k = 0; k < 18; ++k) { const float4 f = __ldg(taprow + k); t[4 * k] = f.x; t[4 * k + 1] = f.y; t[4 * k + 2] = f.z; t[4 * k + 3] = f.w; }
    __syncthreads();
#pragma unroll 1
    for (int step = 0; step < 2; ++step) {
        if (kb + F54_RND * step >= ncf && kb + F54_RND * (2 + step) >= ncf) break;       // block-uniform: neither quad has outputs left
        const int round = 2 * quad + step;
        const int g = lane + 32 * round;
        float ra[F54_NOUT], ia[F54_NOUT];
#pragma unroll
        for (int n = 0; n < F54_NOUT; ++n) { ra[n] = 0.0f; ia[n] = 0.0f; }
        if (kb + 4 * F54_NOUT * g + cc < ncf) {
            const float2* zz = z + (idx_c + 5 * F54_NOUT * g + 5 * (F54_NOUT - 1) - zlo);      // newest sample of this lane's last output
#pragma unroll
            for (int s = 0; s < 72 + 5 * (F54_NOUT - 1); ++s) {
                const float2 v = zz[-s];
#pragma unroll
                for (int n = 0; n < F54_NOUT; ++n) {
                    const int k = 5 * n - 5 * (F54_NOUT - 1) + s;            // tap of output n at this sample (compile-time)
                    if (k >= 0 && k < 72) { ra[n] = fmaf(t[k], v.x, ra[n]); ia[n] = fmaf(t[k], v.y, ia[n]); }
                }
            }
        }
        float2* st = stage + quad * F54_STAGE + lane * F54_SROW + cc;
#pragma unroll
        for (int n = 0; n < F54_NOUT; ++n) st[4 * n] = make_float2(ra[n], ia[n]);
        __syncthreads();
        // both quads' rounds go out as contiguous runs
        for (int o = tid; o < 2 * F54_RND; o += F54_THREADS) {
            const int qd = (o >= F54_RND) ? 1 : 0, oo = o - qd * F54_RND;
            const int r = oo / (4 * F54_NOUT);
            const int ko = kb + F54_RND * (2 * qd + step) + oo;
            if (ko < ncf) c.out[out_base + k0 + ko] = stage[qd * F54_STAGE + r * F54_SROW + (oo - r * 4 * F54_NOUT)];
        }
        __syncthreads();
    }
}

} // namespace b200dsp
