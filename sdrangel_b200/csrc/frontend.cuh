// frontend.cuh — K4: the channel plugin front-end, NCO mix + polyphase Interpolator::decimate, for every channel of a bank.
//
// Replaces (paths relative to the reference tree), per channel:
//   Complex c(re, im); c *= m_nco.nextIQ();                          plugins/channelrx/demodnfm/nfmdemod.cpp:152-153
//   m_interpolator.decimate(&m_interpolatorDistanceRemain, c, &ci)   sdrbase/dsp/interpolator.h:23-36,107-113,183-194
//   m_interpolatorDistanceRemain += m_interpolatorDistance           plugins/channelrx/demodnfm/nfmdemod.cpp:315
//   NCO::nextPhase/nextIQ (phase advanced BEFORE the lookup)         sdrbase/dsp/nco.h:43-50, nco.cpp:60-64
//
// B200 design: one CTA per channel per feed.  The reference's float32 "distance" recurrence decides which inputs emit
// an output and at which of the 16 phases; it depends only on the ratio, not on the data, so lane 0 of warp 0 replays it
// exactly (same float operations) while the other warps mix the channel's new samples with the table NCO (the phase of
// sample i is phase0 + (i+1)*inc, so the mix is data-parallel).  After a block barrier every thread computes outputs
// as 72-tap dot products  y = sum_k taps[phase][k] * z[idx-k]  with the taps in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200dsp {

constexpr int FE_MAX_TAPS = 128;          // taps per phase supported by the kernel's history area (reference default: 72)

struct FrontendChan {            // one per channel with a front-end, device array
    const uint32_t* in;          // channel samples of this feed (packed int16 IQ), m of them
    float2*         z;           // [FE_MAX_TAPS + cap] mixed samples; z[0..FE_MAX_TAPS) = history (newest at FE_MAX_TAPS-1)
    const float*    taps;        // [phase_steps][ntaps]
    float2*         out;         // outputs of this feed are written from out[out_base]
    int*            sched;       // [cap] packed (idx << 8 | phase) of this feed's outputs
    int*            state;       // [4]: nco phase, float distance remain (bits), outputs of the last pass, outputs of this feed
    int             m;           // new samples
    int             first_pass;  // 1: first pass of a feed (the feed's output count restarts at 0)
    int             inc;         // NCO phase increment
    int             ntaps, phase_steps;
    float           ratio;       // m_interpolatorDistance
};

__global__ void frontend_kernel(const FrontendChan* __restrict__ chans, const float* __restrict__ nco_table)
{
    extern __shared__ float fe_taps[];
    __shared__ int s_nout, s_base;
    const FrontendChan c = chans[blockIdx.x];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int m = c.m;
    for (int i = tid; i < c.ntaps * c.phase_steps; i += nthr) fe_taps[i] = c.taps[i];

    const int phase0 = c.state[0];
    if (tid == 0) {
        // exact replay of Interpolator::decimate's schedule (interpolator.h:23-36) and the caller's += (nfmdemod.cpp:315)
        float d = __int_as_float(c.state[1]);
        int n = 0;
        const float steps = (float) c.phase_steps;
        for (int i = 0; i < m; ++i) {
            d = __fadd_rn(d, -1.0f);
            if (d >= 1.0f) continue;
            int ph = (int) floorf(__fmul_rn(d, steps));
            if (ph < 0) ph = 0;
            c.sched[n++] = (i << 8) | ph;
            d = __fadd_rn(d, c.ratio);
        }
        const int base = c.first_pass ? 0 : c.state[3];
        c.state[1] = __float_as_int(d);
        c.state[2] = n;
        c.state[3] = base + n;
        s_base = base;
        int p = (int) (((long long) phase0 + (long long) m * c.inc) % 4096);
        if (p < 0) p += 4096;
        c.state[0] = p;
        s_nout = n;
    }
    // NCO mix of the new samples: phase_i = phase0 + (i+1)*inc  (mod 4096)
    for (int i = tid; i < m; i += nthr) {
        const uint32_t w = c.in[i];
        const float x = (float) (short) (w & 0xffffu), y = (float) ((int) w >> 16);
        int p = (int) (((long long) phase0 + (long long) (i + 1) * c.inc) % 4096);
        if (p < 0) p += 4096;
        const float u = nco_table[p], v = -nco_table[(p + 1024) & 4095];
        c.z[FE_MAX_TAPS + i] = make_float2(x * u - y * v, x * v + y * u);
    }
    __syncthreads();
    const int n = s_nout, out_base = s_base;
    const int nt = c.ntaps;
    for (int o = tid; o < n; o += nthr) {
        const int s = c.sched[o];
        const int idx = s >> 8, ph = s & 0xff;
        const float* t = fe_taps + ph * nt;
        const float2* zz = c.z + FE_MAX_TAPS + idx;
        float ra = 0.0f, ia = 0.0f;
#pragma unroll 8
        for (int k = 0; k < nt; ++k) {
            const float2 v = zz[-k];
            ra = fmaf(t[k], v.x, ra);
            ia = fmaf(t[k], v.y, ia);
        }
        c.out[out_base + o] = make_float2(ra, ia);
    }
    __syncthreads();
    // carry the last FE_MAX_TAPS mixed samples as the next feed's history
    float2 keep[(FE_MAX_TAPS + 127) / 128];
    int cnt = 0;
    for (int i = tid; i < FE_MAX_TAPS; i += nthr) keep[cnt++] = c.z[m + i];
    __syncthreads();
    cnt = 0;
    for (int i = tid; i < FE_MAX_TAPS; i += nthr) c.z[i] = keep[cnt++];
}

} // namespace b200dsp
