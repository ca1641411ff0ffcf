// interpolator.h — drop-in for the Rx use of Interpolator (sdrbase/dsp/interpolator.h:19-36, interpolator.cpp:74-129).
// The reference's decimate() is called once per input sample from the plugin's feed loop
// (plugins/channelrx/demodnfm/nfmdemod.cpp:150-155): a per-sample GPU call would be all launch latency, so the wrapper
// offers the loop itself as one call with the same state (the caller-owned `distance`) and the same results.
#ifndef SDRANGEL_B200_DSP_INTERPOLATOR_H
#define SDRANGEL_B200_DSP_INTERPOLATOR_H
#include "dsptypes.h"

class Interpolator {
public:
    Interpolator() : m_h(nullptr) {}
    ~Interpolator() { free(); }
    void create(int phaseSteps, double sampleRate, double cutoff, double nbTapsPerPhase = 4.5)
    {
        free();
        b200dsp_cxx::check(b200dsp_interp_create(&m_h, phaseSteps, sampleRate, cutoff, nbTapsPerPhase));
    }
    void free() { if (m_h) { b200dsp_interp_destroy(m_h); m_h = nullptr; } }
    /** == for (i < n) if (decimate(distance, in[i], &ci)) { out[m++] = ci; *distance += step; }   returns m */
    size_t decimate(Real* distance, Real step, const Complex* in, size_t n, Complex* out, size_t cap)
    {
        int64_t m = 0;
        b200dsp_cxx::check(b200dsp_interp_decimate(m_h, distance, step, reinterpret_cast<const float*>(in), (int64_t) n,
                                                   reinterpret_cast<float*>(out), (int64_t) cap, &m));
        return (size_t) m;
    }
private:
    Interpolator(const Interpolator&);
    Interpolator& operator=(const Interpolator&);
    b200dsp_interp_t* m_h;
};
#endif
