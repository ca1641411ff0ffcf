// phasediscri.h — drop-in for PhaseDiscriminators (sdrbase/dsp/phasediscri.h:26-198) over b200dsp_demod_*.
// The reference's methods take one sample; a channel plugin restructured for the GPU calls the block forms on the whole
// output of Interpolator::decimate (nfmdemod.cpp:150-165).  The per-sample signatures are kept (one launch per sample) for
// code not yet restructured.  One discriminator kind per handle: the kinds do not share m_m1Sample the way the reference's
// single object would if a plugin mixed them (none does).
#ifndef SDRANGEL_B200_DSP_PHASEDISCRI_H
#define SDRANGEL_B200_DSP_PHASEDISCRI_H
#include "dsptypes.h"

class PhaseDiscriminators {
public:
    PhaseDiscriminators() : m_fmScaling(1.0f) { for (int k = 0; k < 3; k++) m_h[k] = nullptr; }
    ~PhaseDiscriminators() { for (int k = 0; k < 3; k++) b200dsp_demod_destroy(m_h[k]); }
    PhaseDiscriminators(const PhaseDiscriminators&) = delete;
    PhaseDiscriminators& operator=(const PhaseDiscriminators&) = delete;
    void reset() { for (int k = 0; k < 3; k++) if (m_h[k]) b200dsp_cxx::check(b200dsp_demod_reset(m_h[k])); }
    void setFMScaling(Real fmScaling)
    {
        m_fmScaling = fmScaling;
        for (int k = 0; k < 3; k++) if (m_h[k]) b200dsp_cxx::check(b200dsp_demod_set_fm_scaling(m_h[k], fmScaling));
    }
    // the reference's per-sample signatures
    Real phaseDiscriminator(const Complex& sample) { Real o; run(B200DSP_DEMOD_FM_ATAN2, &sample, 1, &o, nullptr, nullptr); return o; }
    Real phaseDiscriminatorDelta(const Complex& sample, double& magsq, Real& fmDev)
    {
        Real o, m;
        run(B200DSP_DEMOD_FM_DELTA, &sample, 1, &o, &m, &fmDev);
        magsq = m;
        return o;
    }
    Real phaseDiscriminator2(const Complex& sample) { Real o; run(B200DSP_DEMOD_FM_DISCRI2, &sample, 1, &o, nullptr, nullptr); return o; }
    // block forms: n samples in, n values out (magsq / fmDev may be null)
    void phaseDiscriminator(const Complex* samples, int n, Real* out) { run(B200DSP_DEMOD_FM_ATAN2, samples, n, out, nullptr, nullptr); }
    void phaseDiscriminatorDelta(const Complex* samples, int n, Real* out, Real* magsq, Real* fmDev) { run(B200DSP_DEMOD_FM_DELTA, samples, n, out, magsq, fmDev); }
    void phaseDiscriminator2(const Complex* samples, int n, Real* out) { run(B200DSP_DEMOD_FM_DISCRI2, samples, n, out, nullptr, nullptr); }
private:
    void run(int kind, const Complex* in, int n, Real* out, Real* a0, Real* a1)
    {
        if (!m_h[kind]) b200dsp_cxx::check(b200dsp_demod_create(&m_h[kind], kind, m_fmScaling, 1));
        b200dsp_cxx::check(b200dsp_demod_run(m_h[kind], reinterpret_cast<const float*>(in), n, out, a0, a1));
    }
    b200dsp_demod_t* m_h[3];
    Real m_fmScaling;
};
#endif
