// iqcorrections.h — the engine-side sample corrections of DSPDeviceSourceEngine (sdrbase/dsp/dspdevicesourceengine.cpp:175-262)
// as a small class with the reference method's name and arguments.  In the reference iqCorrections() is a private member of
// the engine that owns the m_iBeta / m_qBeta moving averages (dspdevicesourceengine.h:106-107); a maintainer replaces the
// body of that member by a call to this object.
#ifndef SDRANGEL_B200_DSP_IQCORRECTIONS_H
#define SDRANGEL_B200_DSP_IQCORRECTIONS_H
#include "dsptypes.h"

class IQCorrections {
public:
    IQCorrections() : m_h(nullptr) { b200dsp_cxx::check(b200dsp_iqcorr_create(&m_h)); }
    ~IQCorrections() { b200dsp_iqcorr_destroy(m_h); }
    IQCorrections(const IQCorrections&) = delete;
    IQCorrections& operator=(const IQCorrections&) = delete;

    /** == DSPDeviceSourceEngine::iqCorrections(begin, end, imbalanceCorrection): corrects [begin, end) in place.
     *  imbalanceCorrection selects the I/Q imbalance branch (:219-252) instead of DC correction only (:254-259). */
    void iqCorrections(SampleVector::iterator begin, SampleVector::iterator end, bool imbalanceCorrection)
    {
        if (begin == end) return;
        b200dsp_cxx::check(b200dsp_iqcorr_run(m_h, reinterpret_cast<int16_t*>(&(*begin)), end - begin, imbalanceCorrection ? 1 : 0));
    }
    void reset() { b200dsp_cxx::check(b200dsp_iqcorr_reset(m_h)); }
    b200dsp_iqcorr_t* handle() { return m_h; }

private:
    b200dsp_iqcorr_t* m_h;
};
#endif
