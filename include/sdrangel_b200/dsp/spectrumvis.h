// spectrumvis.h — drop-in for SpectrumVis (sdrgui/dsp/spectrumvis.h, spectrumvis.cpp:77-254,283-327).
// The reference's only consumer is GLSpectrum::newSpectrum(const std::vector<Real>&, int); any object with that method
// can be plugged in as the display.
#ifndef SDRANGEL_B200_DSP_SPECTRUMVIS_H
#define SDRANGEL_B200_DSP_SPECTRUMVIS_H
#include "basebandsamplesink.h"

struct SpectrumDisplay { virtual ~SpectrumDisplay() {} virtual void newSpectrum(const std::vector<Real>& spectrum, int fftSize) = 0; };

class SpectrumVis : public BasebandSampleSink {
public:
    enum AveragingMode { AvgModeNone, AvgModeMoving, AvgModeFixed };
    SpectrumVis(Real scalef, SpectrumDisplay* glSpectrum = nullptr) : m_glSpectrum(glSpectrum), m_fftSize(1024), m_h(nullptr)
    { b200dsp_cxx::check(b200dsp_spectrum_create(&m_h, scalef)); }
    virtual ~SpectrumVis() { b200dsp_spectrum_destroy(m_h); }
    /** == configure(msgQueue, fftSize, overlapPercent, averagingNb, averagingMode, window, linear) -> handleConfigure */
    void configure(int fftSize, int overlapPercent, unsigned int averagingNb, int averagingMode, int window, bool linear)
    {
        b200dsp_cxx::check(b200dsp_spectrum_configure(m_h, fftSize, overlapPercent, averagingNb, averagingMode, window, linear ? 1 : 0));
        m_fftSize = fftSize > 4096 ? 4096 : (fftSize < 64 ? 64 : fftSize);
    }
    virtual void start() {}
    virtual void stop() {}
    virtual void feed(const SampleVector::const_iterator& begin, const SampleVector::const_iterator& end, bool positiveOnly)
    {
        if (!m_glSpectrum) return;                                                    // spectrumvis.cpp:81-84
        const int64_t n = end - begin;
        int64_t frames = b200dsp_spectrum_frames_for(m_h, n);
        m_frames.resize((size_t) (frames > 0 ? frames : 1) * m_fftSize);
        b200dsp_cxx::check(b200dsp_spectrum_feed(m_h, n > 0 ? (const int16_t*) &(*begin) : nullptr, n, positiveOnly ? 1 : 0, &m_frames[0], frames > 0 ? frames : 1, &frames));
        m_powerSpectrum.resize(m_fftSize);
        for (int64_t f = 0; f < frames; ++f) {
            m_powerSpectrum.assign(m_frames.begin() + f * m_fftSize, m_frames.begin() + (f + 1) * m_fftSize);
            m_glSpectrum->newSpectrum(m_powerSpectrum, m_fftSize);                    // spectrumvis.cpp:147,182,231
        }
    }
private:
    SpectrumDisplay* m_glSpectrum;
    int m_fftSize;
    b200dsp_spectrum_t* m_h;
    std::vector<Real> m_frames, m_powerSpectrum;
};
#endif
