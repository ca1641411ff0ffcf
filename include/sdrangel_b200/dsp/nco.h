// nco.h — drop-in for NCO (sdrbase/dsp/nco.h:27-61, nco.cpp:30-64): same class name and methods; nextIQ() per sample is one GPU
// call (drop-in, not fast), nextIQ(out, n) is the block form (n successive nextIQ() values in one launch).
#ifndef SDRANGEL_B200_DSP_NCO_H
#define SDRANGEL_B200_DSP_NCO_H
#include "dsptypes.h"

class NCO {
public:
    NCO() : m_h(nullptr) { b200dsp_cxx::check(b200dsp_nco_create(&m_h)); }
    ~NCO() { b200dsp_nco_destroy(m_h); }
    void setPhase(int phase) { b200dsp_cxx::check(b200dsp_nco_set_phase(m_h, phase)); }
    void setFreq(Real freq, Real sampleRate) { b200dsp_cxx::check(b200dsp_nco_set_freq(m_h, freq, sampleRate)); }
    /** == nextPhase(); return Complex(table[phase], -table[(phase + 1024) % 4096]) */
    Complex nextIQ()
    {
        float v[2];
        b200dsp_cxx::check(b200dsp_nco_next_iq(m_h, 1, v));
        return Complex(v[0], v[1]);
    }
    /** == nextIQ().real() */
    Real next() { return nextIQ().real(); }
    Complex nextQI() { const Complex c = nextIQ(); return Complex(c.imag(), c.real()); }
    /** n successive nextIQ() values */
    void nextIQ(Complex* out, size_t n) { b200dsp_cxx::check(b200dsp_nco_next_iq(m_h, (int64_t) n, reinterpret_cast<float*>(out))); }
    int phase() const { int p = 0, i = 0; b200dsp_nco_get(m_h, &p, &i); return p; }
    int phaseIncrement() const { int p = 0, i = 0; b200dsp_nco_get(m_h, &p, &i); return i; }
    b200dsp_nco_t* handle() { return m_h; }
private:
    NCO(const NCO&);
    NCO& operator=(const NCO&);
    b200dsp_nco_t* m_h;
};
#endif
