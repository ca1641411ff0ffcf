// fftfilt.h — drop-in for fftfilt (sdrbase/dsp/fftfilt.h:20-48, fftfilt.cpp:49-360), the overlap-add FFT filter behind the
// SSB / DSB demodulators (ssbdemod.cpp:91-92,165-175), over b200dsp_fftfilt_*.  The reference's per-sample signatures are kept:
// a call returns 0 until len/2 samples have been collected, then len/2 and a pointer to the outputs (valid until the next
// block), exactly like the reference; the block goes to the GPU in one call.  The block forms take any number of samples.
#ifndef SDRANGEL_B200_DSP_FFTFILT_H
#define SDRANGEL_B200_DSP_FFTFILT_H
#include <vector>
#include "dsptypes.h"

class fftfilt {
public:
    typedef std::complex<float> cmplx;
    fftfilt(float f1, float f2, int len) : m_h(nullptr), flen2(len >> 1) { b200dsp_cxx::check(b200dsp_fftfilt_create(&m_h, 0, f1, f2, len)); init(); }
    fftfilt(float f2, int len) : m_h(nullptr), flen2(len >> 1) { b200dsp_cxx::check(b200dsp_fftfilt_create(&m_h, 1, 0.0f, f2, len)); init(); }
    ~fftfilt() { b200dsp_fftfilt_destroy(m_h); }
    fftfilt(const fftfilt&) = delete;
    fftfilt& operator=(const fftfilt&) = delete;
    void create_filter(float f1, float f2) { b200dsp_cxx::check(b200dsp_fftfilt_set_filter(m_h, 0, f1, f2)); }
    void create_dsb_filter(float f2) { b200dsp_cxx::check(b200dsp_fftfilt_set_filter(m_h, 1, 0.0f, f2)); }
    int runFilt(const cmplx& in, cmplx** out) { return push(0, false, true, in, out); }
    int runSSB(const cmplx& in, cmplx** out, bool usb, bool getDC = true) { return push(1, usb, getDC, in, out); }
    int runDSB(const cmplx& in, cmplx** out, bool getDC = true) { return push(2, false, getDC, in, out); }
    /** block form: n samples in, the outputs of every block completed by them in `out` (capacity cap); returns their number */
    int runSSB(const cmplx* in, int n, cmplx* out, int cap, bool usb, bool getDC = true) { return block(1, usb, getDC, in, n, out, cap); }
    int runDSB(const cmplx* in, int n, cmplx* out, int cap, bool getDC = true) { return block(2, false, getDC, in, n, out, cap); }
    int runFilt(const cmplx* in, int n, cmplx* out, int cap) { return block(0, false, true, in, n, out, cap); }
private:
    void init() { m_data.reserve((size_t) flen2); m_output.resize((size_t) flen2); }
    int push(int op, bool usb, bool getDC, const cmplx& in, cmplx** out)
    {
        m_data.push_back(in);                                           // data[inptr++] = in
        if ((int) m_data.size() < flen2) return 0;
        const int n = block(op, usb, getDC, &m_data[0], flen2, &m_output[0], flen2);
        m_data.clear();
        *out = &m_output[0];
        return n;
    }
    int block(int op, bool usb, bool getDC, const cmplx* in, int n, cmplx* out, int cap)
    {
        int64_t m = 0;
        b200dsp_cxx::check(b200dsp_fftfilt_run(m_h, op, usb ? 1 : 0, getDC ? 1 : 0, reinterpret_cast<const float*>(in), n, reinterpret_cast<float*>(out), cap, &m));
        return (int) m;
    }
    b200dsp_fftfilt_t* m_h;
    int flen2;
    std::vector<cmplx> m_data, m_output;
};
#endif
