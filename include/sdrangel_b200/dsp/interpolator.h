// interpolator.h — drop-in for Interpolator (sdrbase/dsp/interpolator.h:14-76, interpolator.cpp:74-129): same class name,
// create/free and the three per-sample methods with the reference's signatures, plus the callers' loops as block calls.
//
// The per-sample methods (decimate / interpolate / resample with `const Complex& next`) make one GPU call per sample
// (b200dsp_interp_step): identical results and identical effect on the caller-owned *distance, so plugin code compiles and
// behaves unchanged -- but a launch per sample is all latency.  The block forms run the loop every plugin writes around the
// method as ONE call with the same state and the same results:
//   decimate(distance, step, in, n, out, cap)      plugins/channelrx/demodnfm/nfmdemod.cpp:150-155,315
//   interpolate(distance, step, in, n, out, cap)   plugins/channeltx/modnfm/nfmmod.cpp:126-133 (one call per output sample)
//   resample(distance, step, in, n, out, cap)      interpolator.h:55-76 in its do-while-per-input loop
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

    // ---- the reference's per-sample signatures (interpolator.h:23-76)
    bool decimate(Real* distance, const Complex& next, Complex* result)
    {
        int consumed = 0, produced = 0;
        step(0, distance, next, result, &consumed, &produced);
        return produced != 0;
    }
    bool interpolate(Real* distance, const Complex& next, Complex* result)
    {
        int consumed = 0, produced = 0;
        step(1, distance, next, result, &consumed, &produced);
        return consumed != 0;
    }
    bool resample(Real* distance, const Complex& next, bool* consumed, Complex* result)
    {
        int c = *consumed ? 1 : 0, produced = 0;
        step(2, distance, next, result, &c, &produced);
        *consumed = (c != 0);
        return produced != 0;
    }

    // ---- the callers' loops as one call each; return the number of outputs written
    /** == for (i < n) if (decimate(distance, in[i], &ci)) { out[m++] = ci; *distance += step; } */
    size_t decimate(Real* distance, Real step, const Complex* in, size_t n, Complex* out, size_t cap)
    {
        int64_t m = 0;
        b200dsp_cxx::check(b200dsp_interp_decimate(m_h, distance, step, reinterpret_cast<const float*>(in), (int64_t) n,
                                                   reinterpret_cast<float*>(out), (int64_t) cap, &m));
        return (size_t) m;
    }
    /** == for each output: if (interpolate(distance, in[i], &ci)) ++i; out[m++] = ci; *distance += step;  until in[n] would be needed */
    size_t interpolate(Real* distance, Real step, const Complex* in, size_t n, Complex* out, size_t cap)
    {
        int64_t m = 0;
        b200dsp_cxx::check(b200dsp_interp_interpolate(m_h, distance, step, reinterpret_cast<const float*>(in), (int64_t) n,
                                                      reinterpret_cast<float*>(out), (int64_t) cap, &m));
        return (size_t) m;
    }
    /** == for (i < n) { consumed = false; do { if (resample(distance, in[i], &consumed, &ci)) { out[m++] = ci; *distance += step; } } while (!consumed); } */
    size_t resample(Real* distance, Real step, const Complex* in, size_t n, Complex* out, size_t cap)
    {
        int64_t m = 0;
        b200dsp_cxx::check(b200dsp_interp_resample(m_h, distance, step, reinterpret_cast<const float*>(in), (int64_t) n,
                                                   reinterpret_cast<float*>(out), (int64_t) cap, &m));
        return (size_t) m;
    }
private:
    void step(int op, Real* distance, const Complex& next, Complex* result, int* consumed, int* produced)
    {
        const float nx[2] = { next.real(), next.imag() };
        float rs[2] = { result->real(), result->imag() };
        b200dsp_cxx::check(b200dsp_interp_step(m_h, op, distance, nx, rs, consumed, produced));
        if (*produced) *result = Complex(rs[0], rs[1]);
    }
    Interpolator(const Interpolator&);
    Interpolator& operator=(const Interpolator&);
    b200dsp_interp_t* m_h;
};
#endif
