// ORACLE (test infrastructure, never shipped, never on the product path).
//
// C-ABI driver around the UNMODIFIED reference DSP classes, compiled in place from
// /root/reference by oracle/Makefile into oracle/_ref/libsdrref.so.  No reference source is
// copied into this repository: this file only #includes the reference headers and calls them.
//
// What is reference code and what is glue:
//   * Decimators<>, DecimatorsFI/FF/IF, DownChannelizer, NCO, Interpolator, FFTWindow,
//     FFTEngine/KissEngine, MovingAverage2D, FixedAverage2D : reference code, called as-is.
//   * the 3-line channel plugin front-end (plugins/channelrx/demodnfm/nfmdemod.cpp:152-155,315)
//     and the SpectrumVis glue loop (sdrgui/dsp/spectrumvis.cpp:98-233) are restated here because
//     those translation units pull in Qt audio / OpenGL and cannot be compiled without Qt.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load this.

#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <complex>
#include <list>
#include <random>
#include <functional>
#include <algorithm>

#include "dsp/dsptypes.h"
#include "dsp/decimators.h"
#include "dsp/decimatorsfi.h"
#include "dsp/decimatorsff.h"
#include "dsp/decimatorsif.h"
#include "dsp/dspcommands.h"
#include "dsp/fftwindow.h"
#include "dsp/fftengine.h"
#include "util/movingaverage2d.h"
#include "util/fixedaverage2d.h"
#include "util/messagequeue.h"

// test-only peek into private members (tap tables, NCO table, filter stage list)
#define private public
#define protected public
#include "dsp/nco.h"
#include "dsp/interpolator.h"
#include "dsp/downchannelizer.h"
#undef private
#undef protected

// "fake moc": the two Qt signals the reference declares (no-ops without an event loop)
void MessageQueue::messageEnqueued() {}
void DownChannelizer::inputSampleRateChanged() {}

namespace {

enum { MODE_INF = 0, MODE_SUP = 1, MODE_CEN = 2 };

template<typename D, typename It, typename T>
bool dispatch(D& d, int log2, int mode, It* it, const T* buf, int len)
{
    if (log2 == 0) { d.decimate1(it, buf, len); return true; }
#define ORACLE_CASE(N, L) \
    case L: \
        if (mode == MODE_INF) d.decimate##N##_inf(it, buf, len); \
        else if (mode == MODE_SUP) d.decimate##N##_sup(it, buf, len); \
        else d.decimate##N##_cen(it, buf, len); \
        return true;
    switch (log2) {
        ORACLE_CASE(2, 1)
        ORACLE_CASE(4, 2)
        ORACLE_CASE(8, 3)
        ORACLE_CASE(16, 4)
        ORACLE_CASE(32, 5)
        ORACLE_CASE(64, 6)
    default: return false;
    }
#undef ORACLE_CASE
}

struct DecimII {
    int bits;
    Decimators<qint32, qint16, SDR_RX_SAMP_SZ, 8>  d8;
    Decimators<qint32, qint16, SDR_RX_SAMP_SZ, 12> d12;
    Decimators<qint32, qint16, SDR_RX_SAMP_SZ, 16> d16;
    SampleVector out;
};

// 8-bit signed device format: HackRF (hackrfinputthread.h:57); the unsigned one (DecimatorsU) lives in ref_capi_u8.cpp
struct DecimI8 { Decimators<qint32, qint8, SDR_RX_SAMP_SZ, 8> d; SampleVector out; };
struct DecimFI { DecimatorsFI d; SampleVector out; };
struct DecimFF { DecimatorsFF d; FSampleVector out; };
struct DecimIF {
    int bits;
    DecimatorsIF<qint16, 8>  d8;
    DecimatorsIF<qint16, 12> d12;
    DecimatorsIF<qint16, 16> d16;
    FSampleVector out;
};

// Sink that captures what DownChannelizer hands to the demodulator
class CaptureSink : public BasebandSampleSink {
public:
    SampleVector captured;
    virtual void start() {}
    virtual void stop() {}
    virtual void feed(const SampleVector::const_iterator& begin, const SampleVector::const_iterator& end, bool)
    {
        captured.insert(captured.end(), begin, end);
    }
    virtual bool handleMessage(const Message&) { return false; }
};

struct Chan {
    CaptureSink sink;
    DownChannelizer* chan;
    SampleVector in;
    Chan() : chan(new DownChannelizer(&sink)) {}
    ~Chan() { delete chan; }
};

// plugin front-end as in nfmdemod.cpp:150-155,315,462-470
struct FrontEnd {
    NCO nco;
    Interpolator interp;
    Real distance;
    Real distanceRemain;
};

struct Spectrum {
    FFTEngine* fft;
    FFTWindow window;
    std::vector<Complex> fftBuffer;
    std::vector<Real> powerSpectrum;
    MovingAverage2D<double> movingAverage;
    FixedAverage2D<double> fixedAverage;
    std::size_t fftSize, overlapPercent, overlapSize, refillSize, fftBufferFill;
    unsigned int averageNb;
    int averagingMode; // 0 none 1 moving 2 fixed (spectrumvis.h AveragingMode)
    bool linear;
    Real scalef, ofs, powFFTDiv, mult;
};

} // namespace

extern "C" {

// ---------------------------------------------------------------- Decimators (int -> int)
void* ref_decim_ii_create(int input_bits)
{
    if (input_bits != 8 && input_bits != 12 && input_bits != 16) return 0;
    DecimII* h = new DecimII;
    h->bits = input_bits;
    return h;
}
void ref_decim_ii_destroy(void* p) { delete (DecimII*) p; }

// returns number of Samples written to out (int16 I,Q interleaved), or -1
int ref_decim_ii_run(void* p, int log2, int mode, const int16_t* buf, int len, int16_t* out)
{
    DecimII* h = (DecimII*) p;
    std::size_t need = (std::size_t) (len / 2) + 8;
    if (h->out.size() < need) h->out.resize(need);
    SampleVector::iterator it = h->out.begin();
    bool ok;
    if (h->bits == 8) ok = dispatch(h->d8, log2, mode, &it, buf, len);
    else if (h->bits == 12) ok = dispatch(h->d12, log2, mode, &it, buf, len);
    else ok = dispatch(h->d16, log2, mode, &it, buf, len);
    if (!ok) return -1;
    int n = (int) (it - h->out.begin());
    if (n > 0) memcpy(out, &h->out[0], (std::size_t) n * sizeof(Sample));
    return n;
}

// the split-I/Q overloads (decimators.h:359-371,395-417,2638-2700,...,3888-...) and decimate2_u (:374-393): int16 in, 12-bit
// log2 = 0 decimate1, 1..6 decimateN_cen on separate I and Q arrays of `len` samples each; mode_u = 1: decimate2_u (log2 must be 1)
int ref_decim_ii_run_split(void* p, int log2, int mode_u, const int16_t* bufI, const int16_t* bufQ, int len, int16_t* out)
{
    DecimII* h = (DecimII*) p;
    std::size_t need = (std::size_t) len + 8;
    if (h->out.size() < need) h->out.resize(need);
    SampleVector::iterator it = h->out.begin();
    if (h->bits != 12) return -1;
    if (mode_u) { if (log2 != 1) return -1; h->d12.decimate2_u(&it, bufI, bufQ, len); }
    else switch (log2) {
        case 0: h->d12.decimate1(&it, bufI, bufQ, len); break;
        case 1: h->d12.decimate2_cen(&it, bufI, bufQ, len); break;
        case 2: h->d12.decimate4_cen(&it, bufI, bufQ, len); break;
        case 3: h->d12.decimate8_cen(&it, bufI, bufQ, len); break;
        case 4: h->d12.decimate16_cen(&it, bufI, bufQ, len); break;
        case 5: h->d12.decimate32_cen(&it, bufI, bufQ, len); break;
        case 6: h->d12.decimate64_cen(&it, bufI, bufQ, len); break;
        default: return -1;
    }
    int n = (int) (it - h->out.begin());
    if (n > 0) memcpy(out, &h->out[0], (std::size_t) n * sizeof(Sample));
    return n;
}
int ref_decim_ii_run_2u(void* p, const int16_t* buf, int len, int16_t* out)
{
    DecimII* h = (DecimII*) p;
    std::size_t need = (std::size_t) (len / 2) + 8;
    if (h->out.size() < need) h->out.resize(need);
    SampleVector::iterator it = h->out.begin();
    if (h->bits == 8) h->d8.decimate2_u(&it, buf, len);
    else if (h->bits == 12) h->d12.decimate2_u(&it, buf, len);
    else h->d16.decimate2_u(&it, buf, len);
    int n = (int) (it - h->out.begin());
    if (n > 0) memcpy(out, &h->out[0], (std::size_t) n * sizeof(Sample));
    return n;
}

// ---------------------------------------------------------------- 8-bit inputs (int8 / uint8-127 -> int16)
void* ref_decim_i8_create() { return new DecimI8; }
void ref_decim_i8_destroy(void* p) { delete (DecimI8*) p; }
int ref_decim_i8_run(void* p, int log2, int mode, const int8_t* buf, int len, int16_t* out)
{
    DecimI8* h = (DecimI8*) p;
    std::size_t need = (std::size_t) (len / 2) + 8;
    if (h->out.size() < need) h->out.resize(need);
    SampleVector::iterator it = h->out.begin();
    if (!dispatch(h->d, log2, mode, &it, (const qint8*) buf, len)) return -1;
    int n = (int) (it - h->out.begin());
    if (n > 0) memcpy(out, &h->out[0], (std::size_t) n * sizeof(Sample));
    return n;
}
// ---------------------------------------------------------------- DecimatorsFI (float -> int16)
void* ref_decim_fi_create() { return new DecimFI; }
void ref_decim_fi_destroy(void* p) { delete (DecimFI*) p; }
int ref_decim_fi_run(void* p, int log2, int mode, const float* buf, int len, int16_t* out)
{
    DecimFI* h = (DecimFI*) p;
    std::size_t need = (std::size_t) (len / 2) + 8;
    if (h->out.size() < need) h->out.resize(need);
    SampleVector::iterator it = h->out.begin();
    if (!dispatch(h->d, log2, mode, &it, buf, len)) return -1;
    int n = (int) (it - h->out.begin());
    if (n > 0) memcpy(out, &h->out[0], (std::size_t) n * sizeof(Sample));
    return n;
}

// ---------------------------------------------------------------- DecimatorsFF (float -> float)
void* ref_decim_ff_create() { return new DecimFF; }
void ref_decim_ff_destroy(void* p) { delete (DecimFF*) p; }
int ref_decim_ff_run(void* p, int log2, int mode, const float* buf, int len, float* out)
{
    DecimFF* h = (DecimFF*) p;
    std::size_t need = (std::size_t) (len / 2) + 8;
    if (h->out.size() < need) h->out.resize(need);
    FSampleVector::iterator it = h->out.begin();
    if (!dispatch(h->d, log2, mode, &it, buf, len)) return -1;
    int n = (int) (it - h->out.begin());
    if (n > 0) memcpy(out, &h->out[0], (std::size_t) n * sizeof(FSample));
    return n;
}

// ---------------------------------------------------------------- DecimatorsIF (int16 -> float)
void* ref_decim_if_create(int input_bits)
{
    if (input_bits != 8 && input_bits != 12 && input_bits != 16) return 0;
    DecimIF* h = new DecimIF;
    h->bits = input_bits;
    return h;
}
void ref_decim_if_destroy(void* p) { delete (DecimIF*) p; }
int ref_decim_if_run(void* p, int log2, int mode, const int16_t* buf, int len, float* out)
{
    DecimIF* h = (DecimIF*) p;
    std::size_t need = (std::size_t) (len / 2) + 8;
    if (h->out.size() < need) h->out.resize(need);
    FSampleVector::iterator it = h->out.begin();
    bool ok;
    if (h->bits == 8) ok = dispatch(h->d8, log2, mode, &it, buf, len);
    else if (h->bits == 12) ok = dispatch(h->d12, log2, mode, &it, buf, len);
    else ok = dispatch(h->d16, log2, mode, &it, buf, len);
    if (!ok) return -1;
    int n = (int) (it - h->out.begin());
    if (n > 0) memcpy(out, &h->out[0], (std::size_t) n * sizeof(FSample));
    return n;
}

// ---------------------------------------------------------------- DownChannelizer
void* ref_chan_create() { return new Chan; }
void ref_chan_destroy(void* p) { delete (Chan*) p; }

// Drives the same two messages the engine and the plugin send (dspdevicesourceengine / nfmdemod.cpp:369-371).
// modes[i]: 0 centre, 1 lower half, 2 upper half (downchannelizer.h:72-76). Returns number of stages.
int ref_chan_configure(void* p, int input_rate, int requested_rate, int center_offset,
                       int* out_rate, int* residual_offset, int* modes, int modes_cap)
{
    Chan* h = (Chan*) p;
    DSPSignalNotification sig(input_rate, 0);
    h->chan->handleMessage(sig);
    DSPConfigureChannelizer cfg(requested_rate, center_offset);
    h->chan->handleMessage(cfg);
    // drain what the channelizer posted to the sink; the last MsgChannelizerNotification is current
    Message* m;
    int rate = 0, ofs = 0;
    while ((m = h->sink.getInputMessageQueue()->pop()) != 0) {
        if (DownChannelizer::MsgChannelizerNotification::match(*m)) {
            DownChannelizer::MsgChannelizerNotification* n = (DownChannelizer::MsgChannelizerNotification*) m;
            rate = n->getSampleRate();
            ofs = (int) n->getFrequencyOffset();
        }
        delete m;
    }
    if (out_rate) *out_rate = rate;
    if (residual_offset) *residual_offset = ofs;
    int i = 0;
    for (DownChannelizer::FilterStages::iterator it = h->chan->m_filterStages.begin(); it != h->chan->m_filterStages.end(); ++it, ++i) {
        if (modes && i < modes_cap) modes[i] = (int) (*it)->m_mode;
    }
    return i;
}

// returns samples produced (written to out up to cap), or -1 if cap too small
int ref_chan_feed(void* p, const int16_t* iq, int n, int16_t* out, int cap)
{
    Chan* h = (Chan*) p;
    h->in.resize((std::size_t) n);
    if (n > 0) memcpy(&h->in[0], iq, (std::size_t) n * sizeof(Sample));
    h->sink.captured.clear();
    h->chan->feed(h->in.begin(), h->in.end(), false);
    int m = (int) h->sink.captured.size();
    if (m > cap) return -1;
    if (m > 0) memcpy(out, &h->sink.captured[0], (std::size_t) m * sizeof(Sample));
    return m;
}

// ---------------------------------------------------------------- NCO + Interpolator (plugin front-end)
void* ref_frontend_create(float nco_freq, float nco_rate, int phase_steps, double interp_rate, double cutoff,
                          double taps_per_phase, float distance)
{
    FrontEnd* f = new FrontEnd;
    f->nco.setFreq(nco_freq, nco_rate);
    f->interp.create(phase_steps, interp_rate, cutoff, taps_per_phase);
    f->distance = distance;
    f->distanceRemain = 0;
    return f;
}
void ref_frontend_destroy(void* p) { delete (FrontEnd*) p; }
int ref_frontend_nco_increment(void* p) { return ((FrontEnd*) p)->nco.m_phaseIncrement; }
int ref_frontend_ntaps(void* p) { return ((FrontEnd*) p)->interp.m_nTaps; }
// taps[phase * ntaps + i]
void ref_frontend_taps(void* p, float* taps)
{
    FrontEnd* f = (FrontEnd*) p;
    int n = f->interp.m_nTaps * f->interp.m_phaseSteps;
    for (int i = 0; i < n; i++) taps[i] = f->interp.m_alignedTaps[2 * i];
}
void ref_nco_table(float* table4096)
{
    NCO n; (void) n;
    memcpy(table4096, NCO::m_table, sizeof(Real) * 4096);
}
// feed int16 IQ (what DownChannelizer hands over); out = complex64 pairs; idx/phase = schedule (may be null)
int ref_frontend_feed(void* p, const int16_t* iq, int n, float* out, int cap, int32_t* idx, int32_t* phase)
{
    FrontEnd* f = (FrontEnd*) p;
    int m = 0;
    Complex ci;
    for (int i = 0; i < n; i++) {
        Complex c(iq[2 * i], iq[2 * i + 1]);
        c *= f->nco.nextIQ();
        Real before = f->distanceRemain;
        if (f->interp.decimate(&f->distanceRemain, c, &ci)) {
            if (m >= cap) return -1;
            out[2 * m] = ci.real();
            out[2 * m + 1] = ci.imag();
            if (idx) idx[m] = i;
            if (phase) {
                int ph = (int) floor(f->distanceRemain * (Real) f->interp.m_phaseSteps);
                phase[m] = ph < 0 ? 0 : ph;
            }
            m++;
            f->distanceRemain += f->distance;
        }
        (void) before;
    }
    return m;
}
// complex float input variant (Interpolator::decimate only, no NCO) for unit tests
int ref_interp_decimate(void* p, const float* cin, int n, float* out, int cap)
{
    FrontEnd* f = (FrontEnd*) p;
    int m = 0;
    Complex ci;
    for (int i = 0; i < n; i++) {
        Complex c(cin[2 * i], cin[2 * i + 1]);
        if (f->interp.decimate(&f->distanceRemain, c, &ci)) {
            if (m >= cap) return -1;
            out[2 * m] = ci.real();
            out[2 * m + 1] = ci.imag();
            m++;
            f->distanceRemain += f->distance;
        }
    }
    return m;
}

// Interpolator::interpolate in the loop every Tx plugin writes around it (plugins/channeltx/modnfm/nfmmod.cpp:126-133,
// modam/ammod.cpp:120-127, modssb/ssbmod.cpp:146-153): one call per OUTPUT sample, a new input is fetched when the call
// consumed the current one, then distanceRemain += distance.  Stops when the next call would need input n.
int ref_interp_interpolate(void* p, const float* cin, int n, float* out, int cap)
{
    FrontEnd* f = (FrontEnd*) p;
    int m = 0, i = 0;
    Complex ci;
    while (i < n || f->distanceRemain < 1.0f) {
        if (f->distanceRemain >= 1.0f && i >= n) break;
        Complex c = (i < n) ? Complex(cin[2 * i], cin[2 * i + 1]) : Complex(0, 0);     // (not consumed when i == n: distanceRemain < 1)
        if (f->interp.interpolate(&f->distanceRemain, c, &ci)) i++;
        if (m >= cap) return -1;
        out[2 * m] = ci.real();
        out[2 * m + 1] = ci.imag();
        m++;
        f->distanceRemain += f->distance;
    }
    return m;
}
// Interpolator::resample (the arbitrary P/Q form, interpolator.h:55-76) in its canonical loop: per input, call until consumed;
// every call that returns true yields an output and distanceRemain += distance.
int ref_interp_resample(void* p, const float* cin, int n, float* out, int cap)
{
    FrontEnd* f = (FrontEnd*) p;
    int m = 0;
    Complex ci;
    for (int i = 0; i < n; i++) {
        Complex c(cin[2 * i], cin[2 * i + 1]);
        bool consumed = false;
        do {
            if (f->interp.resample(&f->distanceRemain, c, &consumed, &ci)) {
                if (m >= cap) return -1;
                out[2 * m] = ci.real();
                out[2 * m + 1] = ci.imag();
                m++;
                f->distanceRemain += f->distance;
            }
        } while (!consumed);
    }
    return m;
}
float ref_frontend_remain(void* p) { return ((FrontEnd*) p)->distanceRemain; }
// NCO::nextIQ n times (nco.h:40-53, nco.cpp:48-64): (re, im) pairs
void ref_nco_block(float freq, float rate, int n, float* out)
{
    NCO nco;
    nco.setFreq(freq, rate);
    for (int i = 0; i < n; i++) { Complex c = nco.nextIQ(); out[2 * i] = c.real(); out[2 * i + 1] = c.imag(); }
}

// ---------------------------------------------------------------- SpectrumVis (glue restated, arithmetic is reference code)
void ref_spectrum_configure(void* p, int fftSize, int overlapPercent, unsigned int averageNb, int averagingMode, int window, int linear);

void* ref_spectrum_create(float scalef)
{
    Spectrum* s = new Spectrum;
    s->fft = FFTEngine::create();
    s->fftBuffer.resize(4096);
    s->powerSpectrum.resize(4096);
    s->scalef = scalef;
    s->mult = (10.0f / log2f(10.0f)); // spectrumvis.cpp:17
    ref_spectrum_configure(s, 1024, 0, 0, 0, (int) FFTWindow::BlackmanHarris, 0);
    return s;
}
void ref_spectrum_destroy(void* p)
{
    Spectrum* s = (Spectrum*) p;
    delete s->fft;
    delete s;
}
// spectrumvis.cpp:283-327
void ref_spectrum_configure(void* p, int fftSize, int overlapPercent, unsigned int averageNb, int averagingMode, int window, int linear)
{
    Spectrum* s = (Spectrum*) p;
    if (fftSize > 4096) fftSize = 4096; else if (fftSize < 64) fftSize = 64;
    if (overlapPercent > 100) overlapPercent = 100; else if (overlapPercent < 0) overlapPercent = 0;
    s->overlapPercent = overlapPercent;
    s->fftSize = fftSize;
    s->fft->configure(fftSize, false);
    s->window.create((FFTWindow::Function) window, fftSize);
    s->overlapSize = (s->fftSize * s->overlapPercent) / 100;
    s->refillSize = s->fftSize - s->overlapSize;
    s->fftBufferFill = s->overlapSize;
    s->movingAverage.resize(fftSize, averageNb);
    s->fixedAverage.resize(fftSize, averageNb);
    s->averageNb = averageNb;
    s->averagingMode = averagingMode;
    s->linear = linear != 0;
    s->ofs = 20.0f * log10f(1.0f / s->fftSize);
    s->powFFTDiv = s->fftSize * s->fftSize;
}
void ref_spectrum_window(void* p, float* w)
{
    Spectrum* s = (Spectrum*) p;
    std::vector<Complex> ones(s->fftSize, Complex(1.0f, 0.0f)), out(s->fftSize);
    s->window.apply(&ones[0], &out[0]);
    for (std::size_t i = 0; i < s->fftSize; i++) w[i] = out[i].real();
}
// raw FFT of the reference engine (for FFT unit tests): in/out complex64 pairs of length fftSize
void ref_spectrum_fft(void* p, const float* in, float* out)
{
    Spectrum* s = (Spectrum*) p;
    memcpy(s->fft->in(), in, s->fftSize * sizeof(Complex));
    s->fft->transform();
    memcpy(out, s->fft->out(), s->fftSize * sizeof(Complex));
}

// returns number of frames handed to GLSpectrum::newSpectrum; each frame = fftSize floats. spectrumvis.cpp:77-254
int ref_spectrum_feed(void* p, const int16_t* iq, int n, int positiveOnly, float* frames, int cap_frames)
{
    Spectrum* s = (Spectrum*) p;
    const Sample* begin = (const Sample*) iq;
    const Sample* end = begin + n;
    int nframes = 0;
    const Real m_scalef = s->scalef, m_powFFTDiv = s->powFFTDiv, m_mult = s->mult, m_ofs = s->ofs;
    const bool m_linear = s->linear;
    std::vector<Real>& m_powerSpectrum = s->powerSpectrum;

    while (begin < end)
    {
        std::size_t todo = end - begin;
        std::size_t samplesNeeded = s->refillSize - s->fftBufferFill;

        if (todo >= samplesNeeded)
        {
            std::vector<Complex>::iterator it = s->fftBuffer.begin() + s->fftBufferFill;
            for (std::size_t i = 0; i < samplesNeeded; ++i, ++begin) {
                *it++ = Complex(begin->real() / m_scalef, begin->imag() / m_scalef);
            }
            s->window.apply(&s->fftBuffer[0], s->fft->in());
            s->fft->transform();
            const Complex* fftOut = s->fft->out();
            Complex c;
            Real v;
            std::size_t halfSize = s->fftSize / 2;
            bool emitFrame = false;

            if (s->averagingMode == 0)
            {
                if (positiveOnly) {
                    for (std::size_t i = 0; i < halfSize; i++) {
                        c = fftOut[i];
                        v = c.real() * c.real() + c.imag() * c.imag();
                        v = m_linear ? v / m_powFFTDiv : m_mult * log2f(v) + m_ofs;
                        m_powerSpectrum[i * 2] = v;
                        m_powerSpectrum[i * 2 + 1] = v;
                    }
                } else {
                    for (std::size_t i = 0; i < halfSize; i++) {
                        c = fftOut[i + halfSize];
                        v = c.real() * c.real() + c.imag() * c.imag();
                        v = m_linear ? v / m_powFFTDiv : m_mult * log2f(v) + m_ofs;
                        m_powerSpectrum[i] = v;
                        c = fftOut[i];
                        v = c.real() * c.real() + c.imag() * c.imag();
                        v = m_linear ? v / m_powFFTDiv : m_mult * log2f(v) + m_ofs;
                        m_powerSpectrum[i + halfSize] = v;
                    }
                }
                emitFrame = true;
            }
            else if (s->averagingMode == 1)
            {
                if (positiveOnly) {
                    for (std::size_t i = 0; i < halfSize; i++) {
                        c = fftOut[i];
                        v = c.real() * c.real() + c.imag() * c.imag();
                        v = s->movingAverage.storeAndGetAvg(v, i);
                        v = m_linear ? v / m_powFFTDiv : m_mult * log2f(v) + m_ofs;
                        m_powerSpectrum[i * 2] = v;
                        m_powerSpectrum[i * 2 + 1] = v;
                    }
                } else {
                    for (std::size_t i = 0; i < halfSize; i++) {
                        c = fftOut[i + halfSize];
                        v = c.real() * c.real() + c.imag() * c.imag();
                        v = s->movingAverage.storeAndGetAvg(v, i + halfSize);
                        v = m_linear ? v / m_powFFTDiv : m_mult * log2f(v) + m_ofs;
                        m_powerSpectrum[i] = v;
                        c = fftOut[i];
                        v = c.real() * c.real() + c.imag() * c.imag();
                        v = s->movingAverage.storeAndGetAvg(v, i);
                        v = m_linear ? v / m_powFFTDiv : m_mult * log2f(v) + m_ofs;
                        m_powerSpectrum[i + halfSize] = v;
                    }
                }
                emitFrame = true;
                s->movingAverage.nextAverage();
            }
            else
            {
                double avg;
                if (positiveOnly) {
                    for (std::size_t i = 0; i < halfSize; i++) {
                        c = fftOut[i];
                        v = c.real() * c.real() + c.imag() * c.imag();
                        if (s->fixedAverage.storeAndGetAvg(avg, v, i)) {
                            avg = m_linear ? v / m_powFFTDiv : m_mult * log2f(avg) + m_ofs;
                            m_powerSpectrum[i * 2] = avg;
                            m_powerSpectrum[i * 2 + 1] = avg;
                        }
                    }
                } else {
                    for (std::size_t i = 0; i < halfSize; i++) {
                        c = fftOut[i + halfSize];
                        v = c.real() * c.real() + c.imag() * c.imag();
                        if (s->fixedAverage.storeAndGetAvg(avg, v, i + halfSize)) {
                            avg = m_linear ? v / m_powFFTDiv : m_mult * log2f(avg) + m_ofs;
                            m_powerSpectrum[i] = avg;
                        }
                        c = fftOut[i];
                        v = c.real() * c.real() + c.imag() * c.imag();
                        if (s->fixedAverage.storeAndGetAvg(avg, v, i)) {
                            avg = m_linear ? v / m_powFFTDiv : m_mult * log2f(avg) + m_ofs;
                            m_powerSpectrum[i + halfSize] = avg;
                        }
                    }
                }
                if (s->fixedAverage.nextAverage()) emitFrame = true;
            }

            if (emitFrame) {
                if (nframes >= cap_frames) return -1;
                memcpy(frames + (std::size_t) nframes * s->fftSize, &m_powerSpectrum[0], s->fftSize * sizeof(Real));
                nframes++;
            }

            std::copy(s->fftBuffer.begin() + s->refillSize, s->fftBuffer.end(), s->fftBuffer.begin());
            s->fftBufferFill = s->overlapSize;
        }
        else
        {
            for (std::vector<Complex>::iterator it = s->fftBuffer.begin() + s->fftBufferFill; begin < end; ++begin) {
                *it++ = Complex(begin->real() / m_scalef, begin->imag() / m_scalef);
            }
            s->fftBufferFill += todo;
        }
    }
    return nframes;
}

// sdrbench input generators exactly as sdrbench/mainbench.cpp:30-31,76-79,122,149,176 build them
// (std::bind copies the default-seeded engine; 2N-1 draws; the never-written last element is set to 0)
void ref_sdrbench_gen_s16(int16_t* buf, int n_scalars)
{
    std::mt19937 generator;
    std::uniform_int_distribution<qint16> dist(-2048, 2047);
    auto my_rand = std::bind(dist, generator);
    std::generate(buf, buf + n_scalars - 1, my_rand);
    buf[n_scalars - 1] = 0;
}
void ref_sdrbench_gen_f32(float* buf, int n_scalars)
{
    std::mt19937 generator;
    std::uniform_real_distribution<float> dist(-1.0, 1.0);
    auto my_rand = std::bind(dist, generator);
    std::generate(buf, buf + n_scalars - 1, my_rand);
    buf[n_scalars - 1] = 0;
}

const char* ref_build_info()
{
    return "sdrangel reference compiled from /root/reference (USE_KISSFFT, 16-bit Rx), g++ "
#ifdef __VERSION__
        __VERSION__
#endif
#ifdef __FAST_MATH__
        " -ffast-math"
#endif
        ;
}

} // extern "C"
