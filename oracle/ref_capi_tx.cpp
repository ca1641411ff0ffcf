// ORACLE (test infrastructure, never shipped, never on the product path).
//
// C-ABI driver around the UNMODIFIED reference Tx-side classes, compiled in place from /root/reference by
// oracle/Makefile into oracle/_ref/libsdrref*.so next to ref_capi.cpp:
//   Interpolators<T,SdrBits,OutputBits>::interpolate{1,2,...,64}_cen     sdrbase/dsp/interpolators.h:104-617
//   IntHalfbandFilterEO1<order>::myInterpolate / workInterpolate*         sdrbase/dsp/inthalfbandfiltereo1.h:98-127,291-355,490-554,601-622
//   UpChannelizer::pull / applyConfiguration / createFilterChain          sdrbase/dsp/upchannelizer.cpp:51-104,175-209,252-327
// Nothing of the reference is copied: this file #includes the reference headers and calls them.
#include <cstdint>
#include <cstring>
#include <vector>

#include "dsp/dsptypes.h"
#include "dsp/interpolators.h"
#include "dsp/dspcommands.h"
#include "util/messagequeue.h"

#define private public
#define protected public
#include "dsp/upchannelizer.h"
#undef private
#undef protected

// "fake moc" for the Qt signals these classes declare (no event loop here)
void UpChannelizer::outputSampleRateChanged() {}
void SampleSourceFifo::dataWrite(int) {}
void SampleSourceFifo::dataRead(int) {}

namespace {

struct Interps {
    int bits;
    Interpolators<qint16, 16, 16> d16;
    Interpolators<qint16, 16, 12> d12;
    Interpolators<qint8, 16, 8> d8;
    SampleVector in;
};

template<typename D, typename T>
bool interp_dispatch(D& d, int log2, SampleVector::iterator* it, T* buf, int len)
{
    switch (log2) {
    case 0: d.interpolate1(it, buf, len); return true;
    case 1: d.interpolate2_cen(it, buf, len); return true;
    case 2: d.interpolate4_cen(it, buf, len); return true;
    case 3: d.interpolate8_cen(it, buf, len); return true;
    case 4: d.interpolate16_cen(it, buf, len); return true;
    case 5: d.interpolate32_cen(it, buf, len); return true;
    case 6: d.interpolate64_cen(it, buf, len); return true;
    }
    return false;
}

// the modulator behind an UpChannelizer: hands out the samples of one block in order, zeros when it runs dry
class VectorSource : public BasebandSampleSource {
public:
    std::vector<Sample> data;
    std::size_t pos = 0;
    virtual void start() {}
    virtual void stop() {}
    virtual void pull(Sample& s) { if (pos < data.size()) s = data[pos]; else s = Sample(); ++pos; }
    virtual bool handleMessage(const Message&) { return false; }
};

struct UpChan {
    VectorSource src;
    UpChannelizer* chan;
    UpChan() { chan = new UpChannelizer(&src); }
    ~UpChan() { delete chan; }
};

} // namespace

extern "C" {

void* ref_interps_create(int output_bits)
{
    if (output_bits != 8 && output_bits != 12 && output_bits != 16) return 0;
    Interps* h = new Interps;
    h->bits = output_bits;
    return h;
}
void ref_interps_destroy(void* p) { delete (Interps*) p; }

// == interpolateN_cen(&it, buf, len): `len` counts output scalars; consumes len / (2 << log2) samples from `iq`
// (n_samples must cover that).  buf is int16 (bits 12/16) or int8 (bits 8).  Returns samples consumed.
int ref_interps_run(void* p, int log2, const int16_t* iq, int n_samples, void* buf, int len)
{
    Interps* h = (Interps*) p;
    h->in.resize((std::size_t) n_samples + 1);
    if (n_samples > 0) memcpy(&h->in[0], iq, (std::size_t) n_samples * sizeof(Sample));
    SampleVector::iterator it = h->in.begin();
    bool ok;
    if (h->bits == 16) ok = interp_dispatch(h->d16, log2, &it, (qint16*) buf, len);
    else if (h->bits == 12) ok = interp_dispatch(h->d12, log2, &it, (qint16*) buf, len);
    else ok = interp_dispatch(h->d8, log2, &it, (qint8*) buf, len);
    if (!ok) return -1;
    return (int) (it - h->in.begin());
}

// integer coefficients of the interpolating half-bands (hbfiltertraits.cpp); returns order / 4 and the shift
int ref_hb_coeffs(int order, int32_t* out, int* shift)
{
    switch (order) {
    case 16: for (int i = 0; i < 4; i++) out[i] = HBFIRFilterTraits<16>::hbCoeffs[i]; *shift = HBFIRFilterTraits<16>::hbShift; return 4;
    case 32: for (int i = 0; i < 8; i++) out[i] = HBFIRFilterTraits<32>::hbCoeffs[i]; *shift = HBFIRFilterTraits<32>::hbShift; return 8;
    case 64: for (int i = 0; i < 16; i++) out[i] = HBFIRFilterTraits<64>::hbCoeffs[i]; *shift = HBFIRFilterTraits<64>::hbShift; return 16;
    case 96: for (int i = 0; i < 24; i++) out[i] = HBFIRFilterTraits<96>::hbCoeffs[i]; *shift = HBFIRFilterTraits<96>::hbShift; return 24;
    }
    return -1;
}

void* ref_upchan_create() { return new UpChan; }
void ref_upchan_destroy(void* p) { delete (UpChan*) p; }

// the two messages the sink engine and the modulator plugin send (DSPSignalNotification with the device rate,
// DSPConfigureChannelizer with the modulator's rate and offset); modes[i]: 0 centre, 1 lower half, 2 upper half,
// in the reference's stage order (stage 0 runs at the output rate).  Returns the number of stages.
int ref_upchan_configure(void* p, int output_rate, int requested_rate, int center_offset,
                         int* in_rate, int* residual_offset, int* modes, int modes_cap)
{
    UpChan* h = (UpChan*) p;
    DSPSignalNotification sig(output_rate, 0);
    h->chan->handleMessage(sig);
    DSPConfigureChannelizer cfg(requested_rate, center_offset);
    h->chan->handleMessage(cfg);
    Message* m;
    int rate = 0, ofs = 0;
    while ((m = h->src.getInputMessageQueue()->pop()) != 0) {
        if (UpChannelizer::MsgChannelizerNotification::match(*m)) {
            UpChannelizer::MsgChannelizerNotification* n = (UpChannelizer::MsgChannelizerNotification*) m;
            rate = n->getSampleRate();
            ofs = (int) n->getFrequencyOffset();
        }
        delete m;
    }
    if (in_rate) *in_rate = rate;
    if (residual_offset) *residual_offset = ofs;
    typedef IntHalfbandFilterEO1<UPCHANNELIZER_HB_FILTER_ORDER> F;
    int i = 0;
    for (UpChannelizer::FilterStages::iterator it = h->chan->m_filterStages.begin(); it != h->chan->m_filterStages.end(); ++it, ++i) {
        if (modes && i < modes_cap)
            modes[i] = (*it)->m_workFunction == &F::workInterpolateCenter ? 0 : (*it)->m_workFunction == &F::workInterpolateLowerHalf ? 1 : 2;
    }
    return i;
}

// n_out calls of UpChannelizer::pull; the modulator hands out `iq` (n_in samples, then zeros).  Returns how many
// modulator samples were pulled.
int ref_upchan_pull(void* p, const int16_t* iq, int n_in, int16_t* out, int n_out)
{
    UpChan* h = (UpChan*) p;
    h->src.data.resize((std::size_t) n_in);
    if (n_in > 0) memcpy(&h->src.data[0], iq, (std::size_t) n_in * sizeof(Sample));
    h->src.pos = 0;
    Sample s;
    for (int i = 0; i < n_out; i++) {
        h->chan->pull(s);
        out[2 * i] = s.real();
        out[2 * i + 1] = s.imag();
    }
    return (int) h->src.pos;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------
// demodulator back-ends after Interpolator::decimate, and the .sdriq record format
//   PhaseDiscriminators (sdrbase/dsp/phasediscri.h:26-198): reference code, header-only, called as-is
//   AMDemod::processOneSample's magnitude (plugins/channelrx/demodam/amdemod.cpp:154-156,241): three lines restated here
//     (amdemod.cpp pulls in Qt audio and cannot be compiled without Qt)
//   FileRecord (sdrbase/dsp/filerecord.cpp:72-148): reference code compiled in place
// ---------------------------------------------------------------------------------------------------------
#include <complex>
#include "dsp/phasediscri.h"
#include "dsp/filerecord.h"

extern "C" {

void* ref_discri_create(float fm_scaling)
{
    PhaseDiscriminators* d = new PhaseDiscriminators;
    d->reset();
    d->setFMScaling(fm_scaling);
    // m_prevArg and the phaseDiscriminator3 members are left uninitialised by the reference (phasediscri.h:31-35,133-140); a
    // freshly constructed demodulator object is zero-filled memory in practice: start from zero
    double ms; Real dev;
    Complex z(1.0f, 0.0f);
    d->phaseDiscriminatorDelta(z, ms, dev);       // atan2(0, 1) = 0 -> m_prevArg = 0
    d->reset();
    return d;
}
void ref_discri_destroy(void* p) { delete (PhaseDiscriminators*) p; }

// kind 0: phaseDiscriminator, 1: phaseDiscriminatorDelta (aux0 = magsq as float, aux1 = fmDev), 2: phaseDiscriminator2,
// 3: AM magnitude (out = sqrt(magsq), aux0 = magsq)
void ref_discri_run(void* p, int kind, const float* in_c64, int n, float* out, float* aux0, float* aux1)
{
    PhaseDiscriminators* d = (PhaseDiscriminators*) p;
    for (int i = 0; i < n; i++) {
        Complex ci(in_c64[2 * i], in_c64[2 * i + 1]);
        if (kind == 0) out[i] = d->phaseDiscriminator(ci);
        else if (kind == 1) {
            double magsq; Real dev;
            out[i] = d->phaseDiscriminatorDelta(ci, magsq, dev);
            if (aux0) aux0[i] = (float) magsq;
            if (aux1) aux1[i] = dev;
        } else if (kind == 2) out[i] = d->phaseDiscriminator2(ci);
        else {
            Real re = ci.real() / SDR_RX_SCALEF;          // amdemod.cpp:154-156
            Real im = ci.imag() / SDR_RX_SCALEF;
            Real magsq = re*re + im*im;
            if (aux0) aux0[i] = magsq;
            out[i] = sqrt(magsq);                         // amdemod.cpp:241 (the delay line between them only delays)
        }
    }
}

// FileRecord driven the way the engine drives it: DSPSignalNotification, startRecording, feed ..., stopRecording
int ref_filerecord_write(const char* path, int sample_rate, long long center_frequency, const int16_t* iq, int n1, int n2)
{
    FileRecord rec((QString(path)));
    DSPSignalNotification sig(sample_rate, center_frequency);
    rec.handleMessage(sig);
    rec.startRecording();
    SampleVector v((std::size_t) (n1 + n2));
    if (n1 + n2 > 0) memcpy(&v[0], iq, (std::size_t) (n1 + n2) * sizeof(Sample));
    rec.feed(v.begin(), v.begin() + n1, false);
    rec.feed(v.begin() + n1, v.end(), false);
    rec.stopRecording();
    return (int) rec.getByteCount();
}

int ref_filerecord_read_header(const char* path, int* sample_rate, unsigned long long* center_frequency, long long* ts, unsigned* sample_size)
{
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) return -1;
    FileRecord::Header h;
    FileRecord::readHeader(f, h);
    *sample_rate = h.sampleRate; *center_frequency = h.centerFrequency; *ts = (long long) h.startTimeStamp; *sample_size = h.sampleSize;
    return (int) f.tellg();
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------
// SURVEY.md 8f-4, SSB back-end: fftfilt (sdrbase/dsp/fftfilt.cpp:49-360, Fldigi's overlap-add FFT filter over g_fft, gfft.h),
// reference code compiled in place.  SSBDemod::feed calls runSSB / runDSB per Interpolator::decimate output
// (plugins/channelrx/demodssb/ssbdemod.cpp:91-92,165-175).
// ---------------------------------------------------------------------------------------------------------
#define protected public
#include "dsp/fftfilt.h"
#undef protected

extern "C" {

// kind 0: fftfilt(f1, f2, len) (band-pass, SSB), 1: fftfilt(f2, len) (DSB low-pass)
void* ref_fftfilt_create(int kind, float f1, float f2, int len)
{
    return kind == 0 ? new fftfilt(f1, f2, len) : new fftfilt(f2, len);
}
void ref_fftfilt_destroy(void* p) { delete (fftfilt*) p; }
void ref_fftfilt_set(void* p, int kind, float f1, float f2)
{
    if (kind == 0) ((fftfilt*) p)->create_filter(f1, f2); else ((fftfilt*) p)->create_dsb_filter(f2);
}
void ref_fftfilt_filter(void* p, float* out_c64) { fftfilt* f = (fftfilt*) p; memcpy(out_c64, f->filter, (std::size_t) f->flen * sizeof(fftfilt::cmplx)); }

// op 0: runFilt, 1: runSSB(usb = flag & 1, getDC = flag & 2), 2: runDSB(getDC = flag & 2); one call per input sample, outputs appended.
int ref_fftfilt_run(void* p, int op, int flag, const float* in_c64, int n, float* out_c64, int cap)
{
    fftfilt* f = (fftfilt*) p;
    int m = 0;
    for (int i = 0; i < n; i++) {
        fftfilt::cmplx* o = 0;
        const fftfilt::cmplx ci(in_c64[2 * i], in_c64[2 * i + 1]);
        int k = op == 0 ? f->runFilt(ci, &o) : op == 1 ? f->runSSB(ci, &o, (flag & 1) != 0, (flag & 2) != 0) : f->runDSB(ci, &o, (flag & 2) != 0);
        if (m + k > cap) return -1;
        if (k > 0) memcpy(out_c64 + 2 * m, o, (std::size_t) k * sizeof(fftfilt::cmplx));
        m += k;
    }
    return m;
}

} // extern "C"
