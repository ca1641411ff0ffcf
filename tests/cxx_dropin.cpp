// Drop-in boundary test: code written against the reference's class names and method signatures, compiled against
// include/sdrangel_b200/dsp/*.h and linked to libb200dsp.so.  Prints FNV-1a-64 hashes the pytest wrapper compares with
// the golden values generated from the reference (SURVEY.md Appendix D convention).
#include <cstdio>
#include <cstdlib>
#include <random>
#include <functional>
#include <algorithm>
#include "sdrangel_b200/dsp/decimators.h"
#include "sdrangel_b200/dsp/decimatorsu.h"
#include "sdrangel_b200/dsp/decimatorsfi.h"
#include "sdrangel_b200/dsp/downchannelizer.h"
#include "sdrangel_b200/dsp/spectrumvis.h"
#include "sdrangel_b200/dsp/interpolator.h"
#include "sdrangel_b200/dsp/iqcorrections.h"

static uint64_t fnv(const void* p, size_t n_u16)
{
    const uint16_t* w = (const uint16_t*) p;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n_u16; i++) { h ^= w[i]; h *= 1099511628211ull; }
    return h;
}

struct CaptureSink : BasebandSampleSink {
    SampleVector captured;
    void start() {} void stop() {}
    void feed(const SampleVector::const_iterator& b, const SampleVector::const_iterator& e, bool) { captured.insert(captured.end(), b, e); }
};
struct CountDisplay : SpectrumDisplay { int frames = 0; double last0 = 0; void newSpectrum(const std::vector<Real>& s, int) { frames++; last0 = s[0]; } };

int main()
{
    try {
        if (b200dsp_init(0) != 0) { fprintf(stderr, "%s\n", b200dsp_last_error()); return 2; }
        // sdrbench decimateii exactly as MainBench::testDecimateII + decimateII do it (sdrbench/mainbench.cpp:69-104,193-230)
        const int nbSamples = 1 << 20;
        qint16* buf = new qint16[nbSamples * 2];
        std::mt19937 gen;
        std::uniform_int_distribution<qint16> dist(-2048, 2047);
        auto rnd = std::bind(dist, gen);
        std::generate(buf, buf + nbSamples * 2 - 1, rnd);
        buf[nbSamples * 2 - 1] = 0;
        Decimators<qint32, qint16, SDR_RX_SAMP_SZ, 12> m_decimatorsII;
        SampleVector m_convertBuffer(nbSamples);
        SampleVector::iterator it = m_convertBuffer.begin();
        m_decimatorsII.decimate16_cen(&it, buf, nbSamples * 2);
        const size_t n_out = it - m_convertBuffer.begin();
        printf("decimate16_cen n_out=%zu in=%016llx out=%016llx\n", n_out, (unsigned long long) fnv(buf, 2 * (size_t) nbSamples),
               (unsigned long long) fnv(&m_convertBuffer[0], 2 * n_out));
        // 8-bit device plugins (rtlsdrthread.h:55 / rtlsdrthread.cpp:167, hackrfinputthread.h:57 / hackrfinputthread.cpp:155)
        // on the low bytes of the same buffer
        {
            std::vector<quint8> u8(nbSamples * 2);
            for (int i = 0; i < nbSamples * 2; i++) u8[i] = (quint8) (buf[i] & 0xff);
            DecimatorsU<qint32, quint8, SDR_RX_SAMP_SZ, 8, 127> m_decimatorsU;
            it = m_convertBuffer.begin();
            m_decimatorsU.decimate16_cen(&it, &u8[0], nbSamples * 2);
            size_t nu = it - m_convertBuffer.begin();
            printf("decimatorsu16_cen n_out=%zu out=%016llx\n", nu, (unsigned long long) fnv(&m_convertBuffer[0], 2 * nu));
            Decimators<qint32, qint8, SDR_RX_SAMP_SZ, 8> m_decimators8;
            it = m_convertBuffer.begin();
            m_decimators8.decimate64_inf(&it, (const qint8*) &u8[0], nbSamples * 2);
            nu = it - m_convertBuffer.begin();
            printf("decimators8_64_inf n_out=%zu out=%016llx\n", nu, (unsigned long long) fnv(&m_convertBuffer[0], 2 * nu));
        }
        // engine-side DC correction on the decimated vector, in two parts like DSPDeviceSourceEngine::work's FIFO halves (:343-379)
        {
            IQCorrections corr;
            SampleVector v(1 << 16);
            for (size_t i = 0; i < v.size(); i++) { v[i].setReal(buf[2 * i]); v[i].setImag(buf[2 * i + 1]); }
            corr.iqCorrections(v.begin(), v.begin() + 40000, false);
            corr.iqCorrections(v.begin() + 40000, v.end(), false);
            printf("iqcorrections out=%016llx\n", (unsigned long long) fnv(&v[0], 2 * v.size()));
            bool threw = false;
            try { corr.iqCorrections(v.begin(), v.begin() + 8, true); } catch (const std::runtime_error&) { threw = true; }
            printf("iqcorrections_imbalance %s\n", threw ? "throws" : "silent");
        }
        // DownChannelizer as a plugin wires it (nfmdemod.cpp:93-95): 10 MS/s, 48 kS/s at +1234567 Hz
        CaptureSink sink;
        DownChannelizer chan(&sink);
        chan.setInputSampleRate(10000000);
        chan.configure(48000, 1234567);
        SampleVector in(60000);
        std::mt19937 g2(12);
        for (auto& s : in) { s.setReal((qint16) (g2() >> 16)); s.setImag((qint16) (g2() >> 16)); }
        chan.feed(in.begin(), in.begin() + 1000, false);
        chan.feed(in.begin() + 1000, in.end(), false);
        printf("downchannelizer rate=%d ofs=%d n_out=%zu\n", chan.getCurrentOutputSampleRate(), chan.getCurrentCenterFrequency(), sink.captured.size());
        CountDisplay disp;
        SpectrumVis vis(SDR_RX_SCALEF, &disp);
        vis.configure(4096, 0, 10, SpectrumVis::AvgModeFixed, 1, false);
        SampleVector sv(4096 * 25);
        for (auto& s : sv) { s.setReal((qint16) ((g2() >> 20) - 2048)); s.setImag((qint16) ((g2() >> 20) - 2048)); }
        vis.feed(sv.begin(), sv.end(), false);
        printf("spectrumvis frames=%d\n", disp.frames);
        Interpolator interp;
        interp.create(16, 156250, 12500 / 2.2f);
        std::vector<Complex> cin(20000), cout(20000);
        for (size_t i = 0; i < cin.size(); i++) cin[i] = Complex((Real) ((int) (g2() >> 18) - 8192), (Real) ((int) (g2() >> 18) - 8192));
        Real distance = 0, step = (Real) 156250 / (Real) 48000;
        size_t m = interp.decimate(&distance, step, &cin[0], cin.size(), &cout[0], cout.size());
        printf("interpolator n_out=%zu\n", m);
        delete[] buf;
    } catch (const std::exception& e) {
        fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
    return 0;
}
