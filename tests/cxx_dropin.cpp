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
#include "sdrangel_b200/dsp/nco.h"
#include "sdrangel_b200/dsp/samplesinkfifo.h"
#include "sdrangel_b200/dsp/devicesamplesource.h"
#include "sdrangel_b200/dsp/upchannelizer.h"
#include "sdrangel_b200/dsp/interpolators.h"
#include "sdrangel_b200/dsp/phasediscri.h"
#include "sdrangel_b200/dsp/filerecord.h"
#include "sdrangel_b200/dsp/fftfilt.h"

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
// a modulator: hands out a deterministic sequence, one Sample per pull (what a channel Tx plugin's pull() does)
struct RampSource : BasebandSampleSource {
    int k = 0;
    void start() {} void stop() {}
    void pull(Sample& s) { s.setReal((qint16) (k * 37 - 20000)); s.setImag((qint16) (15000 - k * 91)); ++k; }
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
        // the reference's per-sample signatures (interpolator.h:23-52) give what the block forms give, sample for sample
        {
            Interpolator a, b;
            a.create(16, 156250, 12500 / 2.2f);
            b.create(16, 156250, 12500 / 2.2f);
            Real da = 0, db = 0;
            std::vector<Complex> blk(400), one;
            const size_t mb = a.decimate(&da, step, &cin[0], 1200, &blk[0], blk.size());
            for (size_t i = 0; i < 1200; i++) {
                Complex ci;
                if (b.decimate(&db, cin[i], &ci)) { one.push_back(ci); db += step; }
            }
            bool same = (one.size() == mb) && (da == db);
            for (size_t i = 0; same && i < mb; i++) same = (one[i] == blk[i]);
            printf("interpolator_per_sample_decimate %s n=%zu\n", same ? "same" : "DIFFERENT", one.size());
            // Tx pull loop (nfmmod.cpp:126-133), 48 kS/s -> 156.25 kS/s
            Interpolator c, d;
            c.create(16, 156250, 48000 / 2.2f);
            d.create(16, 156250, 48000 / 2.2f);
            Real dc = 0, dd = 0;
            const Real up = (Real) 48000 / (Real) 156250;
            std::vector<Complex> ublk(2000), uone;
            const size_t mu = c.interpolate(&dc, up, &cin[0], 300, &ublk[0], ublk.size());
            size_t i = 0;
            for (;;) {
                if (dd >= 1.0f && i >= 300) break;
                Complex ci;
                if (d.interpolate(&dd, i < 300 ? cin[i] : Complex(0, 0), &ci)) i++;
                uone.push_back(ci);
                dd += up;
            }
            same = (uone.size() == mu) && (dc == dd);
            for (size_t k = 0; same && k < mu; k++) same = (uone[k] == ublk[k]);
            printf("interpolator_per_sample_interpolate %s n=%zu\n", same ? "same" : "DIFFERENT", uone.size());
        }
        // NCO: per-sample nextIQ() == the block form
        {
            NCO n1, n2;
            n1.setFreq(15433, 156250);
            n2.setFreq(15433, 156250);
            std::vector<Complex> nb(500);
            n1.nextIQ(&nb[0], nb.size());
            bool same = (n1.phaseIncrement() == 404);
            for (size_t k = 0; same && k < 40; k++) same = (n2.nextIQ() == nb[k]);
            printf("nco inc=%d %s\n", n1.phaseIncrement(), same ? "same" : "DIFFERENT");
        }
        // SampleSinkFifo semantics (samplesinkfifo.cpp:113-231): overflow drops, two-span readBegin, readCommit; the device ring behaves alike
        {
            SampleSinkFifo fifo(1000);
            SampleVector v(700);
            for (size_t k = 0; k < v.size(); k++) { v[k].setReal((qint16) k); v[k].setImag((qint16) -(int) k); }
            uint w1 = fifo.write(v.begin(), v.end());
            SampleVector r(500);
            uint r1 = fifo.read(r.begin(), r.end());
            uint w2 = fifo.write(v.begin(), v.end());           // wraps: 300 at the end + 400 at the start
            uint w3 = fifo.write(v.begin(), v.end());           // only 100 fit: overflow drops 600
            SampleVector::iterator p1b, p1e, p2b, p2e;
            uint t = fifo.readBegin(900, &p1b, &p1e, &p2b, &p2e);
            printf("samplesinkfifo w=%u,%u,%u r=%u begin=%u spans=%zu+%zu first=%d second=%d fill=%u\n", w1, w2, w3, r1, t, (size_t) (p1e - p1b), (size_t) (p2e - p2b),
                   (int) p1b->real(), (int) p2b->real(), fifo.fill());
            fifo.readCommit(t);
            DeviceSampleSinkFifo dfifo(1000);
            uint dw1 = dfifo.write(v.begin(), v.end());
            uint dr1 = dfifo.read(r.begin(), r.end());
            uint dw2 = dfifo.write(v.begin(), v.end());
            uint dw3 = dfifo.write(v.begin(), v.end());
            const void *q1, *q2; uint n1 = 0, n2 = 0;
            uint dt = dfifo.readBegin(900, &q1, &n1, &q2, &n2);
            printf("devicesamplesinkfifo w=%u,%u,%u r=%u begin=%u spans=%u+%u fill=%u r0=%d\n", dw1, dw2, dw3, dr1, dt, n1, n2, dfifo.fill(), (int) r[499].real());
            printf("frequencyshift %d %d %d\n", DeviceSampleSource::calculateFrequencyShift(2, DeviceSampleSource::FC_POS_INFRA, 10000000),
                   DeviceSampleSource::calculateFrequencyShift(4, DeviceSampleSource::FC_POS_SUPRA, 10000000),
                   (int) DeviceSampleSource::calculateDeviceCenterFrequency(435000000ull, 0, 4, DeviceSampleSource::FC_POS_SUPRA, 10000000));
        }
        // decimators -> device FIFO -> bank, never through the host: Decimators output written on the device, the bank fed from the FIFO's spans
        {
            b200dsp_decim_t* dh = nullptr;
            b200dsp_cxx::check(b200dsp_decim_create(&dh, B200DSP_FMT_I16, B200DSP_FMT_I16, 12));
            ChannelBank bankA(10000000 / 4), bankB(10000000 / 4);
            CaptureSink sa, sb2;
            bankA.addChannel(&sa, 48000, 300000);
            bankB.addChannel(&sb2, 48000, 300000);
            // host route: decimate4_cen through the wrappers, then ChannelBank::feed
            SampleVector dec(nbSamples / 4);
            it = dec.begin();
            Decimators<qint32, qint16, SDR_RX_SAMP_SZ, 12> d4;
            d4.decimate4_cen(&it, buf, nbSamples * 2);
            bankA.feed(dec.begin(), it, false);
            // device route (C ABI): the same, the samples staying in device memory
            SampleVector viaDevice;
            {
                // a device buffer pair through the FIFO object itself: raw input in, decimated block out
                DeviceSampleSinkFifo rawIn((uint) nbSamples), ring(300000);
                SampleVector raw(nbSamples);
                for (int k = 0; k < nbSamples; k++) { raw[k].setReal(buf[2 * k]); raw[k].setImag(buf[2 * k + 1]); }
                rawIn.write(raw.begin(), raw.end());
                const void *a1, *a2; uint m1 = 0, m2 = 0;
                rawIn.readBegin((uint) nbSamples, &a1, &m1, &a2, &m2);
                DeviceSampleSinkFifo stage((uint) nbSamples / 4);
                const void *o1, *o2; uint k1 = 0, k2 = 0;
                SampleVector zeros(nbSamples / 4);
                stage.write(zeros.begin(), zeros.end());                 // reserve the ring's storage as the decimator's output buffer
                stage.readBegin((uint) nbSamples / 4, &o1, &k1, &o2, &k2);
                int64_t nd = 0;
                b200dsp_cxx::check(b200dsp_decim_run_dev(dh, 2, B200DSP_MODE_CEN, a1, 2ll * m1, (void*) o1, &nd, nullptr));
                b200dsp_cxx::check(b200dsp_decim_sync(dh));
                ring.writeDevice(o1, (uint) nd);                          // decimators -> FIFO (device to device)
                const void *p1, *p2; uint c1 = 0, c2 = 0;
                uint tot = ring.readBegin(300000, &p1, &c1, &p2, &c2);   // engine side: spans -> the bank, still on the device
                if (c1) b200dsp_cxx::check(b200dsp_bank_feed_dev(bankB.handle(), p1, c1, nullptr));
                b200dsp_cxx::check(b200dsp_bank_sync(bankB.handle()));
                ring.readCommit(tot);
                int64_t mch = 0;
                b200dsp_cxx::check(b200dsp_bank_fetch(bankB.handle(), 0, B200DSP_STAGE_CHANNELIZER, nullptr, (int64_t) 1 << 62, &mch));
                viaDevice.resize((size_t) mch);
                if (mch) b200dsp_cxx::check(b200dsp_bank_fetch(bankB.handle(), 0, B200DSP_STAGE_CHANNELIZER, &viaDevice[0], mch, &mch));
            }
            bool same = (viaDevice.size() == sa.captured.size()) && !viaDevice.empty();
            for (size_t k = 0; same && k < viaDevice.size(); k++) same = (viaDevice[k].real() == sa.captured[k].real() && viaDevice[k].imag() == sa.captured[k].imag());
            printf("device_fifo_route %s n=%zu\n", same ? "same" : "DIFFERENT", viaDevice.size());
            b200dsp_decim_destroy(dh);
            // split-I/Q overload == interleaved entry point (decimators.h:2858 vs :2966)
            std::vector<qint16> bi(4096), bq(4096);
            for (int k = 0; k < 4096; k++) { bi[k] = buf[2 * k]; bq[k] = buf[2 * k + 1]; }
            Decimators<qint32, qint16, SDR_RX_SAMP_SZ, 12> ds, dn;
            SampleVector os(4096), on(4096);
            SampleVector::iterator is = os.begin(), in2 = on.begin();
            ds.decimate16_cen(&is, &bi[0], &bq[0], 4096);
            dn.decimate16_cen(&in2, buf, 8192);
            same = (is - os.begin() == in2 - on.begin());
            for (size_t k = 0; same && k < (size_t) (is - os.begin()); k++) same = (os[k].real() == on[k].real() && os[k].imag() == on[k].imag());
            printf("split_iq_overload %s n=%zu\n", same ? "same" : "DIFFERENT", (size_t) (is - os.begin()));
        }
        {
            // Tx mirror: modulator -> UpChannelizer::pull (per sample, as the reference is driven) == pullBlock; then the device
            // plugin's Interpolators<qint16, SDR_TX_SAMP_SZ, 12>::interpolate8_cen into the device buffer (bladerfoutputthread.cpp)
            RampSource m1, m2;
            UpChannelizer u1(&m1, 1000), u2(&m2);
            u1.setOutputSampleRate(4000000); u1.configure(100000, 1200000);
            u2.setOutputSampleRate(4000000); u2.configure(100000, 1200000);
            SampleVector a(10000), b(10000);
            for (size_t k = 0; k < a.size(); k++) u1.pull(a[k]);
            u2.pullBlock(&b[0], (int) b.size());
            bool same = true;
            for (size_t k = 0; same && k < a.size(); k++) same = (a[k].real() == b[k].real() && a[k].imag() == b[k].imag());
            printf("upchannelizer %s rate=%d ofs=%d pulled=%d out=%016llx\n", same ? "same" : "DIFFERENT", u2.getCurrentInputSampleRate(),
                   u2.getCurrentCenterFrequency(), m2.k, (unsigned long long) fnv(&b[0], 2 * b.size()));
            Interpolators<qint16, SDR_TX_SAMP_SZ, 12> interp;
            std::vector<qint16> dev(10000 * 16 + 5, 77);
            SampleVector::iterator it = b.begin();
            interp.interpolate8_cen(&it, &dev[0], (qint32) dev.size());
            printf("interpolators8_cen consumed=%zu tail=%d out=%016llx\n", (size_t) (it - b.begin()), (int) dev[dev.size() - 1],
                   (unsigned long long) fnv(&dev[0], dev.size()));
        }
        {
            // demodulator back-end: per-sample phaseDiscriminatorDelta (the reference's call) == the block form
            PhaseDiscriminators pd1, pd2;
            pd1.setFMScaling(0.25f); pd2.setFMScaling(0.25f);
            std::vector<Complex> z(64);
            for (int k = 0; k < 64; k++) z[k] = Complex(1000.0f * std::cos(0.3f * k * k * 0.01f), 1000.0f * std::sin(0.3f * k * k * 0.01f));
            std::vector<Real> o1(64), o2(64), mg(64), dv(64);
            for (int k = 0; k < 64; k++) { double ms; Real dev; o1[k] = pd1.phaseDiscriminatorDelta(z[k], ms, dev); }
            pd2.phaseDiscriminatorDelta(&z[0], 64, &o2[0], &mg[0], &dv[0]);
            bool same = true;
            for (int k = 0; k < 64; k++) same = same && (o1[k] == o2[k]);
            printf("phasediscri %s\n", same ? "same" : "DIFFERENT");
            // SSB channel filter: the reference's per-sample runSSB (ssbdemod.cpp:171) == the block form
            {
                fftfilt f1(300.0f / 48000.0f, 3000.0f / 48000.0f, 1024), f2(300.0f / 48000.0f, 3000.0f / 48000.0f, 1024);
                std::vector<fftfilt::cmplx> xin(2000), oa, ob(2048);
                for (int k = 0; k < 2000; k++) xin[k] = fftfilt::cmplx(1000.0f * std::cos(0.05f * k), 700.0f * std::sin(0.021f * k));
                for (int k = 0; k < 2000; k++) {
                    fftfilt::cmplx* sideband = 0;
                    int n_out = f1.runSSB(xin[k], &sideband, true);
                    for (int q = 0; q < n_out; q++) oa.push_back(sideband[q]);
                }
                int nb = f2.runSSB(&xin[0], 2000, &ob[0], 2048, true);
                bool sameF = ((int) oa.size() == nb);
                for (int k = 0; sameF && k < nb; k++) sameF = (oa[k] == ob[k]);
                printf("fftfilt %s n=%d\n", sameF ? "same" : "DIFFERENT", nb);
            }
            // FileRecord: record, read the header back the way the file-source plugin does
            const char* path = "/tmp/b200dsp_cxx_dropin.sdriq";
            FileRecord rec(path);
            rec.setSampleRateAndFrequency(2400000, 434000000ull);
            SampleVector v(1000);
            for (int k = 0; k < 1000; k++) { v[k].setReal((qint16) (k - 500)); v[k].setImag((qint16) (3 * k)); }
            rec.feed(v.begin(), v.end(), false);            // not recording yet: dropped
            rec.startRecording();
            rec.feed(v.begin(), v.begin() + 300, false);
            rec.feed(v.begin() + 300, v.end(), false);
            rec.stopRecording();
            std::ifstream f(path, std::ios::binary);
            FileRecord::Header hd;
            FileRecord::readHeader(f, hd);
            f.seekg(0, std::ios::end);
            printf("filerecord rate=%d fc=%llu size=%u count=%llu bytes=%lld\n", hd.sampleRate, (unsigned long long) hd.centerFrequency, hd.sampleSize,
                   (unsigned long long) rec.getByteCount(), (long long) f.tellg());
        }
        delete[] buf;
    } catch (const std::exception& e) {
        fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
    return 0;
}
