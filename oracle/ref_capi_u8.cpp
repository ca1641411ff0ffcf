// ORACLE (test infrastructure) -- C-ABI driver for the reference's DecimatorsU<> (unsigned 8-bit device samples,
// sdrbase/dsp/decimatorsu.h:175-215; RTL-SDR instantiation plugins/samplesource/rtlsdr/rtlsdrthread.h:55).
// A translation unit of its own: decimatorsu.h re-declares decimation_shifts<> and cannot share one with decimators.h
// (in the reference each device plugin includes exactly one of the two).  Compiled with the UNMODIFIED reference headers
// where they lie; never linked into the product.
#include <stdint.h>
#include <string.h>
#include <cstddef>
#include <cmath>
#include "dsp/decimatorsu.h"
#include "util/movingaverage.h"

namespace {

enum { MODE_INF = 0, MODE_SUP = 1, MODE_CEN = 2 };

struct DecimU8 { DecimatorsU<qint32, quint8, SDR_RX_SAMP_SZ, 8, 127> d; SampleVector out; };

bool dispatch_u8(DecimU8* h, int log2, int mode, SampleVector::iterator* it, const quint8* buf, int len)
{
    if (log2 == 0) { h->d.decimate1(it, buf, len); return true; }
#define ORACLE_CASE(N, L) \
    case L: \
        if (mode == MODE_INF) h->d.decimate##N##_inf(it, buf, len); \
        else if (mode == MODE_SUP) h->d.decimate##N##_sup(it, buf, len); \
        else h->d.decimate##N##_cen(it, buf, len); \
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

} // namespace

// DSPDeviceSourceEngine::iqCorrections, DC branch.  dspdevicesourceengine.cpp pulls QThread / the plugin API and is not
// buildable here, so its loop (:175-183,254-261) is restated around the reference's own MovingAverageUtil and Sample types.
struct IqCorr {
    MovingAverageUtil<int32_t, int64_t, 1024> m_iBeta;      // dspdevicesourceengine.h:106-107
    MovingAverageUtil<int32_t, int64_t, 1024> m_qBeta;
    // floating-point imbalance correction (IMBALANCE_INT is not defined in the reference build): dspdevicesourceengine.h:119-125
    MovingAverageUtil<float, double, 128> m_avgII;
    MovingAverageUtil<float, double, 128> m_avgIQ;
    MovingAverageUtil<float, double, 128> m_avgII2;
    MovingAverageUtil<float, double, 128> m_avgQQ2;
    MovingAverageUtil<double, double, 128> m_avgPhi;
    MovingAverageUtil<double, double, 128> m_avgAmp;
};

extern "C" {

void* ref_iqcorr_create() { return new IqCorr; }
void ref_iqcorr_destroy(void* p) { delete (IqCorr*) p; }
// in place on n Samples (int16 I,Q interleaved)
void ref_iqcorr_dc(void* p, int16_t* iq, int n)
{
    IqCorr* h = (IqCorr*) p;
    Sample* begin = (Sample*) iq;
    for (Sample* it = begin; it < begin + n; it++)
    {
        h->m_iBeta(it->real());
        h->m_qBeta(it->imag());
        it->m_real -= (int32_t) h->m_iBeta;
        it->m_imag -= (int32_t) h->m_qBeta;
    }
}

// the imbalance branch of the same loop (dspdevicesourceengine.cpp:175-183,219-252), restated statement by statement
void ref_iqcorr_imbalance(void* p, int16_t* iq, int n)
{
    IqCorr* h = (IqCorr*) p;
    Sample* begin = (Sample*) iq;
    for (Sample* it = begin; it < begin + n; it++)
    {
        h->m_iBeta(it->real());
        h->m_qBeta(it->imag());
        float xi = (it->m_real - (int32_t) h->m_iBeta) / SDR_RX_SCALEF;
        float xq = (it->m_imag - (int32_t) h->m_qBeta) / SDR_RX_SCALEF;
        h->m_avgII(xi*xi);
        h->m_avgIQ(xi*xq);
        if (h->m_avgII.asDouble() != 0) {
            h->m_avgPhi(h->m_avgIQ.asDouble()/h->m_avgII.asDouble());
        }
        float& yi = xi;
        float yq = xq - h->m_avgPhi.asDouble()*xi;
        h->m_avgII2(yi*yi);
        h->m_avgQQ2(yq*yq);
        if (h->m_avgQQ2.asDouble() != 0) {
            h->m_avgAmp(sqrt(h->m_avgII2.asDouble() / h->m_avgQQ2.asDouble()));
        }
        float& zi = yi;
        float zq = h->m_avgAmp.asDouble() * yq;
        it->m_real = zi * SDR_RX_SCALEF;
        it->m_imag = zq * SDR_RX_SCALEF;
    }
}

void* ref_decim_u8_create() { return new DecimU8; }
void ref_decim_u8_destroy(void* p) { delete (DecimU8*) p; }

// returns number of Samples written to out (int16 I,Q interleaved), or -1
int ref_decim_u8_run(void* p, int log2, int mode, const uint8_t* buf, int len, int16_t* out)
{
    DecimU8* h = (DecimU8*) p;
    std::size_t need = (std::size_t) (len / 2) + 8;
    if (h->out.size() < need) h->out.resize(need);
    SampleVector::iterator it = h->out.begin();
    if (!dispatch_u8(h, log2, mode, &it, (const quint8*) buf, len)) return -1;
    int n = (int) (it - h->out.begin());
    if (n > 0) memcpy(out, &h->out[0], (std::size_t) n * sizeof(Sample));
    return n;
}

}
