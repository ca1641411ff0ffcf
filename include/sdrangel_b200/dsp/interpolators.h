// interpolators.h — drop-in for the reference's Interpolators<T, SdrBits, OutputBits> (sdrbase/dsp/interpolators.h:104-135):
// same template, same seven entry points, same observable behaviour (reads through *it and advances it by len / (2N) samples,
// fills buf block by block, leaves a trailing partial block alone, carries the six half-band rings between calls and factors;
// interpolate64_cen leaves scalars 110..127 of every block untouched like the reference's store list) — the arithmetic runs
// in the fused cascade kernel through b200dsp_interps_run.
#ifndef SDRANGEL_B200_DSP_INTERPOLATORS_H
#define SDRANGEL_B200_DSP_INTERPOLATORS_H
#include "dsptypes.h"

template<typename T, uint SdrBits, uint OutputBits>
class Interpolators {
    static_assert(SdrBits == 16 && ((sizeof(T) == 2 && (OutputBits == 12 || OutputBits == 16)) || (sizeof(T) == 1 && OutputBits == 8)),
                  "SDR_TX_SAMP_SZ 16: Interpolators<qint16,16,{12,16}> or Interpolators<qint8,16,8>");
public:
    Interpolators() : m_h(nullptr) { b200dsp_cxx::check(b200dsp_interps_create(&m_h, sizeof(T) == 1 ? B200DSP_FMT_I8 : B200DSP_FMT_I16, OutputBits)); }
    ~Interpolators() { b200dsp_interps_destroy(m_h); }
    Interpolators(const Interpolators&) = delete;
    Interpolators& operator=(const Interpolators&) = delete;
    b200dsp_interps_t* handle() { return m_h; }
    void interpolate1(SampleVector::iterator* it, T* buf, qint32 len) { run(0, it, buf, len); }
    void interpolate2_cen(SampleVector::iterator* it, T* buf, qint32 len) { run(1, it, buf, len); }
    void interpolate4_cen(SampleVector::iterator* it, T* buf, qint32 len) { run(2, it, buf, len); }
    void interpolate8_cen(SampleVector::iterator* it, T* buf, qint32 len) { run(3, it, buf, len); }
    void interpolate16_cen(SampleVector::iterator* it, T* buf, qint32 len) { run(4, it, buf, len); }
    void interpolate32_cen(SampleVector::iterator* it, T* buf, qint32 len) { run(5, it, buf, len); }
    void interpolate64_cen(SampleVector::iterator* it, T* buf, qint32 len) { run(6, it, buf, len); }
private:
    void run(int log2, SampleVector::iterator* it, T* buf, qint32 len)
    {
        int32_t n = 0;
        const bool any = b200dsp_interps_in_count(log2, len) > 0;
        b200dsp_cxx::check(b200dsp_interps_run(m_h, log2, any ? (const int16_t*) &(**it) : nullptr, buf, len, &n));
        *it += n;
    }
    b200dsp_interps_t* m_h;
};
#endif
