// decimators.h — drop-in for the reference's Decimators<StorageType, T, SdrBits, InputBits>
// (sdrbase/dsp/decimators.h:277-341): same template, same 20 interleaved-I/Q entry points, same observable behaviour
// (writes through *it and advances it, drops the trailing partial block, carries filter state between calls) — the
// arithmetic runs in the hb64 cascade kernel through b200dsp_decim_run.
#ifndef SDRANGEL_B200_DSP_DECIMATORS_H
#define SDRANGEL_B200_DSP_DECIMATORS_H
#include "dsptypes.h"

namespace b200dsp_cxx {
template<int IN_FMT, int OUT_FMT, typename TIn, typename OutVec>
class DecimatorsImpl {
public:
    explicit DecimatorsImpl(int inputBits) : m_h(nullptr) { check(b200dsp_decim_create(&m_h, IN_FMT, OUT_FMT, inputBits)); }
    ~DecimatorsImpl() { b200dsp_decim_destroy(m_h); }
    DecimatorsImpl(const DecimatorsImpl&) = delete;
    DecimatorsImpl& operator=(const DecimatorsImpl&) = delete;
    b200dsp_decim_t* handle() { return m_h; }
    void setExactFloat(bool exact) { check(b200dsp_decim_set_exact_float(m_h, exact ? 1 : 0)); }
protected:
    void run(int log2, int mode, typename OutVec::iterator* it, const TIn* buf, qint32 len)
    {
        int32_t n = 0;
        // the caller sized the vector (e.g. airspythread.cpp:31); &(**it) is its contiguous storage
        check(b200dsp_decim_run(m_h, log2, mode, buf, len, len > 1 ? (void*) &(**it) : nullptr, &n));
        *it += n;
    }
    /** the overloads on separate I and Q arrays (decimators.h:359-371,395-417,2638-...): len samples in each */
    void runSplit(int log2, int mode, typename OutVec::iterator* it, const TIn* bufI, const TIn* bufQ, qint32 len)
    {
        int32_t n = 0;
        check(b200dsp_decim_run_split(m_h, log2, mode, bufI, bufQ, len, len > 0 ? (void*) &(**it) : nullptr, &n));
        *it += n;
    }
    b200dsp_decim_t* m_h;
};
} // namespace b200dsp_cxx

#define B200DSP_DECIM_ENTRY_POINTS(ItVec, TIn)                                                                                      \
    void decimate1(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(0, B200DSP_MODE_CEN, it, buf, len); }               \
    void decimate2_inf(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(1, B200DSP_MODE_INF, it, buf, len); }           \
    void decimate2_sup(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(1, B200DSP_MODE_SUP, it, buf, len); }           \
    void decimate2_cen(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(1, B200DSP_MODE_CEN, it, buf, len); }           \
    void decimate4_inf(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(2, B200DSP_MODE_INF, it, buf, len); }           \
    void decimate4_sup(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(2, B200DSP_MODE_SUP, it, buf, len); }           \
    void decimate4_cen(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(2, B200DSP_MODE_CEN, it, buf, len); }           \
    void decimate8_inf(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(3, B200DSP_MODE_INF, it, buf, len); }           \
    void decimate8_sup(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(3, B200DSP_MODE_SUP, it, buf, len); }           \
    void decimate8_cen(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(3, B200DSP_MODE_CEN, it, buf, len); }           \
    void decimate16_inf(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(4, B200DSP_MODE_INF, it, buf, len); }          \
    void decimate16_sup(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(4, B200DSP_MODE_SUP, it, buf, len); }          \
    void decimate16_cen(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(4, B200DSP_MODE_CEN, it, buf, len); }          \
    void decimate32_inf(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(5, B200DSP_MODE_INF, it, buf, len); }          \
    void decimate32_sup(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(5, B200DSP_MODE_SUP, it, buf, len); }          \
    void decimate32_cen(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(5, B200DSP_MODE_CEN, it, buf, len); }          \
    void decimate64_inf(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(6, B200DSP_MODE_INF, it, buf, len); }          \
    void decimate64_sup(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(6, B200DSP_MODE_SUP, it, buf, len); }          \
    void decimate64_cen(ItVec::iterator* it, const TIn* buf, qint32 len) { this->run(6, B200DSP_MODE_CEN, it, buf, len); }

/** Decimators with integer input and integer output (decimators.h:277-341).  StorageType is the reference's filter
 *  accumulator type (qint32 in 16-bit Rx mode); T is qint16 (Airspy, LimeSDR, PlutoSDR, BladeRF, SDRplay, TestSource) or
 *  qint8 with InputBits 8 (HackRF, hackrfinputthread.h:57).  The unsigned variant is DecimatorsU (decimatorsu.h). */
template<typename StorageType, typename T, uint SdrBits, uint InputBits>
class Decimators : public b200dsp_cxx::DecimatorsImpl<(sizeof(T) == 1 ? B200DSP_FMT_I8 : B200DSP_FMT_I16), B200DSP_FMT_I16, T, SampleVector> {
    static_assert((sizeof(T) == 2 || (sizeof(T) == 1 && InputBits == 8)) && SdrBits == 16 && sizeof(StorageType) == 4,
                  "16-bit Rx mode: Decimators<qint32, qint16, 16, {8,12,16}> or Decimators<qint32, qint8, 16, 8>");
public:
    Decimators() : b200dsp_cxx::DecimatorsImpl<(sizeof(T) == 1 ? B200DSP_FMT_I8 : B200DSP_FMT_I16), B200DSP_FMT_I16, T, SampleVector>(InputBits) {}
    B200DSP_DECIM_ENTRY_POINTS(SampleVector, T)
    /** decimators.h:374-393: unfiltered /2 of offset-255 data */
    void decimate2_u(SampleVector::iterator* it, const T* buf, qint32 len) { this->run(1, B200DSP_MODE_U, it, buf, len); }
    // the overloads on separate I and Q arrays that the reference defines (decimators.h:359-371,395-417,2638-3888)
    void decimate1(SampleVector::iterator* it, const T* bufI, const T* bufQ, qint32 len) { this->runSplit(0, B200DSP_MODE_CEN, it, bufI, bufQ, len); }
    void decimate2_u(SampleVector::iterator* it, const T* bufI, const T* bufQ, qint32 len) { this->runSplit(1, B200DSP_MODE_U, it, bufI, bufQ, len); }
    void decimate2_cen(SampleVector::iterator* it, const T* bufI, const T* bufQ, qint32 len) { this->runSplit(1, B200DSP_MODE_CEN, it, bufI, bufQ, len); }
    void decimate4_cen(SampleVector::iterator* it, const T* bufI, const T* bufQ, qint32 len) { this->runSplit(2, B200DSP_MODE_CEN, it, bufI, bufQ, len); }
    void decimate8_cen(SampleVector::iterator* it, const T* bufI, const T* bufQ, qint32 len) { this->runSplit(3, B200DSP_MODE_CEN, it, bufI, bufQ, len); }
    void decimate16_cen(SampleVector::iterator* it, const T* bufI, const T* bufQ, qint32 len) { this->runSplit(4, B200DSP_MODE_CEN, it, bufI, bufQ, len); }
    void decimate32_cen(SampleVector::iterator* it, const T* bufI, const T* bufQ, qint32 len) { this->runSplit(5, B200DSP_MODE_CEN, it, bufI, bufQ, len); }
    void decimate64_cen(SampleVector::iterator* it, const T* bufI, const T* bufQ, qint32 len) { this->runSplit(6, B200DSP_MODE_CEN, it, bufI, bufQ, len); }
};
#endif
