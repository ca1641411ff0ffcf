// decimatorsu.h — drop-in for the reference's DecimatorsU<StorageType, T, SdrBits, InputBits, Shift>
// (sdrbase/dsp/decimatorsu.h:175-215): unsigned 8-bit device samples, sample = byte - Shift, then the Decimators<>
// arithmetic (the two headers differ only by that subtraction).  RTL-SDR: DecimatorsU<qint32, quint8, SDR_RX_SAMP_SZ, 8, 127>
// (plugins/samplesource/rtlsdr/rtlsdrthread.h:55, call sites rtlsdrthread.cpp:97-173).
#ifndef SDRANGEL_B200_DSP_DECIMATORSU_H
#define SDRANGEL_B200_DSP_DECIMATORSU_H
#include "decimators.h"

template<typename StorageType, typename T, uint SdrBits, uint InputBits, int Shift>
class DecimatorsU : public b200dsp_cxx::DecimatorsImpl<B200DSP_FMT_U8, B200DSP_FMT_I16, T, SampleVector> {
    static_assert(sizeof(T) == 1 && SdrBits == 16 && InputBits == 8 && sizeof(StorageType) == 4 && Shift >= 0 && Shift <= 255,
                  "16-bit Rx mode: DecimatorsU<qint32, quint8, 16, 8, Shift>");
public:
    DecimatorsU() : b200dsp_cxx::DecimatorsImpl<B200DSP_FMT_U8, B200DSP_FMT_I16, T, SampleVector>(InputBits)
    {
        b200dsp_cxx::check(b200dsp_decim_set_shift(this->m_h, Shift));
    }
    B200DSP_DECIM_ENTRY_POINTS(SampleVector, T)
};
#endif
