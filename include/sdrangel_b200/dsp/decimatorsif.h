// decimatorsif.h — drop-in for DecimatorsIF<T, InputBits> (sdrbase/dsp/decimatorsif.h:53-83): int16 in, FSample out,
// scaled by decimation_scale<InputBits>::scaleIn.
#ifndef SDRANGEL_B200_DSP_DECIMATORSIF_H
#define SDRANGEL_B200_DSP_DECIMATORSIF_H
#include "decimators.h"
template<typename T, uint InputBits>
class DecimatorsIF : public b200dsp_cxx::DecimatorsImpl<B200DSP_FMT_I16, B200DSP_FMT_F32, T, FSampleVector> {
    static_assert(sizeof(T) == 2, "DecimatorsIF<qint16, {8,12,16}>");
public:
    DecimatorsIF() : b200dsp_cxx::DecimatorsImpl<B200DSP_FMT_I16, B200DSP_FMT_F32, T, FSampleVector>(InputBits) {}
    B200DSP_DECIM_ENTRY_POINTS(FSampleVector, T)
};
#endif
