// decimatorsfi.h — drop-in for DecimatorsFI (sdrbase/dsp/decimatorsfi.h:26-55): float in, int16 Sample out.
#ifndef SDRANGEL_B200_DSP_DECIMATORSFI_H
#define SDRANGEL_B200_DSP_DECIMATORSFI_H
#include "decimators.h"
class DecimatorsFI : public b200dsp_cxx::DecimatorsImpl<B200DSP_FMT_F32, B200DSP_FMT_I16, float, SampleVector> {
public:
    DecimatorsFI() : b200dsp_cxx::DecimatorsImpl<B200DSP_FMT_F32, B200DSP_FMT_I16, float, SampleVector>(16) {}
    B200DSP_DECIM_ENTRY_POINTS(SampleVector, float)
};
#endif
