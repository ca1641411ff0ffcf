// decimatorsff.h — drop-in for DecimatorsFF (sdrbase/dsp/decimatorsff.h): float in, FSample out.
#ifndef SDRANGEL_B200_DSP_DECIMATORSFF_H
#define SDRANGEL_B200_DSP_DECIMATORSFF_H
#include "decimators.h"
class DecimatorsFF : public b200dsp_cxx::DecimatorsImpl<B200DSP_FMT_F32, B200DSP_FMT_F32, float, FSampleVector> {
public:
    DecimatorsFF() : b200dsp_cxx::DecimatorsImpl<B200DSP_FMT_F32, B200DSP_FMT_F32, float, FSampleVector>(16) {}
    B200DSP_DECIM_ENTRY_POINTS(FSampleVector, float)
};
#endif
