// basebandsamplesource.h — Qt-free restatement of the source interface the Tx path sits behind
// (sdrbase/dsp/basebandsamplesource.h:30-77): a modulator hands out one Sample per pull().  pullBlock() is the block form a
// GPU-side consumer uses (default: n calls of pull); message plumbing and the SampleSourceFifo are control plane, out of scope.
#ifndef SDRANGEL_B200_DSP_BASEBANDSAMPLESOURCE_H
#define SDRANGEL_B200_DSP_BASEBANDSAMPLESOURCE_H
#include "dsptypes.h"
class BasebandSampleSource {
public:
    virtual ~BasebandSampleSource() {}
    virtual void start() = 0;
    virtual void stop() = 0;
    virtual void pull(Sample& sample) = 0;
    virtual void pullAudio(int nbSamples) { (void) nbSamples; }
    virtual void pullBlock(Sample* samples, int nbSamples) { for (int i = 0; i < nbSamples; i++) pull(samples[i]); }
};
#endif
