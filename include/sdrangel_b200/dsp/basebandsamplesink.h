// basebandsamplesink.h — Qt-free restatement of the sink interface the hot path sits behind
// (sdrbase/dsp/basebandsamplesink.h:29-71).  Message plumbing (QObject, MessageQueue) is control plane and out of scope;
// the sample-carrying virtuals keep their signatures.
#ifndef SDRANGEL_B200_DSP_BASEBANDSAMPLESINK_H
#define SDRANGEL_B200_DSP_BASEBANDSAMPLESINK_H
#include "dsptypes.h"
class BasebandSampleSink {
public:
    virtual ~BasebandSampleSink() {}
    virtual void start() = 0;
    virtual void stop() = 0;
    virtual void feed(const SampleVector::const_iterator& begin, const SampleVector::const_iterator& end, bool positiveOnly) = 0;
};
#endif
