// devicesamplesource.h — Qt-free DeviceSampleSource (sdrbase/dsp/devicesamplesource.h:36-115, devicesamplesource.cpp:43-117):
// the interface a device plugin implements and the engine pulls from.  The sample-carrying part keeps its shape -- the source
// owns a SampleSinkFifo the plugin's thread writes decimated samples into and the engine reads from (getSampleFifo()) -- and
// the two static frequency helpers are restated; serialisation, message queues and the REST hooks are control plane.
#ifndef SDRANGEL_B200_DSP_DEVICESAMPLESOURCE_H
#define SDRANGEL_B200_DSP_DEVICESAMPLESOURCE_H
#include <string>
#include "samplesinkfifo.h"

class DeviceSampleSource {
public:
    typedef enum { FC_POS_INFRA = 0, FC_POS_SUPRA, FC_POS_CENTER } fcPos_t;     // == B200DSP_MODE_INF / _SUP / _CEN

    DeviceSampleSource() {}
    virtual ~DeviceSampleSource() {}
    virtual void init() = 0;
    virtual bool start() = 0;
    virtual void stop() = 0;
    virtual const std::string& getDeviceDescription() const = 0;
    virtual int getSampleRate() const = 0;                  //!< Sample rate exposed by the source
    virtual quint64 getCenterFrequency() const = 0;         //!< Center frequency exposed by the source
    virtual void setCenterFrequency(qint64 centerFrequency) = 0;
    SampleSinkFifo* getSampleFifo() { return &m_sampleFifo; }

    /** devicesamplesource.cpp:43-69 */
    static qint64 calculateDeviceCenterFrequency(quint64 centerFrequency, qint64 transverterDeltaFrequency, int log2Decim, fcPos_t fcPos,
                                                 quint32 devSampleRate, bool transverterMode = false)
    {
        qint64 deviceCenterFrequency = (qint64) centerFrequency;
        deviceCenterFrequency -= transverterMode ? transverterDeltaFrequency : 0;
        deviceCenterFrequency = deviceCenterFrequency < 0 ? 0 : deviceCenterFrequency;
        deviceCenterFrequency -= calculateFrequencyShift(log2Decim, fcPos, devSampleRate);
        return deviceCenterFrequency;
    }
    /** devicesamplesource.cpp:87-117: where the decimators' Inf / Sup variants put the wanted band */
    static qint32 calculateFrequencyShift(int log2Decim, fcPos_t fcPos, quint32 devSampleRate)
    {
        if (log2Decim == 0) return 0;
        const quint32 div = (log2Decim < 3) ? (1u << (log2Decim + 1)) : (1u << log2Decim);
        if (fcPos == FC_POS_INFRA) return -(qint32) (devSampleRate / div);
        if (fcPos == FC_POS_SUPRA) return (qint32) (devSampleRate / div);
        return 0;
    }
protected:
    SampleSinkFifo m_sampleFifo;
};
#endif
