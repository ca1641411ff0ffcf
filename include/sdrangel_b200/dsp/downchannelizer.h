// downchannelizer.h — drop-in for DownChannelizer (sdrbase/dsp/downchannelizer.h:28-119, downchannelizer.cpp:50-91,165-189):
// a BasebandSampleSink that decimates towards the requested channel and feeds the result to its own sink.
// The reference configures it with two messages (DSPSignalNotification, DSPConfigureChannelizer) and answers with
// MsgChannelizerNotification; here the same information travels through plain methods.  One DownChannelizer == a
// one-channel bank; ChannelBank (below) is the multi-channel form that shares the half-band tree between channels.
#ifndef SDRANGEL_B200_DSP_DOWNCHANNELIZER_H
#define SDRANGEL_B200_DSP_DOWNCHANNELIZER_H
#include "basebandsamplesink.h"

class DownChannelizer : public BasebandSampleSink {
public:
    explicit DownChannelizer(BasebandSampleSink* sampleSink) :
        m_sampleSink(sampleSink), m_bank(nullptr), m_inputSampleRate(0), m_requestedOutputSampleRate(0), m_requestedCenterFrequency(0),
        m_currentOutputSampleRate(0), m_currentCenterFrequency(0), m_chan(-1) {}
    virtual ~DownChannelizer() { if (m_bank) b200dsp_bank_destroy(m_bank); }
    /** == handleMessage(DSPSignalNotification(sampleRate, ...)) (downchannelizer.cpp:111-124) */
    void setInputSampleRate(int sampleRate) { m_inputSampleRate = sampleRate; applyConfiguration(); }
    /** == configure(messageQueue, sampleRate, centerFrequency) -> DSPConfigureChannelizer (downchannelizer.cpp:105-109,125-140) */
    void configure(int sampleRate, int centerFrequency)
    {
        m_requestedOutputSampleRate = sampleRate;
        m_requestedCenterFrequency = centerFrequency;
        applyConfiguration();
    }
    int getInputSampleRate() const { return m_inputSampleRate; }
    int getRequestedCenterFrequency() const { return m_requestedCenterFrequency; }
    /** what MsgChannelizerNotification carries (downchannelizer.cpp:186-187) */
    int getCurrentOutputSampleRate() const { return m_currentOutputSampleRate; }
    int getCurrentCenterFrequency() const { return m_currentCenterFrequency; }

    virtual void start() {}
    virtual void stop() {}
    virtual void feed(const SampleVector::const_iterator& begin, const SampleVector::const_iterator& end, bool positiveOnly)
    {
        if (!m_sampleSink) return;
        if (!m_bank) { m_sampleSink->feed(begin, end, positiveOnly); return; }     // no filter chain: forwarded unchanged (:57-60)
        const int64_t n = end - begin;
        b200dsp_cxx::check(b200dsp_bank_feed(m_bank, n > 0 ? (const int16_t*) &(*begin) : nullptr, n));
        int64_t m = 0;
        b200dsp_cxx::check(b200dsp_bank_fetch(m_bank, m_chan, B200DSP_STAGE_CHANNELIZER, nullptr, (int64_t) 1 << 62, &m));
        m_sampleBuffer.resize((size_t) m);
        if (m > 0) b200dsp_cxx::check(b200dsp_bank_fetch(m_bank, m_chan, B200DSP_STAGE_CHANNELIZER, &m_sampleBuffer[0], m, &m));
        m_sampleSink->feed(m_sampleBuffer.begin(), m_sampleBuffer.end(), positiveOnly);   // (:88)
        m_sampleBuffer.clear();
    }
protected:
    void applyConfiguration()
    {
        if (m_bank) { b200dsp_bank_destroy(m_bank); m_bank = nullptr; }
        if (m_inputSampleRate == 0 || m_requestedOutputSampleRate == 0) return;         // (:165-167)
        b200dsp_cxx::check(b200dsp_bank_create(&m_bank, m_inputSampleRate));
        b200dsp_cxx::check(b200dsp_bank_add_channel(m_bank, m_requestedOutputSampleRate, m_requestedCenterFrequency, &m_chan,
                                                    &m_currentOutputSampleRate, &m_currentCenterFrequency));
    }
    BasebandSampleSink* m_sampleSink;
    b200dsp_bank_t* m_bank;
    int m_inputSampleRate, m_requestedOutputSampleRate, m_requestedCenterFrequency, m_currentOutputSampleRate, m_currentCenterFrequency;
    int m_chan;
    SampleVector m_sampleBuffer;
};
#endif
