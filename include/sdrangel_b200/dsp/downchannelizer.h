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

/** The multi-channel form: many (DownChannelizer [+ NCO + Interpolator]) channels fed from ONE baseband, sharing the
 *  half-band tree between channels (b200dsp_bank_*).  Each channel has its own sink, fed once per feed() with that
 *  channel's decimated samples, exactly what each DownChannelizer of the reference would hand to its plugin
 *  (downchannelizer.cpp:86-89); channels with a front-end also deliver the NCO-mixed, resampled complex stream
 *  (nfmdemod.cpp:150-155,315) through feedFrontend(). */
class ChannelBank : public BasebandSampleSink {
public:
    struct ChannelSink : BasebandSampleSink {
        virtual void feedFrontend(const std::vector<Complex>& samples) { (void) samples; }
    };
    explicit ChannelBank(int inputSampleRate) : m_bank(nullptr) { b200dsp_cxx::check(b200dsp_bank_create(&m_bank, inputSampleRate)); }
    virtual ~ChannelBank() { b200dsp_bank_destroy(m_bank); }
    /** == DSPConfigureChannelizer(sampleRate, centerFrequency) for a new channel; returns its index.  outRate / offset: what
     *  MsgChannelizerNotification would report */
    int addChannel(BasebandSampleSink* sink, int sampleRate, int centerFrequency, int* outRate = nullptr, int* frequencyOffset = nullptr)
    {
        int id = -1, r = 0, o = 0;
        b200dsp_cxx::check(b200dsp_bank_add_channel(m_bank, sampleRate, centerFrequency, &id, &r, &o));
        if (outRate) *outRate = r;
        if (frequencyOffset) *frequencyOffset = o;
        m_sinks.push_back(sink); m_fe.push_back(nullptr);
        return id;
    }
    /** == m_nco.setFreq(ncoFreq, rate); m_interpolator.create(16, rate, cutoff); distance = rate / outRate (nfmdemod.cpp:462-470) */
    void setFrontend(int channel, ChannelSink* sink, float ncoFreq, double cutoff, int outRate, int phaseSteps = 16, double tapsPerPhase = 4.5)
    {
        b200dsp_cxx::check(b200dsp_bank_set_frontend(m_bank, channel, ncoFreq, phaseSteps, cutoff, tapsPerPhase, outRate));
        m_fe[(size_t) channel] = sink;
    }
    virtual void start() {}
    virtual void stop() {}
    virtual void feed(const SampleVector::const_iterator& begin, const SampleVector::const_iterator& end, bool positiveOnly)
    {
        const int64_t n = end - begin;
        b200dsp_cxx::check(b200dsp_bank_feed(m_bank, n > 0 ? (const int16_t*) &(*begin) : nullptr, n));
        for (size_t c = 0; c < m_sinks.size(); ++c) {
            int64_t m = 0;
            if (m_sinks[c]) {
                b200dsp_cxx::check(b200dsp_bank_fetch(m_bank, (int) c, B200DSP_STAGE_CHANNELIZER, nullptr, (int64_t) 1 << 62, &m));
                m_buf.resize((size_t) m);
                if (m > 0) b200dsp_cxx::check(b200dsp_bank_fetch(m_bank, (int) c, B200DSP_STAGE_CHANNELIZER, &m_buf[0], m, &m));
                m_sinks[c]->feed(m_buf.begin(), m_buf.end(), positiveOnly);
            }
            if (m_fe[c]) {
                b200dsp_cxx::check(b200dsp_bank_fetch(m_bank, (int) c, B200DSP_STAGE_FRONTEND, nullptr, (int64_t) 1 << 62, &m));
                m_cbuf.resize((size_t) m);
                if (m > 0) b200dsp_cxx::check(b200dsp_bank_fetch(m_bank, (int) c, B200DSP_STAGE_FRONTEND, &m_cbuf[0], m, &m));
                m_fe[c]->feedFrontend(m_cbuf);
            }
        }
    }
    b200dsp_bank_t* handle() { return m_bank; }
private:
    ChannelBank(const ChannelBank&);
    ChannelBank& operator=(const ChannelBank&);
    b200dsp_bank_t* m_bank;
    std::vector<BasebandSampleSink*> m_sinks;
    std::vector<ChannelSink*> m_fe;
    SampleVector m_buf;
    std::vector<Complex> m_cbuf;
};
#endif
