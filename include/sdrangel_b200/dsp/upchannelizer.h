// upchannelizer.h — drop-in for UpChannelizer (sdrbase/dsp/upchannelizer.h:36-123, upchannelizer.cpp:51-104,175-209): a
// BasebandSampleSource that interpolates its modulator's samples up to the device sink's rate at the channel's offset.
// The reference is driven one output sample at a time; the arithmetic here runs in blocks on the GPU (b200dsp_upchan_pull):
// pull(Sample&) hands out samples of the current block and computes the next block of `blockSize` outputs when it runs dry,
// pulling from the modulator exactly the samples those outputs consume, in order, only earlier than the reference would.
// pullBlock() is the block call a sink engine restructured for the GPU makes (BasebandSampleSource::feed's loop,
// basebandsamplesource.h:45-56).  Configuration travels through plain methods instead of messages.
#ifndef SDRANGEL_B200_DSP_UPCHANNELIZER_H
#define SDRANGEL_B200_DSP_UPCHANNELIZER_H
#include "basebandsamplesource.h"

class UpChannelizer : public BasebandSampleSource {
public:
    explicit UpChannelizer(BasebandSampleSource* sampleSource, int blockSize = 4096) :
        m_sampleSource(sampleSource), m_h(nullptr), m_outputSampleRate(0), m_requestedInputSampleRate(0), m_requestedCenterFrequency(0),
        m_currentInputSampleRate(0), m_currentCenterFrequency(0), m_blockSize(blockSize > 0 ? blockSize : 1), m_pos(0)
    { b200dsp_cxx::check(b200dsp_upchan_create(&m_h)); }
    virtual ~UpChannelizer() { b200dsp_upchan_destroy(m_h); }
    UpChannelizer(const UpChannelizer&) = delete;
    UpChannelizer& operator=(const UpChannelizer&) = delete;
    /** == handleMessage(DSPSignalNotification(sampleRate, ...)) (upchannelizer.cpp:130-146) */
    void setOutputSampleRate(int sampleRate) { m_outputSampleRate = sampleRate; applyConfiguration(); }
    /** == configure(messageQueue, sampleRate, centerFrequency) -> DSPConfigureChannelizer (upchannelizer.cpp:45-49,147-160) */
    void configure(int sampleRate, int centerFrequency)
    {
        m_requestedInputSampleRate = sampleRate;
        m_requestedCenterFrequency = centerFrequency;
        applyConfiguration();
    }
    int getOutputSampleRate() const { return m_outputSampleRate; }
    /** what MsgChannelizerNotification carries (upchannelizer.cpp:204-208) */
    int getCurrentInputSampleRate() const { return m_currentInputSampleRate; }
    int getCurrentCenterFrequency() const { return m_currentCenterFrequency; }

    virtual void start() { if (m_sampleSource) m_sampleSource->start(); }
    virtual void stop() { if (m_sampleSource) m_sampleSource->stop(); }
    virtual void pullAudio(int nbSamples) { if (m_sampleSource) m_sampleSource->pullAudio(nbSamples); }
    virtual void pull(Sample& sample)
    {
        if (!m_sampleSource) return;                                     // (:53-56)
        if (m_pos >= m_block.size()) { m_block.resize((size_t) m_blockSize); pullBlock(&m_block[0], m_blockSize); m_pos = 0; }
        sample = m_block[m_pos++];
    }
    virtual void pullBlock(Sample* samples, int nbSamples)
    {
        if (!m_sampleSource || nbSamples <= 0) return;
        const int64_t need = b200dsp_upchan_source_count(m_h, nbSamples);
        m_source.resize((size_t) (need > 0 ? need : 1));
        if (need > 0) m_sampleSource->pullBlock(&m_source[0], (int) need);
        b200dsp_cxx::check(b200dsp_upchan_pull(m_h, (const int16_t*) &m_source[0], need, (int16_t*) samples, nbSamples));
    }
protected:
    void applyConfiguration()
    {
        if (m_outputSampleRate == 0) return;                             // (:177-185)
        m_block.clear(); m_pos = 0;
        if (m_requestedInputSampleRate <= 0) {
            // not configured by the modulator yet: an empty channel fits no half, the chain has no stage (signalContainsChannel, :240-250)
            b200dsp_cxx::check(b200dsp_upchan_set_path(m_h, nullptr, 0));
            m_currentInputSampleRate = m_outputSampleRate; m_currentCenterFrequency = m_requestedCenterFrequency;
            return;
        }
        b200dsp_cxx::check(b200dsp_upchan_configure(m_h, m_outputSampleRate, m_requestedInputSampleRate, m_requestedCenterFrequency,
                                                    &m_currentInputSampleRate, &m_currentCenterFrequency));
    }
    BasebandSampleSource* m_sampleSource;
    b200dsp_upchan_t* m_h;
    int m_outputSampleRate, m_requestedInputSampleRate, m_requestedCenterFrequency, m_currentInputSampleRate, m_currentCenterFrequency;
    int m_blockSize;
    size_t m_pos;
    SampleVector m_block, m_source;
};
#endif
