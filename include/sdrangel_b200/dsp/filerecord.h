// filerecord.h — drop-in for FileRecord (sdrbase/dsp/filerecord.h:13-55, filerecord.cpp:25-148): the BasebandSampleSink that
// writes .sdriq record files (24-byte header + raw Samples), and the static readHeader the file-source plugin uses
// (plugins/samplesource/filesource/filesourceinput.cpp).  Host-side I/O over b200dsp_sdriq_*; the DSPSignalNotification the
// reference receives as a message arrives through setSampleRateAndFrequency().
#ifndef SDRANGEL_B200_DSP_FILERECORD_H
#define SDRANGEL_B200_DSP_FILERECORD_H
#include <ctime>
#include <fstream>
#include <string>
#include "basebandsamplesink.h"

class FileRecord : public BasebandSampleSink {
public:
    struct Header {
        qint32      sampleRate;
        quint64     centerFrequency;
        std::time_t startTimeStamp;
        quint32     sampleSize;
    };
    FileRecord() : m_fileName("test.sdriq"), m_sampleRate(0), m_centerFrequency(0), m_recordOn(false), m_w(nullptr), m_byteCount(0) {}
    explicit FileRecord(const std::string& filename) : m_fileName(filename), m_sampleRate(0), m_centerFrequency(0), m_recordOn(false), m_w(nullptr), m_byteCount(0) {}
    virtual ~FileRecord() { stopRecording(); }
    FileRecord(const FileRecord&) = delete;
    FileRecord& operator=(const FileRecord&) = delete;
    quint64 getByteCount() const { return m_byteCount; }             // counts Samples, like the reference (filerecord.cpp:89)
    void setFileName(const std::string& filename) { if (!m_recordOn) m_fileName = filename; }
    /** == handleMessage(DSPSignalNotification(sampleRate, centerFrequency)) (filerecord.cpp:117-127) */
    void setSampleRateAndFrequency(int sampleRate, quint64 centerFrequency) { m_sampleRate = sampleRate; m_centerFrequency = centerFrequency; }
    virtual void feed(const SampleVector::const_iterator& begin, const SampleVector::const_iterator& end, bool)
    {
        if (!m_recordOn) return;                                     // "send the samples to /dev/null"
        if (begin < end) {
            b200dsp_cxx::check(b200dsp_sdriq_write(m_w, (const int16_t*) &(*begin), end - begin));    // the header goes out with the first samples
            m_byteCount += end - begin;
        }
    }
    virtual void start() {}
    virtual void stop() { stopRecording(); }
    void startRecording()
    {
        if (m_w) return;
        b200dsp_cxx::check(b200dsp_sdriq_create(&m_w, m_fileName.c_str(), m_sampleRate, m_centerFrequency, (int64_t) time(0)));
        m_recordOn = true;
        m_byteCount = 0;
    }
    void stopRecording()
    {
        if (!m_w) return;
        b200dsp_sdriq_close(m_w);
        m_w = nullptr;
        m_recordOn = false;
    }
    static void readHeader(std::ifstream& sampleFile, Header& header)
    {
        char raw[B200DSP_SDRIQ_HEADER_BYTES] = { 0 };
        sampleFile.read(raw, sizeof(raw));
        int32_t r = 0; uint64_t c = 0; int64_t t = 0; uint32_t s = 0;
        b200dsp_sdriq_header_decode(raw, &r, &c, &t, &s);
        header.sampleRate = r; header.centerFrequency = c; header.startTimeStamp = (std::time_t) t; header.sampleSize = s;
    }
private:
    std::string m_fileName;
    qint32 m_sampleRate;
    quint64 m_centerFrequency;
    bool m_recordOn;
    b200dsp_sdriq_t* m_w;
    quint64 m_byteCount;
};
#endif
