// samplesinkfifo.h — Qt-free SampleSinkFifo (sdrbase/dsp/samplesinkfifo.h:27-65, samplesinkfifo.cpp:113-231): the ring between a
// device plugin's decimators (producer thread) and the engine's work loop (consumer).  Same methods and semantics:
// write() takes min(count, size - fill) samples and drops the rest, read()/readBegin() hand out min(count, fill), readBegin
// returns at most two contiguous spans and consumes nothing until readCommit.  The Qt signal dataReady() becomes an
// optional callback.  DeviceSampleSinkFifo is the same ring in device memory (b200dsp_fifo_*): decimator output goes in with
// a device-to-device copy and a bank is fed from the spans, so the samples never visit the host.
#ifndef SDRANGEL_B200_DSP_SAMPLESINKFIFO_H
#define SDRANGEL_B200_DSP_SAMPLESINKFIFO_H
#include <algorithm>
#include <functional>
#include <mutex>
#include "dsptypes.h"

class SampleSinkFifo {
public:
    SampleSinkFifo() : m_size(0), m_fill(0), m_head(0), m_tail(0) {}
    explicit SampleSinkFifo(int size) : m_size(0), m_fill(0), m_head(0), m_tail(0) { create((uint) size); }
    bool setSize(int size) { create((uint) size); return m_data.size() == (size_t) size; }
    uint size() const { return m_size; }
    uint fill() { std::lock_guard<std::mutex> g(m_mutex); return m_fill; }
    void setDataReadyCallback(std::function<void()> cb) { m_dataReady = cb; }

    /** raw bytes of Samples (samplesinkfifo.cpp:70-111) */
    uint write(const quint8* data, uint count)
    {
        const Sample* begin = reinterpret_cast<const Sample*>(data);
        return writeSamples(begin, count / (uint) sizeof(Sample));
    }
    uint write(SampleVector::const_iterator begin, SampleVector::const_iterator end)
    {
        const uint count = (uint) (end - begin);
        return writeSamples(count ? &(*begin) : nullptr, count);
    }
    uint read(SampleVector::iterator begin, SampleVector::iterator end)
    {
        std::lock_guard<std::mutex> g(m_mutex);
        const uint count = (uint) (end - begin);
        const uint total = std::min(count, m_fill);              // underflow: what there is
        uint remaining = total;
        while (remaining > 0) {
            const uint len = std::min(remaining, m_size - m_head);
            std::copy(m_data.begin() + m_head, m_data.begin() + m_head + len, begin);
            m_head = (m_head + len) % m_size;
            m_fill -= len;
            begin += len;
            remaining -= len;
        }
        return total;
    }
    uint readBegin(uint count, SampleVector::iterator* part1Begin, SampleVector::iterator* part1End,
                   SampleVector::iterator* part2Begin, SampleVector::iterator* part2End)
    {
        std::lock_guard<std::mutex> g(m_mutex);
        uint head = m_head;
        const uint total = std::min(count, m_fill);
        uint remaining = total;
        if (remaining > 0) {
            const uint len = std::min(remaining, m_size - head);
            *part1Begin = m_data.begin() + head;
            *part1End = m_data.begin() + head + len;
            head = (head + len) % m_size;
            remaining -= len;
        } else {
            *part1Begin = m_data.end();
            *part1End = m_data.end();
        }
        if (remaining > 0) {
            const uint len = std::min(remaining, m_size - head);
            *part2Begin = m_data.begin() + head;
            *part2End = m_data.begin() + head + len;
        } else {
            *part2Begin = m_data.end();
            *part2End = m_data.end();
        }
        return total;
    }
    uint readCommit(uint count)
    {
        std::lock_guard<std::mutex> g(m_mutex);
        if (count > m_fill) count = m_fill;
        m_head = (m_head + count) % m_size;
        m_fill -= count;
        return count;
    }
private:
    void create(uint s)
    {
        m_size = 0; m_fill = 0; m_head = 0; m_tail = 0;
        m_data.resize(s);
        m_size = (uint) m_data.size();
    }
    uint writeSamples(const Sample* begin, uint count)
    {
        uint total;
        {
            std::lock_guard<std::mutex> g(m_mutex);
            total = std::min(count, m_size - m_fill);            // overflow: the excess is dropped (samplesinkfifo.cpp:122-137)
            uint remaining = total;
            while (remaining > 0) {
                const uint len = std::min(remaining, m_size - m_tail);
                std::copy(begin, begin + len, m_data.begin() + m_tail);
                m_tail = (m_tail + len) % m_size;
                m_fill += len;
                begin += len;
                remaining -= len;
            }
        }
        if (m_fill > 0 && m_dataReady) m_dataReady();
        return total;
    }
    std::mutex m_mutex;
    SampleVector m_data;
    uint m_size, m_fill, m_head, m_tail;
    std::function<void()> m_dataReady;
};

/** The same ring in device memory.  Spans are device pointers to packed int16 I/Q (4 bytes per sample). */
class DeviceSampleSinkFifo {
public:
    explicit DeviceSampleSinkFifo(uint size) : m_h(nullptr) { b200dsp_cxx::check(b200dsp_fifo_create(&m_h, size)); }
    ~DeviceSampleSinkFifo() { b200dsp_fifo_destroy(m_h); }
    uint size() const { return b200dsp_fifo_size(m_h); }
    uint fill() { return b200dsp_fifo_fill(m_h); }
    /** from host memory (a device plugin's callback buffer after its decimators ran on the host side of the C ABI) */
    uint write(SampleVector::const_iterator begin, SampleVector::const_iterator end, void* cudaStream = nullptr)
    {
        uint32_t w = 0;
        const uint count = (uint) (end - begin);
        b200dsp_cxx::check(b200dsp_fifo_write(m_h, count ? &(*begin) : nullptr, count, 0, cudaStream, &w));
        return w;
    }
    /** from device memory: the output of b200dsp_decim_run_dev */
    uint writeDevice(const void* dSamples, uint count, void* cudaStream = nullptr)
    {
        uint32_t w = 0;
        b200dsp_cxx::check(b200dsp_fifo_write(m_h, dSamples, count, 1, cudaStream, &w));
        return w;
    }
    uint readBegin(uint count, const void** part1, uint* n1, const void** part2, uint* n2)
    {
        uint32_t total = 0, a = 0, b = 0;
        b200dsp_cxx::check(b200dsp_fifo_read_begin(m_h, count, part1, &a, part2, &b, &total));
        *n1 = a; *n2 = b;
        return total;
    }
    uint readCommit(uint count) { uint32_t c = 0; b200dsp_cxx::check(b200dsp_fifo_read_commit(m_h, count, &c)); return c; }
    uint read(SampleVector::iterator begin, SampleVector::iterator end, void* cudaStream = nullptr)
    {
        uint32_t r = 0;
        const uint count = (uint) (end - begin);
        b200dsp_cxx::check(b200dsp_fifo_read(m_h, count ? &(*begin) : nullptr, count, cudaStream, &r));
        return r;
    }
    b200dsp_fifo_t* handle() { return m_h; }
private:
    DeviceSampleSinkFifo(const DeviceSampleSinkFifo&);
    DeviceSampleSinkFifo& operator=(const DeviceSampleSinkFifo&);
    b200dsp_fifo_t* m_h;
};
#endif
