// dsptypes.h — Qt-free restatement of the reference's sample containers (sdrbase/dsp/dsptypes.h:31-96), 16-bit Rx mode.
// Same names, same layout (Sample = 2 x int16 packed, FSample = 2 x float), so SampleVector storage can be handed to
// the C ABI as interleaved I/Q without a copy.
#ifndef SDRANGEL_B200_DSP_DSPTYPES_H
#define SDRANGEL_B200_DSP_DSPTYPES_H
#include <complex>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../b200dsp.h"

typedef int8_t qint8;   typedef uint8_t quint8;
typedef int16_t qint16; typedef uint16_t quint16;
typedef int32_t qint32; typedef uint32_t quint32;
typedef int64_t qint64; typedef uint64_t quint64;
typedef unsigned int uint;

#define SDR_RX_SAMP_SZ 16
#define SDR_RX_SCALEF 32768.0f
#define SDR_TX_SAMP_SZ 16       // sdrbase/dsp/dsptypes.h:37
#define SDR_RX_SCALED 32768.0
typedef qint16 FixReal;
typedef float Real;
typedef std::complex<Real> Complex;

#pragma pack(push, 1)
struct Sample {
    Sample() : m_real(0), m_imag(0) {}
    Sample(FixReal real) : m_real(real), m_imag(0) {}
    Sample(FixReal real, FixReal imag) : m_real(real), m_imag(imag) {}
    inline void setReal(FixReal v) { m_real = v; }
    inline void setImag(FixReal v) { m_imag = v; }
    inline FixReal real() const { return m_real; }
    inline FixReal imag() const { return m_imag; }
    FixReal m_real;
    FixReal m_imag;
};
struct FSample {
    FSample() : m_real(0), m_imag(0) {}
    FSample(Real real) : m_real(real), m_imag(0) {}
    FSample(Real real, Real imag) : m_real(real), m_imag(imag) {}
    inline void setReal(Real v) { m_real = v; }
    inline void setImag(Real v) { m_imag = v; }
    inline Real real() const { return m_real; }
    inline Real imag() const { return m_imag; }
    Real m_real;
    Real m_imag;
};
#pragma pack(pop)
static_assert(sizeof(Sample) == 4 && sizeof(FSample) == 8, "sample layout must match the C ABI's interleaved I/Q");

typedef std::vector<Sample> SampleVector;
typedef std::vector<FSample> FSampleVector;

namespace b200dsp_cxx {
// The reference's DSP calls return void and cannot fail; the GPU layer can (no device, out of memory).  The wrappers
// turn a failing C-ABI call into an exception carrying b200dsp_last_error(): loud, never a silent CPU fallback.
inline void check(int rc)
{
    if (rc != 0) throw std::runtime_error(std::string("b200dsp: ") + b200dsp_last_error());
}
} // namespace b200dsp_cxx
#endif
