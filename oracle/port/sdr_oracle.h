/* ORACLE (test infrastructure; never shipped, never on the product path).
 *
 * Plain-C restatement of the reference's baseband-to-channel DSP hot path, written from the closed
 * forms in SURVEY.md Appendix A.  Each function cites the reference file:line it follows (paths are
 * relative to /root/reference).  Pinned by tests/ against (a) oracle/_ref/libsdrref.so, the unmodified
 * reference compiled in place, and (b) the golden vectors in tests/golden/ that were generated from it.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#ifndef SDR_ORACLE_H
#define SDR_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MODE_INF = 0, ORC_MODE_SUP = 1, ORC_MODE_CEN = 2 };
enum { ORC_FMT_I16 = 0, ORC_FMT_F32 = 1 };

/* sdrbench input generators: sdrbench/mainbench.cpp:30-31,76-79 (mt19937 default seed, 2N-1 draws, last = 0) */
void orc_sdrbench_gen_s16(int16_t* buf, int n_scalars);
void orc_sdrbench_gen_f32(float* buf, int n_scalars);

/* Decimators<qint32,qint16,16,InputBits>: sdrbase/dsp/decimators.h:279-341 and the entry points below it */
void* orc_decim_ii_create(int input_bits);
void  orc_decim_ii_destroy(void* h);
int   orc_decim_ii_run(void* h, int log2, int mode, const int16_t* buf, int len, int16_t* out);

/* 8-bit inputs: Decimators<qint32,qint8,16,8> (is_unsigned 0) / DecimatorsU<qint32,quint8,16,8,shift> (is_unsigned 1):
 * hackrfinputthread.h:57, rtlsdrthread.h:55, sdrbase/dsp/decimatorsu.h:175-249 */
void* orc_decim_x8_create(int is_unsigned, int shift);
void  orc_decim_x8_destroy(void* h);
int   orc_decim_x8_run(void* h, int log2, int mode, const uint8_t* buf, int len, int16_t* out);

/* DSPDeviceSourceEngine::iqCorrections(begin, end, false): sdrbase/dsp/dspdevicesourceengine.cpp:175-183,254-261 */
void* orc_iqcorr_create(void);
void  orc_iqcorr_destroy(void* h);
void  orc_iqcorr_dc(void* h, int16_t* iq, int n_samples);
/* ... (begin, end, true), floating-point flavour: dspdevicesourceengine.cpp:219-252 */
void  orc_iqcorr_imbalance(void* h, int16_t* iq, int n_samples);

/* DecimatorsFI / DecimatorsFF / DecimatorsIF: sdrbase/dsp/decimatorsfi.cpp, decimatorsff.cpp, decimatorsif.h */
void* orc_decim_f_create(int in_fmt, int out_fmt, int input_bits);
void  orc_decim_f_destroy(void* h);
int   orc_decim_f_run(void* h, int log2, int mode, const void* buf, int len, void* out);

/* DownChannelizer: sdrbase/dsp/downchannelizer.cpp:50-91,165-189,250-287 */
void* orc_chan_create(void);
void  orc_chan_destroy(void* h);
int   orc_chan_configure(void* h, int input_rate, int requested_rate, int center_offset,
                         int* out_rate, int* residual_offset, int* modes, int modes_cap);
int   orc_chan_feed(void* h, const int16_t* iq, int n, int16_t* out, int cap);

/* NCO + Interpolator as wired by a channel plugin: nco.cpp:30-64, interpolator.cpp:21-129, interpolator.h:23-36,
 * plugins/channelrx/demodnfm/nfmdemod.cpp:152-155,315,462-470 */
void  orc_nco_table(float* table4096);
int   orc_nco_increment(float freq, float rate);
int   orc_interp_ntaps(int phase_steps, double taps_per_phase);
void  orc_interp_taps(int phase_steps, double rate, double cutoff, double taps_per_phase, float* taps);
void* orc_frontend_create(float nco_freq, float nco_rate, int phase_steps, double interp_rate, double cutoff,
                          double taps_per_phase, float distance);
void  orc_frontend_destroy(void* h);
int   orc_frontend_feed(void* h, const int16_t* iq, int n, float* out, int cap, int32_t* idx, int32_t* phase);
/* Interpolator::decimate (mode 0) / interpolate (1) / resample (2) on complex float input in their callers' loops */
int   orc_interp_run(void* h, int mode, const float* cin, int n, float* out, int cap);
float orc_frontend_remain(void* h);
void  orc_nco_block(float freq, float rate, int n, float* out);

/* FFTWindow / KissFFT / SpectrumVis: fftwindow.cpp:20-73, kissfft.h:44-80,127-238, sdrgui/dsp/spectrumvis.cpp:77-327 */
void  orc_fft_window(int function, int n, float* w);
void  orc_kissfft_forward(int n, const float* in, float* out);
void* orc_spectrum_create(float scalef);
void  orc_spectrum_destroy(void* h);
void  orc_spectrum_configure(void* h, int fft_size, int overlap_pct, unsigned avg_nb, int avg_mode, int window, int linear);
int   orc_spectrum_feed(void* h, const int16_t* iq, int n, int positive_only, float* frames, int cap_frames);

/* Tx mirror.  Interpolators<T,16,OutputBits>: sdrbase/dsp/interpolators.h:104-617 over
 * IntHalfbandFilterEO1<>::myInterpolate (inthalfbandfiltereo1.h:601-622); buf int16 (12/16 bits) or int8 (8 bits) */
void* orc_interps_create(int output_bits);
void  orc_interps_destroy(void* h);
int   orc_interps_run(void* h, int log2, const int16_t* iq, void* buf, int len);
/* UpChannelizer: sdrbase/dsp/upchannelizer.cpp:51-104,175-209,252-327, IntHalfbandFilterEO1<96>::workInterpolate* */
void* orc_upchan_create(void);
void  orc_upchan_destroy(void* h);
int   orc_upchan_configure(void* h, int output_rate, int requested_rate, int center_offset, int* in_rate, int* residual_offset,
                           int* modes, int modes_cap);
int   orc_upchan_pull(void* h, const int16_t* iq, int n_in, int16_t* out, int n_out);
void  orc_hb_interp_coeffs(int order, int32_t* out);

/* SURVEY.md 8f-4.  PhaseDiscriminators: sdrbase/dsp/phasediscri.h:26-198; AM magnitude: plugins/channelrx/demodam/amdemod.cpp:154-156,241.
 * kind 0 phaseDiscriminator, 1 phaseDiscriminatorDelta (aux0 magsq, aux1 fmDev), 2 phaseDiscriminator2, 3 AM magnitude (aux0 magsq) */
void* orc_discri_create(float fm_scaling);
void  orc_discri_destroy(void* h);
void  orc_discri_run(void* h, int kind, const float* in_c64, int n, float* out, float* aux0, float* aux1);
/* fftfilt: sdrbase/dsp/fftfilt.cpp:49-360 (kind 0 fftfilt(f1,f2,len), 1 fftfilt(f2,len); op 0 runFilt, 1 runSSB, 2 runDSB; flag bit 0 usb, bit 1 getDC) */
void* orc_fftfilt_create(int kind, float f1, float f2, int len);
void  orc_fftfilt_destroy(void* h);
void  orc_fftfilt_set(void* h, int kind, float f1, float f2);
void  orc_fftfilt_filter(void* h, float* out_c64);
int   orc_fftfilt_run(void* h, int op, int flag, const float* in_c64, int n, float* out_c64, int cap);
/* .sdriq header: FileRecord::writeHeader, sdrbase/dsp/filerecord.cpp:129-137 */
void  orc_sdriq_header(int32_t rate, uint64_t center, int64_t ts, uint32_t sample_size, uint8_t* out24);

#ifdef __cplusplus
}
#endif
#endif
