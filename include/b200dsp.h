/* b200dsp.h — C ABI of the B200-native SDRangel baseband-to-channel DSP hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, opaque handles, no C++/torch/Qt types.
 * The reference has no C plugin ABI (its plugins instantiate the DSP classes by value), so every
 * entry point below names the reference class/method it replaces (paths relative to the reference
 * tree); include/sdrangel_b200/dsp/ holds header-only C++ wrappers with the reference's class names and
 * method signatures that forward here.
 *
 * Conventions
 *   - every function returns 0 on success, a negative B200DSP_E* code on failure; b200dsp_last_error()
 *     returns a thread-local message for the last failure on the calling thread.
 *   - one handle == one reference object; a handle is single-writer (like the reference objects) and
 *     owns a CUDA stream; different handles may be driven from different host threads concurrently.
 *   - host-pointer calls return when the outputs are host-visible; `_dev` calls take device pointers
 *     (16-byte aligned) and are asynchronous on the given stream (NULL = the handle's stream).
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with B200DSP_ENODEV.
 */
#ifndef B200DSP_H
#define B200DSP_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DSP_OK        0
#define B200DSP_EINVAL   -1   /* bad argument */
#define B200DSP_ENODEV   -2   /* no CUDA device / driver */
#define B200DSP_ECUDA    -3   /* CUDA runtime error (message in last_error) */
#define B200DSP_ENOMEM   -4
#define B200DSP_ESTATE   -5   /* call not valid in this state (e.g. not configured) */

/* sample formats */
#define B200DSP_FMT_I16   0   /* int16 interleaved I,Q  (reference: qint16 buffers / Sample, dsptypes.h:44-65) */
#define B200DSP_FMT_F32   1   /* float interleaved I,Q  (reference: FSample, dsptypes.h:67-93) */
#define B200DSP_FMT_I8    2   /* int8 interleaved I,Q   (reference: qint8 device buffers, hackrfinputthread.h:57) */
#define B200DSP_FMT_U8    3   /* uint8 interleaved I,Q  (reference: quint8 device buffers, rtlsdrthread.h:55) */
/* fc position, same meaning as the reference's m_fcPos switch (plugins/samplesource/airspy/airspythread.cpp:118-202) */
#define B200DSP_MODE_INF  0
#define B200DSP_MODE_SUP  1
#define B200DSP_MODE_CEN  2
#define B200DSP_MODE_U    3   /* Decimators<>::decimate2_u (decimators.h:374-393): unfiltered /2 of offset-255 data; log2_decim 1, signed integer inputs */

/* ---- library ------------------------------------------------------------------------------------------- */
int         b200dsp_init(int device_ordinal);      /* selects the device for handles created afterwards */
int         b200dsp_device_count(void);            /* 0 when no usable CUDA device */
const char* b200dsp_last_error(void);
const char* b200dsp_version(void);
int         b200dsp_sm_count(void);

/* ---- K1/K2: half-band decimation cascades ---------------------------------------------------------------
 * One handle == one reference decimator object with its six half-band stages of state:
 *   in I16, out I16 : Decimators<qint32,qint16,16,input_bits>   sdrbase/dsp/decimators.h:277-341
 *   in F32, out I16 : DecimatorsFI                               sdrbase/dsp/decimatorsfi.h:26-55
 *   in F32, out F32 : DecimatorsFF                               sdrbase/dsp/decimatorsff.h
 *   in I16, out F32 : DecimatorsIF<qint16,input_bits>            sdrbase/dsp/decimatorsif.h:53-83
 *   in I8,  out I16 : Decimators<qint32,qint8,16,8>              plugins/samplesource/hackrfinput/hackrfinputthread.h:57
 *   in U8,  out I16 : DecimatorsU<qint32,quint8,16,8,Shift>      sdrbase/dsp/decimatorsu.h:175-215 (Shift 127: rtlsdrthread.h:55)
 * input_bits in {8,12,16} selects decimation_shifts<16,input_bits> (decimators.h:79-167) / the IF scale.
 */
typedef struct b200dsp_decim b200dsp_decim_t;

int b200dsp_decim_create(b200dsp_decim_t** h, int in_fmt, int out_fmt, int input_bits);
int b200dsp_decim_destroy(b200dsp_decim_t* h);

/* DecimatorsU's Shift template argument (sample = byte - Shift, decimatorsu.h:225-226,241-249); U8 handles only, default 127 */
int b200dsp_decim_set_shift(b200dsp_decim_t* h, int shift);

/* float arithmetic flavour: 0 (default) = fused multiply-add accumulation (within 1e-5 rel. RMS of the
 * reference), 1 = separately rounded add/mul/add in the reference's order (bit-identical to the reference
 * compiled without -ffast-math).  No effect on the integer cascade, which is always bit-exact. */
int b200dsp_decim_set_exact_float(b200dsp_decim_t* h, int exact);

/* == decimate{1,2,4,...,64}_{inf,sup,cen}(SampleVector::iterator* it, const T* buf, qint32 len)
 *    (decimators.h:344-3886, decimatorsfi.cpp:19-1172): `len` counts scalars (2 per IQ sample); the trailing
 *    partial block is dropped exactly as the reference's loops do; filter state carries to the next call.
 *    out receives *n_out IQ samples (2 scalars each).  log2_decim 0..6. */
int b200dsp_decim_run(b200dsp_decim_t* h, int log2_decim, int mode,
                      const void* in, int32_t len_scalars, void* out, int32_t* n_out);
int b200dsp_decim_run_dev(b200dsp_decim_t* h, int log2_decim, int mode,
                          const void* d_in, int64_t len_scalars, void* d_out, int64_t* n_out, void* cuda_stream);
/* == the overloads on separate I and Q arrays, decimate1 / decimate2_u / decimateN_cen(it, bufI, bufQ, len)
 *    (decimators.h:359-371,395-417,2638-2700,...): len_per_array samples in each array; mode B200DSP_MODE_CEN or B200DSP_MODE_U */
int b200dsp_decim_run_split(b200dsp_decim_t* h, int log2_decim, int mode,
                            const void* in_i, const void* in_q, int32_t len_per_array, void* out, int32_t* n_out);
/* number of output IQ samples decimate*_ would write for `len_scalars` (pure host arithmetic, no device) */
int64_t b200dsp_decim_out_count(int in_fmt, int out_fmt, int log2_decim, int mode, int64_t len_scalars);

/* Filter state = for each of the six stages the last 64 stage inputs (I then Q, oldest first), as the
 * reference keeps them in IntHalfbandFilterEO::m_even/m_odd (inthalfbandfiltereo.h:743-749).
 * 6*2*64 elements of int32 (integer cascade) or float (float cascades). */
#define B200DSP_DECIM_STATE_ELEMS (6 * 2 * 64)
int b200dsp_decim_get_state(b200dsp_decim_t* h, void* state_host);
int b200dsp_decim_set_state(b200dsp_decim_t* h, const void* state_host);
int b200dsp_decim_reset(b200dsp_decim_t* h);
int b200dsp_decim_sync(b200dsp_decim_t* h);        /* wait for the handle's stream */

/* ---- K3/K4: DownChannelizer bank + channel plugin front-end -----------------------------------------------
 * A bank == many reference (DownChannelizer, NCO, Interpolator) triples fed from ONE baseband stream:
 *   DownChannelizer::feed / applyConfiguration / createFilterChain   sdrbase/dsp/downchannelizer.cpp:50-91,165-189,250-287
 *   the channel plugin front-end  c *= nco.nextIQ(); interpolator.decimate(...)
 *                                                                     plugins/channelrx/demodnfm/nfmdemod.cpp:150-155,315,453-476
 * Channels are added first (each add runs the reference's filter-chain selection and reports what
 * MsgChannelizerNotification would: output rate and residual offset), then samples are fed.  Adding a channel or a
 * front-end after feeding restarts all filter state (the plan is rebuilt).
 * The outputs of a feed stay valid until the next feed, like DownChannelizer's m_sampleBuffer (downchannelizer.cpp:86-89).
 */
typedef struct b200dsp_bank b200dsp_bank_t;

#define B200DSP_STAGE_CHANNELIZER 0   /* int16 IQ at input_rate / 2^S   (what DownChannelizer hands to its sink) */
#define B200DSP_STAGE_FRONTEND    1   /* float IQ after NCO mix + Interpolator::decimate (complex64) */

/* == DownChannelizer::applyConfiguration's filter-chain selection (downchannelizer.cpp:165-177,250-287) as pure host
 *    arithmetic (no device needed): returns the number of stages S, writes out rate, residual offset and the modes */
int b200dsp_filter_chain(int input_rate_hz, int requested_rate_hz, int center_offset_hz,
                         int* out_rate_hz, int* residual_offset_hz, int* modes, int cap);
int b200dsp_bank_create(b200dsp_bank_t** b, int input_rate_hz);
int b200dsp_bank_destroy(b200dsp_bank_t* b);
/* internal time-chunk (input samples per pass over the tree); default 12582912, rounded to a multiple of 768 */
int b200dsp_bank_set_chunk(b200dsp_bank_t* b, int64_t samples);
/* == DSPConfigureChannelizer(requested_rate, center_offset) -> MsgChannelizerNotification(out_rate, residual_offset) */
int b200dsp_bank_add_channel(b200dsp_bank_t* b, int requested_rate_hz, int center_offset_hz,
                             int* chan_id, int* out_rate_hz, int* residual_offset_hz);
/* a channel given by its filter stages (0 centre, 1 lower half, 2 upper half) instead of (rate, offset); its output is
 * trunc(stage output / 2^out_shift).  Building block for splitting one bank over several GPUs: a bank
 * fed with the output of a depth-k tree node holds the channels below it by their path suffix with out_shift = k + n_modes;
 * out_shift = 0 exposes a node's raw stage output. */
int b200dsp_bank_add_channel_path(b200dsp_bank_t* b, const int* modes, int n_modes, int out_shift, int* chan_id);
/* zero all filter state (as freshly constructed reference objects) without rebuilding the plan; in stream order on
 * cuda_stream (NULL = the bank's own stream, the one the host-pointer calls use) */
int b200dsp_bank_reset(b200dsp_bank_t* b, void* cuda_stream);
/* filter stages chosen for the channel: 0 centre, 1 lower half, 2 upper half (downchannelizer.h:72-76); returns S */
int b200dsp_bank_channel_path(b200dsp_bank_t* b, int chan_id, int* modes, int cap);
/* number of distinct half-band stages (tree nodes) the bank evaluates for all its channels */
int b200dsp_bank_node_count(b200dsp_bank_t* b);
/* == m_nco.setFreq(nco_freq, rate); m_interpolator.create(phase_steps, rate, cutoff, taps_per_phase);
 *    m_interpolatorDistance = rate / out_rate  with rate = the channel's channelizer output rate (nfmdemod.cpp:462-470) */
int b200dsp_bank_set_frontend(b200dsp_bank_t* b, int chan_id, float nco_freq_hz, int phase_steps, double cutoff_hz,
                              double taps_per_phase, int out_rate_hz);
int b200dsp_bank_frontend_info(b200dsp_bank_t* b, int chan_id, int* nco_increment, int* taps_per_phase, float* taps, int taps_cap);
/* == DownChannelizer::feed(begin, end, positiveOnly) for every channel of the bank (+ the front-ends) */
int b200dsp_bank_feed(b200dsp_bank_t* b, const int16_t* iq, int64_t n_samples);
/* device-pointer form, asynchronous on cuda_stream (NULL = the bank's own stream).  A bank is single-writer: successive
 * feeds must be ordered on one stream (or by events); fetch / fetch_all synchronise only the stream they are given (fetch:
 * the bank's own), so after a feed_dev on a caller stream synchronise that stream, or pass it to fetch_all, first. */
int b200dsp_bank_feed_dev(b200dsp_bank_t* b, const void* d_iq, int64_t n_samples, void* cuda_stream);
/* outputs produced by the last feed for one channel; stage selects int16 IQ (4 bytes/sample) or complex64 (8 bytes) */
int b200dsp_bank_fetch(b200dsp_bank_t* b, int chan_id, int stage, void* out, int64_t cap_samples, int64_t* n_samples);
int b200dsp_bank_fetch_dev(b200dsp_bank_t* b, int chan_id, int stage, const void** d_ptr, int64_t* n_samples);
/* what Interpolator::decimate's float32 distance recurrence decided (interpolator.h:23-36, nfmdemod.cpp:155,315) in the
 * channel's last internal pass: for every front-end output the index (within that pass) of the channel sample that emitted
 * it and the polyphase phase.  For parity checks of the schedule itself; a feed no longer than the chunk is one pass. */
int b200dsp_bank_fetch_schedule(b200dsp_bank_t* b, int chan_id, int32_t* idx, int32_t* phase, int64_t cap, int64_t* n);
/* every channel's outputs of the last feed in one transfer (what DSPDeviceSourceEngine::work's loop over the channel sinks
 * hands out, dspdevicesourceengine.cpp:325-408): channel c's samples land at out + c * stride_samples (samples of 4 or 8
 * bytes by stage), counts[c] = its sample count (0 for a channel without that stage); waits for the stream (NULL: the
 * bank's own) before returning.  Pinned `out` makes it one DMA. */
int b200dsp_bank_fetch_all(b200dsp_bank_t* b, int stage, void* out, int64_t stride_samples, int64_t* counts, void* cuda_stream);
/* == one engine work cycle for the whole bank, host buffers in and out (DSPDeviceSourceEngine::work's FIFO read + loop over
 * the channel sinks, dspdevicesourceengine.cpp:325-408): the results of b200dsp_bank_feed(iq, n) followed by
 * b200dsp_bank_fetch_all(stage, out, stride, counts), with the host-to-device copy of pass p+1, the kernels of pass p and
 * the device-to-host copy of the output columns already complete overlapped on three streams.  Pin iq and out. */
int b200dsp_bank_process(b200dsp_bank_t* b, const int16_t* iq, int64_t n_samples, int stage,
                         void* out, int64_t stride_samples, int64_t* counts);
/* the bank's own CUDA stream (cudaStream_t), the one the host-pointer calls and NULL-stream _dev calls use */
void* b200dsp_bank_stream(b200dsp_bank_t* b);
/* the device half of fetch_all: gathers into a caller-owned device array [channel][stride_samples] and device counts,
 * asynchronously on the stream (for callers that overlap the device-to-host copy with the next block's kernels) */
int b200dsp_bank_gather_dev(b200dsp_bank_t* b, int stage, void* d_out, int64_t stride_samples, int64_t* d_counts, void* cuda_stream);
/* device-to-device copy of samples [skip, skip + count) of a channel's channelizer output of the last feed */
int b200dsp_bank_copy_out_dev(b200dsp_bank_t* b, int chan_id, int64_t skip, int64_t count, void* d_dst, void* cuda_stream);
int b200dsp_bank_sync(b200dsp_bank_t* b);
/* device time (ms) and count of the tree-level kernel launches of the last internal pass (instrumentation for bench.py) */
int b200dsp_bank_tree_time(b200dsp_bank_t* b, float* ms, int* launches);

/* ---- K6: one bank's channels sharded over the GPUs of one box --------------------------------------------------------
 * No reference counterpart: the reference hands one SampleVector span to every channel sink of one process
 * (dspdevicesourceengine.cpp:325-408).  Here the channels are split by frequency block over N GPUs, one process (or host
 * thread) per GPU, each with an ordinary bank for its channels; every GPU needs the whole baseband:
 *   device-resident source: b200dsp_dist_bcast_begin  -- NCCL broadcast over NVLink from the GPU that holds it;
 *   host-fed source:        b200dsp_dist_ingest_begin -- every rank copies ITS 1/N time slice over its own PCIe link, an
 *                                                        in-place NCCL all-gather completes the block on every GPU.
 * Two receive slots: the transfer of block k+1 runs on the handle's collective stream under the kernels of block k.
 * NCCL (libnccl.so.2) is loaded at run time by these calls only. */
typedef struct b200dsp_dist b200dsp_dist_t;
#define B200DSP_DIST_ID_BYTES 128
/* channels [lo, hi) of n_channels served by `rank`: contiguous blocks in the order given (frequency order) */
int b200dsp_dist_shard(int n_channels, int world, int rank, int* lo, int* hi);
int b200dsp_dist_unique_id(void* id_out);                  /* one rank makes it, all ranks pass it to _create (ncclGetUniqueId) */
int b200dsp_dist_create(b200dsp_dist_t** d, const void* id, int rank, int world);       /* collective; device = b200dsp_init's */
int b200dsp_dist_destroy(b200dsp_dist_t* d);
int b200dsp_dist_reserve(b200dsp_dist_t* d, int64_t n_samples);                          /* two slots of n_samples (grows on demand) */
/* collective: block of n_samples int16 IQ from `root`'s device buffer d_iq (ignored elsewhere) into `slot` (0/1);
 * after_stream (root, may be NULL): a stream whose work so far produced d_iq */
int b200dsp_dist_bcast_begin(b200dsp_dist_t* d, int slot, const void* d_iq, int64_t n_samples, int root, void* after_stream);
/* collective: host_slice = this rank's samples [rank * n/world, (rank+1) * n/world) of the block (n_samples_total a multiple of 4 * world) */
int b200dsp_dist_ingest_begin(b200dsp_dist_t* d, int slot, const int16_t* host_slice, int64_t n_samples_total);
/* feed `bank` from the slot once its transfer has landed (stream order; NULL = the bank's stream); the slot may be refilled
 * by the next _begin, which waits for this feed */
int b200dsp_dist_feed(b200dsp_dist_t* d, int slot, b200dsp_bank_t* bank, void* cuda_stream);
int b200dsp_dist_slot(b200dsp_dist_t* d, int slot, const void** d_ptr, int64_t* n_samples);     /* the received block (for checks) */
int b200dsp_dist_sync(b200dsp_dist_t* d);
/* Copy-engine chain for the device-resident case (one process per GPU): the block travels rank 0 -> 1 -> ... -> N-1 in 16
 * sub-blocks, each rank forwarding by DMA copies into the next rank's receive slot (CUDA IPC mapping), ordered by counters
 * in device memory that the streams wait on / write (cuStreamWaitValue32 and 4-byte copies).  No SM is taken from the FIR
 * kernels and every NVLink port carries the block once in, once out.  Set-up: every rank _export()s (allocates 3 slots of
 * n_samples), the B200DSP_DIST_P2P_BLOB_BYTES blobs of all ranks are gathered by the caller's own means, in rank order, and
 * every rank _import()s the array.  Then per block, on every rank, in the same order: _p2p_begin(slot) ... _p2p_feed(slot). */
#define B200DSP_DIST_P2P_BLOB_BYTES 512
int b200dsp_dist_p2p_export(b200dsp_dist_t* d, int64_t n_samples, void* blob_out);
int b200dsp_dist_p2p_import(b200dsp_dist_t* d, const void* blobs_all);
int b200dsp_dist_p2p_begin(b200dsp_dist_t* d, int slot, const void* d_iq_rank0, int64_t n_samples, void* after_stream);
int b200dsp_dist_p2p_feed(b200dsp_dist_t* d, int slot, b200dsp_bank_t* bank, void* cuda_stream);
int b200dsp_dist_p2p_slot(b200dsp_dist_t* d, int slot, const void** d_ptr, int64_t* n_samples);

/* ---- engine-side sample corrections ----------------------------------------------------------------
 * == DSPDeviceSourceEngine::iqCorrections(begin, end, imbalanceCorrection)  sdrbase/dsp/dspdevicesourceengine.cpp:175-262,
 *    the step DSPDeviceSourceEngine::work applies between the device Decimators<> and the channel sinks (:343-347).
 * One handle == one engine's m_iBeta / m_qBeta state (MovingAverageUtil<int32_t,int64_t,1024>, dspdevicesourceengine.h:106-107).
 * imbalance 0: DC correction (:254-259), bit-exact.  imbalance 1: the I/Q imbalance branch (:219-252, the floating-point flavour
 * the reference compiles): segments of 2048 samples replayed in the reference's operation order after a 2048-sample
 * warm-up -- within 1 LSB of the reference (its running totals carry rounding drift from before the warm-up). */
typedef struct b200dsp_iqcorr b200dsp_iqcorr_t;
int b200dsp_iqcorr_create(b200dsp_iqcorr_t** h);
int b200dsp_iqcorr_destroy(b200dsp_iqcorr_t* h);
int b200dsp_iqcorr_reset(b200dsp_iqcorr_t* h);
int b200dsp_iqcorr_run(b200dsp_iqcorr_t* h, int16_t* iq, int64_t n_samples, int imbalance);      /* in place, host buffer */
int b200dsp_iqcorr_run_dev(b200dsp_iqcorr_t* h, const void* d_in, void* d_out, int64_t n_samples, int imbalance, void* cuda_stream);

/* ---- stand-alone Interpolator (the polyphase resampler of K4 without the bank) -------------------------------------
 * == Interpolator::create(phaseSteps, sampleRate, cutoff, nbTapsPerPhase)   sdrbase/dsp/interpolator.cpp:74-129
 * b200dsp_interp_decimate is the block form of the loop every Rx plugin writes (plugins/channelrx/demodnfm/nfmdemod.cpp:150-155,315):
 *     for each input c:  if (interp.decimate(&remain, c, &ci)) { out[m++] = ci; remain += distance; }
 * `distance_remain` is the caller-owned Real the reference passes by pointer: read on entry, updated on return.
 * Complex samples are interleaved float pairs.  At most 2^24-1 input samples per call. */
typedef struct b200dsp_interp b200dsp_interp_t;
int b200dsp_interp_create(b200dsp_interp_t** h, int phase_steps, double sample_rate, double cutoff, double taps_per_phase);
int b200dsp_interp_destroy(b200dsp_interp_t* h);
int b200dsp_interp_info(b200dsp_interp_t* h, int* taps_per_phase, float* taps, int taps_cap);
int b200dsp_interp_decimate(b200dsp_interp_t* h, float* distance_remain, float distance, const float* in_c64, int64_t n_samples,
                            float* out_c64, int64_t cap_samples, int64_t* n_out);
/* == Interpolator::interpolate (interpolator.h:39-52) in the loop every Tx plugin writes around it
 *    (plugins/channeltx/modnfm/nfmmod.cpp:126-133, modam/ammod.cpp:120-127, modssb/ssbmod.cpp:146-153):
 *      for each OUTPUT: if (interp.interpolate(&remain, x[i], &y)) ++i;  out[m++] = y;  remain += distance;
 *    until the next call would need x[n_samples].  All n_samples inputs are consumed; about n_samples / distance outputs. */
int b200dsp_interp_interpolate(b200dsp_interp_t* h, float* distance_remain, float distance, const float* in_c64, int64_t n_samples,
                               float* out_c64, int64_t cap_samples, int64_t* n_out);
/* == Interpolator::resample (interpolator.h:55-76), the arbitrary P/Q form, in its canonical loop:
 *      for each input c: consumed = false; do { if (interp.resample(&remain, c, &consumed, &y)) { out[m++] = y; remain += distance; } } while (!consumed); */
int b200dsp_interp_resample(b200dsp_interp_t* h, float* distance_remain, float distance, const float* in_c64, int64_t n_samples,
                            float* out_c64, int64_t cap_samples, int64_t* n_out);

/* one reference call (op 0 decimate, 1 interpolate, 2 resample) with the reference's exact effect on *distance; one launch
 * per sample -- for plugin code not yet restructured into blocks.  *produced: a result was written; *consumed: `next` was taken */
int b200dsp_interp_step(b200dsp_interp_t* h, int op, float* distance, const float* next_c64, float* result_c64, int* consumed, int* produced);

/* ---- device-resident SampleSinkFifo -------------------------------------------------------------------------------------
 * == SampleSinkFifo (sdrbase/dsp/samplesinkfifo.h:27-65, samplesinkfifo.cpp:113-231): the ring between a device plugin's
 *    decimators and the engine's work loop, same semantics (overflow drops, two-span readBegin + readCommit), ring in device
 *    memory: decimator output in with a device-to-device copy, b200dsp_bank_feed_dev straight from the spans. */
typedef struct b200dsp_fifo b200dsp_fifo_t;
int      b200dsp_fifo_create(b200dsp_fifo_t** f, uint32_t size_samples);
int      b200dsp_fifo_destroy(b200dsp_fifo_t* f);
uint32_t b200dsp_fifo_size(b200dsp_fifo_t* f);
uint32_t b200dsp_fifo_fill(b200dsp_fifo_t* f);
int      b200dsp_fifo_write(b200dsp_fifo_t* f, const void* samples, uint32_t count, int src_is_device, void* cuda_stream, uint32_t* written);
int      b200dsp_fifo_read_begin(b200dsp_fifo_t* f, uint32_t count, const void** d_part1, uint32_t* n1, const void** d_part2, uint32_t* n2, uint32_t* total);
int      b200dsp_fifo_read_commit(b200dsp_fifo_t* f, uint32_t count, uint32_t* committed);
int      b200dsp_fifo_read(b200dsp_fifo_t* f, void* out_host, uint32_t count, void* cuda_stream, uint32_t* read);

/* ---- stand-alone NCO -------------------------------------------------------------------------------------------------
 * == NCO (sdrbase/dsp/nco.h:40-53, nco.cpp:30-64): 4096-entry cosine table, integer phase advanced BEFORE each lookup,
 *    nextIQ() = (T[p], -T[(p + 1024) mod 4096]).  One handle == one NCO object (its phase). */
typedef struct b200dsp_nco b200dsp_nco_t;
int b200dsp_nco_create(b200dsp_nco_t** h);
int b200dsp_nco_destroy(b200dsp_nco_t* h);
int b200dsp_nco_set_freq(b200dsp_nco_t* h, float freq, float sample_rate);       /* == NCO::setFreq: increment = (int) (freq * 4096 / rate) */
int b200dsp_nco_set_phase(b200dsp_nco_t* h, int phase);                            /* == NCO::setPhase */
int b200dsp_nco_get(b200dsp_nco_t* h, int* phase, int* phase_increment);
/* n x nextIQ(): interleaved (re, im) floats; the phase advances by n increments */
int b200dsp_nco_next_iq(b200dsp_nco_t* h, int64_t n, float* out_c64);
int b200dsp_nco_next_iq_dev(b200dsp_nco_t* h, int64_t n, float* d_out_c64, void* cuda_stream);
/* the mix a channel plugin does per sample, as a block: out[i] = Complex(in[i].re, in[i].im) * nco.nextIQ()  (nfmdemod.cpp:152-153) */
int b200dsp_nco_mix_dev(b200dsp_nco_t* h, const void* d_in_i16, int64_t n, float* d_out_c64, void* cuda_stream);

/* ---- Tx mirror: device-side interpolators and the UpChannelizer -------------------------------------------
 * K8 == Interpolators<T, SDR_TX_SAMP_SZ 16, OutputBits> (sdrbase/dsp/interpolators.h:104-617): the step a sample-sink plugin
 *    runs between the engine's SampleVector and the device buffer,
 *      out I16, output_bits 16 : Interpolators<qint16,16,16>   plugins/samplesink/filesink/filesinkthread.h:73, plutosdroutputthread.h:57
 *      out I16, output_bits 12 : Interpolators<qint16,16,12>   plugins/samplesink/bladerfoutput/bladerfoutputthread.h:53, limesdroutputthread.h:57
 *      out I8,  output_bits 8  : Interpolators<qint8,16,8>     plugins/samplesink/hackrfoutput/hackrfoutputthread.h:52
 *    One handle == one Interpolators object: six interpolating half-bands (orders 64, 32, 16, 16, 16, 16;
 *    IntHalfbandFilterEO1<>::myInterpolate, inthalfbandfiltereo1.h:601-622) whose rings persist across calls and factors. */
typedef struct b200dsp_interps b200dsp_interps_t;
int b200dsp_interps_create(b200dsp_interps_t** h, int out_fmt, int output_bits);
int b200dsp_interps_destroy(b200dsp_interps_t* h);
int b200dsp_interps_reset(b200dsp_interps_t* h);
/* Samples interpolate*_ consumes for `len_scalars` output scalars: len / (2 << log2) (pure host arithmetic) */
int64_t b200dsp_interps_in_count(int log2_interp, int64_t len_scalars);
/* == interpolate{1,2,4,...,64}_cen(SampleVector::iterator* it, T* buf, qint32 len): `len` counts OUTPUT scalars; reads
 *    *n_consumed = len / (2 << log2) Samples (the iterator's advance), writes n_consumed * (2 << log2) scalars of `buf`;
 *    a trailing partial block is left alone like the reference's loops do.  interpolate64_cen writes only scalars 0..109 of
 *    every block of 128 (its store list stops there, interpolators.h): those 18 scalars of `buf` stay untouched here too. */
int b200dsp_interps_run(b200dsp_interps_t* h, int log2_interp, const int16_t* samples_iq, void* buf, int32_t len_scalars, int32_t* n_consumed);
int b200dsp_interps_run_dev(b200dsp_interps_t* h, int log2_interp, const void* d_samples_iq, void* d_buf, int64_t len_scalars,
                            int64_t* n_consumed, void* cuda_stream);

/* K9 == UpChannelizer (sdrbase/dsp/upchannelizer.cpp:51-104,175-209,252-327) with its IntHalfbandFilterEO1<96> stages
 *    (workInterpolateCenter / LowerHalf / UpperHalf, inthalfbandfiltereo1.h:98-127,291-355,490-554): the modulator's samples
 *    at rate out / 2^S interpolated to the device sink's rate and shifted to the channel's offset.  One handle == one object. */
typedef struct b200dsp_upchan b200dsp_upchan_t;
int b200dsp_upchan_create(b200dsp_upchan_t** h);
int b200dsp_upchan_destroy(b200dsp_upchan_t* h);
/* == DSPSignalNotification(output rate) + DSPConfigureChannelizer(requested rate, offset) -> applyConfiguration: the filter
 *    chain is rebuilt (fresh filters); reports what MsgChannelizerNotification carries: modulator rate and residual offset */
int b200dsp_upchan_configure(b200dsp_upchan_t* h, int output_rate_hz, int requested_rate_hz, int center_offset_hz,
                             int* in_rate_hz, int* residual_offset_hz);
/* the stage list given directly: 0 centre, 1 lower half, 2 upper half; stage 0 runs at the output rate (upchannelizer.h:83-104) */
int b200dsp_upchan_set_path(b200dsp_upchan_t* h, const int* modes, int n_modes);
int b200dsp_upchan_path(b200dsp_upchan_t* h, int* modes, int cap);            /* returns S */
/* modulator samples the next n_out calls of pull() take from m_sampleSource->pull (depends on the stages' phases; host arithmetic) */
int64_t b200dsp_upchan_source_count(b200dsp_upchan_t* h, int64_t n_out);
/* == n_out x UpChannelizer::pull(sample): source = the samples m_sampleSource->pull hands out during those calls, in order
 *    (at least b200dsp_upchan_source_count(n_out) of them; exactly that many are consumed) */
int b200dsp_upchan_pull(b200dsp_upchan_t* h, const int16_t* source_iq, int64_t n_source, int16_t* out_iq, int64_t n_out);
int b200dsp_upchan_pull_dev(b200dsp_upchan_t* h, const void* d_source_iq, int64_t n_source, void* d_out_iq, int64_t n_out, void* cuda_stream);

/* ---- demodulator back-ends after Interpolator::decimate ------------------------------------------------
 * == PhaseDiscriminators (sdrbase/dsp/phasediscri.h:26-198) and the magnitude lines of AMDemod::processOneSample
 *    (plugins/channelrx/demodam/amdemod.cpp:154-156,241), as block calls over the complex64 outputs of the front-end.
 * One handle == n_channels discriminator objects of one kind (their m_m1Sample / m_m2Sample / m_prevArg carried between calls).
 * Kinds 1-3 are bit-identical to the reference compiled without -ffast-math; kind 0 differs by the device's atan2f (<= 2 ulp).
 * Squelch, audio filters, AGC and the audio FIFO behind them are audio back-end, out of scope. */
typedef struct b200dsp_demod b200dsp_demod_t;
#define B200DSP_DEMOD_FM_ATAN2    0   /* phaseDiscriminator(sample)                       phasediscri.h:48-53  */
#define B200DSP_DEMOD_FM_DELTA    1   /* phaseDiscriminatorDelta(sample, magsq, fmDev)    phasediscri.h:59-77 (NFMDemod, nfmdemod.cpp:165): aux0 = magsq, aux1 = fmDev */
#define B200DSP_DEMOD_FM_DISCRI2  2   /* phaseDiscriminator2(sample)                      phasediscri.h:84-96  */
#define B200DSP_DEMOD_AM_MAG      3   /* re, im / SDR_RX_SCALEF; magsq; sqrt(magsq)       amdemod.cpp:154-156,241: aux0 = magsq */
int b200dsp_demod_create(b200dsp_demod_t** h, int kind, float fm_scaling, int n_channels);
int b200dsp_demod_destroy(b200dsp_demod_t* h);
int b200dsp_demod_reset(b200dsp_demod_t* h);                                   /* == PhaseDiscriminators::reset (+ m_prevArg = 0) */
int b200dsp_demod_set_fm_scaling(b200dsp_demod_t* h, float fm_scaling);        /* == setFMScaling */
/* one stream, host buffers (1-channel handles); aux0 / aux1 may be NULL */
int b200dsp_demod_run(b200dsp_demod_t* h, const float* in_c64, int64_t n_samples, float* out, float* aux0, float* aux1);
/* every channel of a pooled block: channel c's samples at d_pool + c * stride (complex64), d_counts[c] of them (device array) --
 * the layout b200dsp_bank_gather_dev(STAGE_FRONTEND) writes; outputs at d_out + c * out_stride */
int b200dsp_demod_run_pool_dev(b200dsp_demod_t* h, const void* d_pool_c64, int64_t stride_samples, const int64_t* d_counts, int n_channels,
                               float* d_out, int64_t out_stride, float* d_aux0, float* d_aux1, void* cuda_stream);

/* K11 == fftfilt (sdrbase/dsp/fftfilt.cpp:49-360, Fldigi's overlap-add FFT filter over g_fft): the SSB / DSB demodulators' channel
 *    filter, called per Interpolator::decimate output (plugins/channelrx/demodssb/ssbdemod.cpp:91-92,165-175; amdemod.cpp:199-207).
 *    kind 0 == fftfilt(f1, f2, len) / create_filter (band-pass, frequencies relative to the sample rate), kind 1 == fftfilt(f2, len) /
 *    create_dsb_filter.  len a power of two, 16..4096.  Block form of runFilt (op 0) / runSSB(usb, getDC) (op 1) / runDSB(getDC)
 *    (op 2): every len/2 input samples (counted across calls, like inptr) yield len/2 outputs.  Agreement with the reference is
 *    to float32 rounding (another FFT factorisation than g_fft). */
typedef struct b200dsp_fftfilt b200dsp_fftfilt_t;
int b200dsp_fftfilt_create(b200dsp_fftfilt_t** h, int kind, float f1, float f2, int len);
int b200dsp_fftfilt_destroy(b200dsp_fftfilt_t* h);
int b200dsp_fftfilt_set_filter(b200dsp_fftfilt_t* h, int kind, float f1, float f2);
int b200dsp_fftfilt_filter(b200dsp_fftfilt_t* h, float* out_c64, int cap_samples);            /* the frequency response; returns len */
int64_t b200dsp_fftfilt_out_count(b200dsp_fftfilt_t* h, int64_t n_samples);                      /* outputs the next run of n_samples yields */
int b200dsp_fftfilt_run(b200dsp_fftfilt_t* h, int op, int usb, int get_dc, const float* in_c64, int64_t n_samples, float* out_c64,
                        int64_t cap_samples, int64_t* n_out);
int b200dsp_fftfilt_run_dev(b200dsp_fftfilt_t* h, int op, int usb, int get_dc, const void* d_in_c64, int64_t n_samples, void* d_out_c64,
                            int64_t cap_samples, int64_t* n_out, void* cuda_stream);

/* ---- .sdriq record files: FileRecord / the file-source plugin's on-disk format -----------------------
 * == FileRecord::writeHeader / readHeader / feed (sdrbase/dsp/filerecord.cpp:72-148): 24-byte header (qint32 sample rate,
 *    quint64 centre frequency, 8-byte time_t start, quint32 sample size), then raw Samples.  Pure host code, no device. */
typedef struct b200dsp_sdriq b200dsp_sdriq_t;
#define B200DSP_SDRIQ_HEADER_BYTES 24
int b200dsp_sdriq_header_encode(int32_t sample_rate, uint64_t center_frequency, int64_t start_timestamp, uint32_t sample_size, void* out24);
int b200dsp_sdriq_header_decode(const void* in24, int32_t* sample_rate, uint64_t* center_frequency, int64_t* start_timestamp, uint32_t* sample_size);
/* reader (what FileSourceThread streams from): header fields + number of Samples in the file */
int b200dsp_sdriq_open(b200dsp_sdriq_t** r, const char* path, int32_t* sample_rate, uint64_t* center_frequency, int64_t* start_timestamp,
                       uint32_t* sample_size, int64_t* n_samples);
int b200dsp_sdriq_read(b200dsp_sdriq_t* r, int16_t* iq, int64_t cap_samples, int64_t* got);
/* writer == FileRecord::startRecording ... feed ... stopRecording: the header goes out with the first non-empty feed */
int b200dsp_sdriq_create(b200dsp_sdriq_t** w, const char* path, int32_t sample_rate, uint64_t center_frequency, int64_t start_timestamp);
int b200dsp_sdriq_write(b200dsp_sdriq_t* w, const int16_t* iq, int64_t n_samples);
int b200dsp_sdriq_close(b200dsp_sdriq_t* r);

/* ---- K5: SpectrumVis -------------------------------------------------------------------------------------
 * One handle == one reference SpectrumVis sink (sdrgui/dsp/spectrumvis.cpp:77-254,283-327) with its FFTWindow
 * (sdrbase/dsp/fftwindow.cpp:20-73), FFT engine (sdrbase/dsp/kissengine.cpp, kissfft.h) and per-bin averagers
 * (sdrbase/util/movingaverage2d.h, fixedaverage2d.h).  Every frame the reference would hand to
 * GLSpectrum::newSpectrum(m_powerSpectrum, fftSize) (spectrumvis.cpp:147,182,231) is returned as one row of fft_size floats.
 */
typedef struct b200dsp_spectrum b200dsp_spectrum_t;

#define B200DSP_AVG_NONE    0   /* SpectrumVis::AvgModeNone   */
#define B200DSP_AVG_MOVING  1   /* SpectrumVis::AvgModeMoving */
#define B200DSP_AVG_FIXED   2   /* SpectrumVis::AvgModeFixed  */
/* window: FFTWindow::Function 0 Bartlett, 1 BlackmanHarris, 2 Flattop, 3 Hamming, 4 Hanning, 5 Rectangle (fftwindow.h:30-37) */

int b200dsp_spectrum_create(b200dsp_spectrum_t** s, float scalef);            /* scalef: SpectrumVis(Real scalef), 32768 for 16-bit Rx */
int b200dsp_spectrum_destroy(b200dsp_spectrum_t* s);
/* == SpectrumVis::handleConfigure(fftSize, overlapPercent, averageNb, averagingMode, window, linear); restarts the frame
 *    buffer and the averagers like the reference.  Only overlap 0 is supported. */
int b200dsp_spectrum_configure(b200dsp_spectrum_t* s, int fft_size, int overlap_percent, unsigned int average_nb,
                               int averaging_mode, int window, int linear);
/* frames the next feed of n_samples would emit (pure host arithmetic) */
int64_t b200dsp_spectrum_frames_for(b200dsp_spectrum_t* s, int64_t n_samples);
/* == SpectrumVis::feed(begin, end, positiveOnly); out_frames receives *n_frames rows of fft_size floats */
int b200dsp_spectrum_feed(b200dsp_spectrum_t* s, const int16_t* iq, int64_t n_samples, int positive_only,
                          float* out_frames, int64_t cap_frames, int64_t* n_frames);
int b200dsp_spectrum_feed_dev(b200dsp_spectrum_t* s, const void* d_iq, int64_t n_samples, int positive_only,
                              float* d_out_frames, int64_t cap_frames, int64_t* n_frames, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DSP_H */
