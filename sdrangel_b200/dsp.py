"""Host-side mirrors of the reference's decimator classes over the C ABI.

Names and call semantics follow the reference (sdrbase/dsp/decimators.h:277-341, decimatorsfi.h:26-55,
decimatorsff.h, decimatorsif.h:53-83): one object holds six half-band stages of state; `decimateN_{inf,sup,cen}`
take an interleaved I/Q buffer, drop the trailing partial block, carry filter state to the next call and return
the produced samples (the C++ wrappers in include/sdrangel_b200/dsp/ write through `SampleVector::iterator*`
exactly like the reference; in Python the samples are returned as an (n, 2) array).
"""
import ctypes as C
import numpy as np
from . import capi

MODE_INF, MODE_SUP, MODE_CEN = capi.MODE_INF, capi.MODE_SUP, capi.MODE_CEN
_MODES = {"inf": MODE_INF, "sup": MODE_SUP, "cen": MODE_CEN}


class _DecimatorsBase:
    IN_FMT = capi.FMT_I16
    OUT_FMT = capi.FMT_I16

    def __init__(self, input_bits=12, device=None):
        L = capi.lib()
        if device is not None:
            capi.init(device)
        h = C.c_void_p()
        capi.check(L.b200dsp_decim_create(C.byref(h), self.IN_FMT, self.OUT_FMT, input_bits))
        self._h = h
        self.input_bits = input_bits
        self.in_dtype = {capi.FMT_I16: np.int16, capi.FMT_F32: np.float32, capi.FMT_I8: np.int8, capi.FMT_U8: np.uint8}[self.IN_FMT]
        self.out_dtype = np.int16 if self.OUT_FMT == capi.FMT_I16 else np.float32
        self.state_dtype = np.int32 if (self.IN_FMT != capi.FMT_F32 and self.OUT_FMT == capi.FMT_I16) else np.float32

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().b200dsp_decim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- generic entry points -------------------------------------------------------------------------
    def out_count(self, log2, mode, len_scalars):
        return int(capi.lib().b200dsp_decim_out_count(self.IN_FMT, self.OUT_FMT, log2, mode, len_scalars))

    def run(self, log2, mode, buf):
        """Host-buffer call == decimate<2^log2>_<mode>(&it, buf, len): returns the (n_out, 2) samples written."""
        buf = np.ascontiguousarray(buf, dtype=self.in_dtype).reshape(-1)
        n = self.out_count(log2, mode, buf.size)
        if n < 0:
            raise ValueError("bad log2/mode")
        out = np.empty((max(n, 1), 2), dtype=self.out_dtype)
        n_out = C.c_int32(0)
        capi.check(capi.lib().b200dsp_decim_run(self._h, log2, mode, buf.ctypes.data, buf.size, out.ctypes.data, C.byref(n_out)))
        assert n_out.value == n
        return out[:n]

    def run_split(self, log2, mode, buf_i, buf_q):
        """== decimate1 / decimate2_u / decimateN_cen(&it, bufI, bufQ, len): the overloads on separate I and Q arrays."""
        bi = np.ascontiguousarray(buf_i, dtype=self.in_dtype).reshape(-1)
        bq = np.ascontiguousarray(buf_q, dtype=self.in_dtype).reshape(-1)
        assert bi.size == bq.size
        out = np.empty((max(bi.size, 1), 2), dtype=self.out_dtype)
        n_out = C.c_int32(0)
        capi.check(capi.lib().b200dsp_decim_run_split(self._h, log2, mode, bi.ctypes.data, bq.ctypes.data, bi.size, out.ctypes.data, C.byref(n_out)))
        return out[:n_out.value]

    def decimate2_u(self, buf):
        """== Decimators::decimate2_u(&it, buf, len) (decimators.h:374-393)."""
        return self.run(1, capi.MODE_U, buf)

    def run_dev(self, log2, mode, d_in, len_scalars, d_out, stream=None):
        """Device-resident call (metric path): d_in / d_out are device pointers (ints); asynchronous."""
        n_out = C.c_int64(0)
        capi.check(capi.lib().b200dsp_decim_run_dev(self._h, log2, mode, C.c_void_p(d_in), len_scalars, C.c_void_p(d_out),
                                                    C.byref(n_out), C.c_void_p(stream or 0)))
        return n_out.value

    def set_exact_float(self, exact=True):
        capi.check(capi.lib().b200dsp_decim_set_exact_float(self._h, int(bool(exact))))

    def get_state(self):
        st = np.empty((6, 2, 64), dtype=self.state_dtype)
        capi.check(capi.lib().b200dsp_decim_get_state(self._h, st.ctypes.data))
        return st

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=self.state_dtype).reshape(6, 2, 64)
        capi.check(capi.lib().b200dsp_decim_set_state(self._h, st.ctypes.data))

    def reset(self):
        capi.check(capi.lib().b200dsp_decim_reset(self._h))

    def sync(self):
        capi.check(capi.lib().b200dsp_decim_sync(self._h))

    def decimate1(self, buf):
        return self.run(0, MODE_CEN, buf)


def _add_entry_points(cls):
    for log2 in range(1, 7):
        for name, mode in _MODES.items():
            def f(self, buf, _l=log2, _m=mode):
                return self.run(_l, _m, buf)
            f.__name__ = "decimate%d_%s" % (1 << log2, name)
            f.__doc__ = "== %s::decimate%d_%s(it, buf, len)" % (cls.__name__, 1 << log2, name)
            setattr(cls, f.__name__, f)
    return cls


@_add_entry_points
class Decimators(_DecimatorsBase):
    """Decimators<qint32, qint16, 16, input_bits> (sdrbase/dsp/decimators.h:277-341)."""
    IN_FMT, OUT_FMT = capi.FMT_I16, capi.FMT_I16


@_add_entry_points
class Decimators8(_DecimatorsBase):
    """Decimators<qint32, qint8, 16, 8> (plugins/samplesource/hackrfinput/hackrfinputthread.h:57): int8 in, int16 out."""
    IN_FMT, OUT_FMT = capi.FMT_I8, capi.FMT_I16

    def __init__(self, device=None):
        super().__init__(8, device)


@_add_entry_points
class DecimatorsU(_DecimatorsBase):
    """DecimatorsU<qint32, quint8, 16, 8, shift> (sdrbase/dsp/decimatorsu.h:175-215; RTL-SDR: shift 127): uint8 in,
    sample = byte - shift, int16 out."""
    IN_FMT, OUT_FMT = capi.FMT_U8, capi.FMT_I16

    def __init__(self, shift=127, device=None):
        super().__init__(8, device)
        if shift != 127:
            capi.check(capi.lib().b200dsp_decim_set_shift(self._h, int(shift)))
        self.shift = shift


@_add_entry_points
class DecimatorsFI(_DecimatorsBase):
    """DecimatorsFI (sdrbase/dsp/decimatorsfi.h:26-55): float in, int16 out."""
    IN_FMT, OUT_FMT = capi.FMT_F32, capi.FMT_I16


@_add_entry_points
class DecimatorsFF(_DecimatorsBase):
    """DecimatorsFF (sdrbase/dsp/decimatorsff.h): float in, float out."""
    IN_FMT, OUT_FMT = capi.FMT_F32, capi.FMT_F32


@_add_entry_points
class DecimatorsIF(_DecimatorsBase):
    """DecimatorsIF<qint16, input_bits> (sdrbase/dsp/decimatorsif.h:53-83): int16 in, float out."""
    IN_FMT, OUT_FMT = capi.FMT_I16, capi.FMT_F32


class DownChannelizerBank:
    """A bank of reference (DownChannelizer [+ NCO + Interpolator]) channels fed from one baseband stream.

    Mirrors, per channel: DSPConfigureChannelizer -> MsgChannelizerNotification (add_channel), the NFM plugin's
    applyChannelSettings front-end wiring (set_frontend) and DownChannelizer::feed (feed); sdrbase/dsp/downchannelizer.cpp,
    plugins/channelrx/demodnfm/nfmdemod.cpp:150-155,453-476."""

    def __init__(self, input_rate, device=None):
        L = capi.lib()
        if device is not None:
            capi.init(device)
        h = C.c_void_p()
        capi.check(L.b200dsp_bank_create(C.byref(h), int(input_rate)))
        self._h = h
        self.input_rate = int(input_rate)
        self.channels = []
        self._n_channels = 0

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().b200dsp_bank_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_chunk(self, samples):
        capi.check(capi.lib().b200dsp_bank_set_chunk(self._h, int(samples)))

    def add_channel(self, requested_rate, center_offset):
        """Returns (chan_id, out_rate, residual_offset, path) with path a string over 'C','L','U'."""
        cid, rate, ofs = C.c_int32(), C.c_int32(), C.c_int32()
        capi.check(capi.lib().b200dsp_bank_add_channel(self._h, int(requested_rate), int(center_offset), C.byref(cid), C.byref(rate), C.byref(ofs)))
        modes = (C.c_int32 * 32)()
        n = capi.lib().b200dsp_bank_channel_path(self._h, cid.value, modes, 32)
        path = "".join("CLU"[modes[i]] for i in range(n))
        self.channels.append((cid.value, rate.value, ofs.value, path))
        self._n_channels = max(self._n_channels, cid.value + 1)
        return cid.value, rate.value, ofs.value, path

    def add_channel_path(self, path, out_shift):
        """Channel by explicit stages ('C','L','U' string); output = trunc(stage output / 2^out_shift).  Returns chan_id."""
        modes = (C.c_int32 * max(len(path), 1))(*["CLU".index(ch) for ch in path])
        cid = C.c_int32()
        capi.check(capi.lib().b200dsp_bank_add_channel_path(self._h, modes, len(path), int(out_shift), C.byref(cid)))
        self._n_channels = max(self._n_channels, cid.value + 1)
        return cid.value

    def reset(self, stream=None):
        capi.check(capi.lib().b200dsp_bank_reset(self._h, C.c_void_p(stream or 0)))

    def fetch_dev(self, chan_id, stage=capi.STAGE_CHANNELIZER):
        """(device pointer, n_samples) of the channel's output of the last feed (n is -1 for the front-end stage)."""
        ptr, n = C.c_void_p(), C.c_int64(0)
        capi.check(capi.lib().b200dsp_bank_fetch_dev(self._h, chan_id, stage, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def copy_out_dev(self, chan_id, skip, count, d_dst, stream=None):
        capi.check(capi.lib().b200dsp_bank_copy_out_dev(self._h, chan_id, int(skip), int(count), C.c_void_p(d_dst), C.c_void_p(stream or 0)))

    def gather_dev(self, stage, d_out, stride, d_counts, stream=None):
        """Device half of fetch_all: [channel][stride] samples into d_out and int64 counts into d_counts, asynchronously."""
        capi.check(capi.lib().b200dsp_bank_gather_dev(self._h, stage, C.c_void_p(d_out), int(stride), C.c_void_p(d_counts), C.c_void_p(stream or 0)))

    def tree_time(self):
        """(ms, launches) of the tree-level kernels of the last internal pass (waits for it)."""
        ms, k = C.c_float(0), C.c_int32(0)
        capi.check(capi.lib().b200dsp_bank_tree_time(self._h, C.byref(ms), C.byref(k)))
        return ms.value, k.value

    def node_count(self):
        n = capi.lib().b200dsp_bank_node_count(self._h)
        if n < 0:
            capi.check(n)
        return n

    def set_frontend(self, chan_id, nco_freq, cutoff, out_rate, phase_steps=16, taps_per_phase=4.5):
        capi.check(capi.lib().b200dsp_bank_set_frontend(self._h, chan_id, float(nco_freq), phase_steps, float(cutoff), float(taps_per_phase), int(out_rate)))

    def frontend_info(self, chan_id):
        inc, nt = C.c_int32(), C.c_int32()
        capi.check(capi.lib().b200dsp_bank_frontend_info(self._h, chan_id, C.byref(inc), C.byref(nt), None, 0))
        taps = np.empty(nt.value * 256, dtype=np.float32)
        capi.check(capi.lib().b200dsp_bank_frontend_info(self._h, chan_id, C.byref(inc), C.byref(nt), taps.ctypes.data, taps.size))
        return inc.value, nt.value, taps

    def feed(self, iq):
        iq = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1)
        capi.check(capi.lib().b200dsp_bank_feed(self._h, iq.ctypes.data, iq.size // 2))

    def feed_dev(self, d_iq, n_samples, stream=None):
        capi.check(capi.lib().b200dsp_bank_feed_dev(self._h, C.c_void_p(d_iq), int(n_samples), C.c_void_p(stream or 0)))

    def sync(self):
        capi.check(capi.lib().b200dsp_bank_sync(self._h))

    def fetch(self, chan_id, stage=capi.STAGE_CHANNELIZER):
        n = C.c_int64(0)
        capi.check(capi.lib().b200dsp_bank_fetch(self._h, chan_id, stage, None, 1 << 62, C.byref(n)))
        dt = np.int16 if stage == capi.STAGE_CHANNELIZER else np.float32
        out = np.empty((max(n.value, 1), 2), dtype=dt)
        capi.check(capi.lib().b200dsp_bank_fetch(self._h, chan_id, stage, out.ctypes.data, out.shape[0], C.byref(n)))
        return out[:n.value]

    def fetch_schedule(self, chan_id):
        """(input index, phase) of every front-end output of the channel's last internal pass."""
        n = C.c_int64(0)
        capi.check(capi.lib().b200dsp_bank_fetch_schedule(self._h, chan_id, None, None, 1 << 62, C.byref(n)))
        idx, ph = np.empty(n.value, dtype=np.int32), np.empty(n.value, dtype=np.int32)
        if n.value:
            capi.check(capi.lib().b200dsp_bank_fetch_schedule(self._h, chan_id, idx.ctypes.data, ph.ctypes.data, n.value, C.byref(n)))
        return idx, ph

    def process(self, iq_ptr, n_samples, stage, out_ptr, stride):
        """== b200dsp_bank_process: host buffer in (pointer), every channel's outputs to out_ptr [channel][stride]; returns counts."""
        counts = np.zeros(max(self._n_channels, 1), dtype=np.int64)
        capi.check(capi.lib().b200dsp_bank_process(self._h, C.c_void_p(iq_ptr), int(n_samples), stage, C.c_void_p(out_ptr), int(stride), counts.ctypes.data))
        return counts[:self._n_channels]

    def fetch_all(self, stage=capi.STAGE_CHANNELIZER, stride=None, out_ptr=None, stream=None, n_channels=None):
        """Every channel's outputs of the last feed in one transfer.  Returns (array [n_channels, stride, 2], counts);
        with `out_ptr` (e.g. pinned memory of n_channels * stride samples) only the counts array is returned."""
        nc = int(n_channels if n_channels is not None else self._n_channels)
        counts = np.zeros(nc, dtype=np.int64)
        if nc == 0:
            return (np.empty((0, 0, 2), np.int16 if stage == capi.STAGE_CHANNELIZER else np.float32), counts) if out_ptr is None else counts
        if stride is None:
            stride = max(1, max(self.fetch_dev(c, capi.STAGE_CHANNELIZER)[1] for c in range(nc)))
        L = capi.lib()
        if out_ptr is not None:
            capi.check(L.b200dsp_bank_fetch_all(self._h, stage, C.c_void_p(out_ptr), int(stride), counts.ctypes.data, C.c_void_p(stream or 0)))
            return counts
        out = np.empty((nc, int(stride), 2), dtype=np.int16 if stage == capi.STAGE_CHANNELIZER else np.float32)
        capi.check(L.b200dsp_bank_fetch_all(self._h, stage, out.ctypes.data, int(stride), counts.ctypes.data, C.c_void_p(stream or 0)))
        return out, counts


class ShardedBank:
    """One rank's share of a DownChannelizer bank sharded by channel over the GPUs of one box (include/b200dsp.h, K6):
    an ordinary bank for channels [lo, hi) of the plan plus the b200dsp_dist object that brings the whole baseband to this
    GPU (NCCL broadcast from a device buffer, or sliced host-to-device copies + in-place all-gather).  The 128-byte NCCL id
    made by one rank has to reach the others by the caller's own means (torch.distributed in bench.py)."""

    def __init__(self, input_rate, offsets, requested_rate, rank, world, nccl_id, frontend=None, chunk=None):
        L = capi.lib()
        lo, hi = C.c_int32(), C.c_int32()
        capi.check(L.b200dsp_dist_shard(len(offsets), world, rank, C.byref(lo), C.byref(hi)))
        self.lo, self.hi, self.rank, self.world = lo.value, hi.value, rank, world
        self.bank = DownChannelizerBank(input_rate)
        if chunk:
            self.bank.set_chunk(chunk)
        self.info = []
        for fc in offsets[self.lo:self.hi]:
            cid, rate, ofs, path = self.bank.add_channel(requested_rate, fc)
            if frontend is not None:
                cutoff, out_rate = frontend
                self.bank.set_frontend(cid, -ofs, cutoff, out_rate)
            self.info.append((cid, rate, ofs, path))
        h = C.c_void_p()
        idb = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
        capi.check(L.b200dsp_dist_create(C.byref(h), idb, rank, world))
        self._h = h

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        capi.check(capi.lib().b200dsp_dist_unique_id(buf))
        return bytes(buf)

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().b200dsp_dist_destroy(self._h)
            self._h = None
        if getattr(self, "bank", None):
            self.bank.close()
            self.bank = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reserve(self, n_samples):
        capi.check(capi.lib().b200dsp_dist_reserve(self._h, int(n_samples)))

    def bcast_begin(self, slot, d_iq, n_samples, root=0, after_stream=None):
        capi.check(capi.lib().b200dsp_dist_bcast_begin(self._h, slot, C.c_void_p(d_iq or 0), int(n_samples), root, C.c_void_p(after_stream or 0)))

    def ingest_begin(self, slot, host_slice_ptr, n_samples_total):
        capi.check(capi.lib().b200dsp_dist_ingest_begin(self._h, slot, C.c_void_p(host_slice_ptr), int(n_samples_total)))

    def feed(self, slot, stream=None):
        capi.check(capi.lib().b200dsp_dist_feed(self._h, slot, self.bank._h, C.c_void_p(stream or 0)))

    def slot(self, slot):
        ptr, n = C.c_void_p(), C.c_int64(0)
        capi.check(capi.lib().b200dsp_dist_slot(self._h, slot, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def sync(self):
        capi.check(capi.lib().b200dsp_dist_sync(self._h))

    # copy-engine chain (one process per GPU): export -> gather the blobs of all ranks in rank order -> import
    P2P_SLOTS = 3

    def p2p_export(self, n_samples):
        blob = (C.c_char * 512)()
        capi.check(capi.lib().b200dsp_dist_p2p_export(self._h, int(n_samples), blob))
        return bytes(blob)

    def p2p_import(self, blobs):
        data = b"".join(bytes(b) for b in blobs)
        buf = (C.c_char * len(data)).from_buffer_copy(data)
        capi.check(capi.lib().b200dsp_dist_p2p_import(self._h, buf))

    def p2p_begin(self, slot, d_iq, n_samples, after_stream=None):
        capi.check(capi.lib().b200dsp_dist_p2p_begin(self._h, slot, C.c_void_p(d_iq or 0), int(n_samples), C.c_void_p(after_stream or 0)))

    def p2p_feed(self, slot, stream=None, consume_only=False):
        capi.check(capi.lib().b200dsp_dist_p2p_feed(self._h, slot, None if consume_only else self.bank._h, C.c_void_p(stream or 0)))

    def p2p_slot(self, slot):
        ptr, n = C.c_void_p(), C.c_int64(0)
        capi.check(capi.lib().b200dsp_dist_p2p_slot(self._h, slot, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value


class SpectrumVis:
    """SpectrumVis (sdrgui/dsp/spectrumvis.cpp): feed() returns the frames the reference would pass to
    GLSpectrum::newSpectrum, as an (n_frames, fft_size) float32 array."""
    AvgModeNone, AvgModeMoving, AvgModeFixed = 0, 1, 2

    def __init__(self, scalef=32768.0, device=None):
        L = capi.lib()
        if device is not None:
            capi.init(device)
        h = C.c_void_p()
        capi.check(L.b200dsp_spectrum_create(C.byref(h), float(scalef)))
        self._h = h
        self.fft_size = 1024

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().b200dsp_spectrum_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def configure(self, fft_size, overlap_pct=0, avg_nb=0, avg_mode=0, window=1, linear=False):
        capi.check(capi.lib().b200dsp_spectrum_configure(self._h, fft_size, overlap_pct, avg_nb, avg_mode, window, int(linear)))
        self.fft_size = min(max(fft_size, 64), 4096)

    def frames_for(self, n_samples):
        return int(capi.lib().b200dsp_spectrum_frames_for(self._h, int(n_samples)))

    def feed(self, iq, positive_only=False):
        iq = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1)
        n = iq.size // 2
        nf = self.frames_for(n)
        out = np.empty((max(nf, 1), self.fft_size), dtype=np.float32)
        got = C.c_int64(0)
        capi.check(capi.lib().b200dsp_spectrum_feed(self._h, iq.ctypes.data, n, int(positive_only), out.ctypes.data, out.shape[0], C.byref(got)))
        return out[:got.value]

    def feed_dev(self, d_iq, n_samples, d_out, cap_frames, positive_only=False, stream=None):
        got = C.c_int64(0)
        capi.check(capi.lib().b200dsp_spectrum_feed_dev(self._h, C.c_void_p(d_iq), int(n_samples), int(positive_only), C.c_void_p(d_out),
                                                        int(cap_frames), C.byref(got), C.c_void_p(stream or 0)))
        return got.value


class Interpolator:
    """Interpolator (sdrbase/dsp/interpolator.h:19-36): create() + the block form of the Rx plugins' decimate loop
    (plugins/channelrx/demodnfm/nfmdemod.cpp:150-155,315).  `distance_remain` is caller-owned, like the Real* in the reference."""

    def __init__(self, phase_steps, sample_rate, cutoff, taps_per_phase=4.5, device=None):
        L = capi.lib()
        if device is not None:
            capi.init(device)
        h = C.c_void_p()
        capi.check(L.b200dsp_interp_create(C.byref(h), int(phase_steps), float(sample_rate), float(cutoff), float(taps_per_phase)))
        self._h = h
        self.phase_steps = int(phase_steps)

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().b200dsp_interp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def taps(self):
        nt = C.c_int32()
        capi.check(capi.lib().b200dsp_interp_info(self._h, C.byref(nt), None, 0))
        t = np.empty(nt.value * self.phase_steps, dtype=np.float32)
        capi.check(capi.lib().b200dsp_interp_info(self._h, C.byref(nt), t.ctypes.data, t.size))
        return t.reshape(self.phase_steps, nt.value)

    def decimate(self, distance_remain, distance, samples):
        """Returns (outputs complex64, new distance_remain)."""
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        out = np.empty(x.size + 2, dtype=np.complex64)
        rem = C.c_float(float(distance_remain))
        n = C.c_int64(0)
        capi.check(capi.lib().b200dsp_interp_decimate(self._h, C.byref(rem), float(distance), x.ctypes.data, x.size, out.ctypes.data, out.size, C.byref(n)))
        return out[:n.value], rem.value

    def _up(self, fn, distance_remain, distance, samples):
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        out = np.empty(int((x.size + 2) / float(distance)) + x.size + 8, dtype=np.complex64)
        rem = C.c_float(float(distance_remain))
        n = C.c_int64(0)
        capi.check(fn(self._h, C.byref(rem), float(distance), x.ctypes.data if x.size else None, x.size, out.ctypes.data, out.size, C.byref(n)))
        return out[:n.value], rem.value

    def interpolate(self, distance_remain, distance, samples):
        """== the Tx plugins' loop around Interpolator::interpolate (nfmmod.cpp:126-133): (outputs, new distance_remain)."""
        return self._up(capi.lib().b200dsp_interp_interpolate, distance_remain, distance, samples)

    def resample(self, distance_remain, distance, samples):
        """== the canonical loop around Interpolator::resample (interpolator.h:55-76): (outputs, new distance_remain)."""
        return self._up(capi.lib().b200dsp_interp_resample, distance_remain, distance, samples)


class NCO:
    """NCO (sdrbase/dsp/nco.h:40-53, nco.cpp:30-64): setFreq, setPhase, block form of nextIQ()."""

    def __init__(self, device=None):
        if device is not None:
            capi.init(device)
        h = C.c_void_p()
        capi.check(capi.lib().b200dsp_nco_create(C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().b200dsp_nco_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setFreq(self, freq, sample_rate):
        capi.check(capi.lib().b200dsp_nco_set_freq(self._h, float(freq), float(sample_rate)))

    def setPhase(self, phase):
        capi.check(capi.lib().b200dsp_nco_set_phase(self._h, int(phase)))

    def state(self):
        p, i = C.c_int32(), C.c_int32()
        capi.check(capi.lib().b200dsp_nco_get(self._h, C.byref(p), C.byref(i)))
        return p.value, i.value

    def nextIQ(self, n):
        out = np.empty(int(n), dtype=np.complex64)
        capi.check(capi.lib().b200dsp_nco_next_iq(self._h, int(n), out.ctypes.data))
        return out


class IQCorrections:
    """DSPDeviceSourceEngine::iqCorrections (sdrbase/dsp/dspdevicesourceengine.cpp:175-262): the engine's DC (and I/Q
    imbalance) correction of the decimated baseband before it reaches the sinks.  One object == one engine's state."""

    def __init__(self, device=None):
        if device is not None:
            capi.init(device)
        h = C.c_void_p()
        capi.check(capi.lib().b200dsp_iqcorr_create(C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().b200dsp_iqcorr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        capi.check(capi.lib().b200dsp_iqcorr_reset(self._h))

    def iqCorrections(self, iq, imbalanceCorrection=False):
        """Corrects the (n, 2) int16 samples in place, like the reference does on the FIFO's vector; returns the array."""
        a = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1, 2)
        capi.check(capi.lib().b200dsp_iqcorr_run(self._h, a.ctypes.data, a.shape[0], int(bool(imbalanceCorrection))))
        if a is not iq and isinstance(iq, np.ndarray) and iq.dtype == np.int16 and iq.size == a.size:
            iq.reshape(-1, 2)[...] = a
        return a

    def run_dev(self, d_in, d_out, n_samples, stream=None):
        capi.check(capi.lib().b200dsp_iqcorr_run_dev(self._h, C.c_void_p(d_in), C.c_void_p(d_out), int(n_samples), 0, C.c_void_p(stream or 0)))


class Interpolators:
    """Interpolators<T, 16, OutputBits> (sdrbase/dsp/interpolators.h:104-617): the device-side Tx interpolators.
    output_bits 16 / 12 -> int16 device buffers, 8 -> int8 (HackRF).  interpolateN_cen(samples, len) returns the device buffer
    (len output scalars, the trailing partial block untouched = the fill value) and the number of Samples consumed."""

    def __init__(self, output_bits=16, device=None):
        if device is not None:
            capi.init(device)
        self.output_bits = int(output_bits)
        self.fmt = capi.FMT_I8 if self.output_bits == 8 else capi.FMT_I16
        self.dtype = np.int8 if self.output_bits == 8 else np.int16
        h = C.c_void_p()
        capi.check(capi.lib().b200dsp_interps_create(C.byref(h), self.fmt, self.output_bits))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().b200dsp_interps_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        capi.check(capi.lib().b200dsp_interps_reset(self._h))

    def run(self, log2, samples, length=None, fill=0):
        """== interpolate{2^log2}_cen(&it, buf, len): samples int16 [n, 2]; len defaults to everything the samples give."""
        x = np.ascontiguousarray(samples, dtype=np.int16).reshape(-1, 2)
        if length is None:
            length = x.shape[0] * (2 << log2)
        need = capi.lib().b200dsp_interps_in_count(log2, length)
        if need < 0 or need > x.shape[0]:
            raise ValueError("interpolate: %d output scalars need %d samples, %d given" % (length, need, x.shape[0]))
        buf = np.full(int(length), fill, dtype=self.dtype)
        n = C.c_int32(0)
        capi.check(capi.lib().b200dsp_interps_run(self._h, log2, x.ctypes.data if x.size else None, buf.ctypes.data if buf.size else None, int(length), C.byref(n)))
        return buf, n.value

    def run_dev(self, log2, d_samples, d_buf, len_scalars, stream=None):
        n = C.c_int64(0)
        capi.check(capi.lib().b200dsp_interps_run_dev(self._h, log2, d_samples, d_buf, len_scalars, C.byref(n), stream))
        return n.value

    def interpolate1(self, samples, length=None):
        return self.run(0, samples, length)


for _l in range(1, 7):
    setattr(Interpolators, "interpolate%d_cen" % (1 << _l), (lambda l: lambda self, samples, length=None, fill=0: self.run(l, samples, length, fill))(_l))


class UpChannelizer:
    """UpChannelizer (sdrbase/dsp/upchannelizer.cpp:51-104,175-209,252-327): block form of pull()."""

    def __init__(self, device=None):
        if device is not None:
            capi.init(device)
        h = C.c_void_p()
        capi.check(capi.lib().b200dsp_upchan_create(C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().b200dsp_upchan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def configure(self, output_rate, requested_rate, center_offset):
        """Returns (modulator rate, residual offset, path) -- what MsgChannelizerNotification reports + the stage modes."""
        r, o = C.c_int32(), C.c_int32()
        capi.check(capi.lib().b200dsp_upchan_configure(self._h, output_rate, requested_rate, center_offset, C.byref(r), C.byref(o)))
        return r.value, o.value, self.path()

    def set_path(self, modes):
        m = (C.c_int32 * max(1, len(modes)))(*modes)
        capi.check(capi.lib().b200dsp_upchan_set_path(self._h, m, len(modes)))

    def path(self):
        m = (C.c_int32 * 32)()
        n = capi.lib().b200dsp_upchan_path(self._h, m, 32)
        return [int(m[i]) for i in range(n)]

    def source_count(self, n_out):
        return int(capi.lib().b200dsp_upchan_source_count(self._h, n_out))

    def pull(self, source, n_out):
        """n_out calls of pull(); source int16 [>= source_count(n_out), 2].  Returns (out int16 [n_out, 2], samples consumed)."""
        need = self.source_count(n_out)
        x = np.ascontiguousarray(source, dtype=np.int16).reshape(-1, 2)
        out = np.empty((n_out, 2), dtype=np.int16)
        capi.check(capi.lib().b200dsp_upchan_pull(self._h, x.ctypes.data if x.size else None, x.shape[0], out.ctypes.data if n_out else None, n_out))
        return out, need

    def pull_dev(self, d_source, n_source, d_out, n_out, stream=None):
        capi.check(capi.lib().b200dsp_upchan_pull_dev(self._h, d_source, n_source, d_out, n_out, stream))


class Demod:
    """Demodulator back-ends after Interpolator::decimate: PhaseDiscriminators (phasediscri.h:26-198; kinds
    0 atan2, 1 delta, 2 discri2) and the AM magnitude of AMDemod::processOneSample (amdemod.cpp:154-156,241; kind 3)."""
    FM_ATAN2, FM_DELTA, FM_DISCRI2, AM_MAG = 0, 1, 2, 3

    def __init__(self, kind, fm_scaling=1.0, n_channels=1, device=None):
        if device is not None:
            capi.init(device)
        h = C.c_void_p()
        capi.check(capi.lib().b200dsp_demod_create(C.byref(h), kind, float(fm_scaling), n_channels))
        self._h, self.kind, self.n_channels = h, kind, n_channels

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().b200dsp_demod_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        capi.check(capi.lib().b200dsp_demod_reset(self._h))

    def setFMScaling(self, fm_scaling):
        capi.check(capi.lib().b200dsp_demod_set_fm_scaling(self._h, float(fm_scaling)))

    def run(self, x):
        """One stream: (out, aux0, aux1) float32 arrays (aux0 = magsq for kinds 1 and 3, aux1 = fmDev for kind 1)."""
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out, a0, a1 = (np.zeros(x.size, dtype=np.float32) for _ in range(3))
        capi.check(capi.lib().b200dsp_demod_run(self._h, x.ctypes.data if x.size else None, x.size, out.ctypes.data, a0.ctypes.data, a1.ctypes.data))
        return out, a0, a1

    def run_pool_dev(self, d_pool, stride, d_counts, d_out, out_stride, d_aux0=None, d_aux1=None, stream=None):
        capi.check(capi.lib().b200dsp_demod_run_pool_dev(self._h, d_pool, stride, d_counts, self.n_channels, d_out, out_stride, d_aux0, d_aux1, stream))


class SdriqFile:
    """.sdriq record files (FileRecord, sdrbase/dsp/filerecord.cpp:72-148): host-side reader / writer."""

    def __init__(self, path, mode="r", sample_rate=0, center_frequency=0, timestamp=0):
        h = C.c_void_p()
        self.mode = mode
        if mode == "r":
            r, c, t, s, n = C.c_int32(), C.c_uint64(), C.c_int64(), C.c_uint32(), C.c_int64()
            capi.check(capi.lib().b200dsp_sdriq_open(C.byref(h), path.encode(), C.byref(r), C.byref(c), C.byref(t), C.byref(s), C.byref(n)))
            self.sample_rate, self.center_frequency, self.timestamp, self.sample_size, self.n_samples = r.value, c.value, t.value, s.value, n.value
        else:
            capi.check(capi.lib().b200dsp_sdriq_create(C.byref(h), path.encode(), sample_rate, center_frequency, timestamp))
        self._h = h

    def read(self, n):
        out = np.empty((n, 2), dtype=np.int16)
        got = C.c_int64(0)
        capi.check(capi.lib().b200dsp_sdriq_read(self._h, out.ctypes.data, n, C.byref(got)))
        return out[:got.value]

    def write(self, iq):
        iq = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1, 2)
        capi.check(capi.lib().b200dsp_sdriq_write(self._h, iq.ctypes.data if iq.size else None, iq.shape[0]))

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().b200dsp_sdriq_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FftFilt:
    """fftfilt (sdrbase/dsp/fftfilt.cpp:49-360): kind 0 == fftfilt(f1, f2, len), 1 == fftfilt(f2, len); block forms of runFilt /
    runSSB / runDSB (ops 0 / 1 / 2)."""

    def __init__(self, kind, f1, f2, length, device=None):
        if device is not None:
            capi.init(device)
        h = C.c_void_p()
        capi.check(capi.lib().b200dsp_fftfilt_create(C.byref(h), kind, float(f1), float(f2), length))
        self._h, self.flen = h, length

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().b200dsp_fftfilt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_filter(self, kind, f1, f2):
        capi.check(capi.lib().b200dsp_fftfilt_set_filter(self._h, kind, float(f1), float(f2)))

    def filter(self):
        out = np.zeros(self.flen, dtype=np.complex64)
        capi.lib().b200dsp_fftfilt_filter(self._h, out.ctypes.data, self.flen)
        return out

    def run(self, op, x, usb=True, get_dc=True):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.zeros(x.size + self.flen, dtype=np.complex64)
        n = C.c_int64(0)
        capi.check(capi.lib().b200dsp_fftfilt_run(self._h, op, int(bool(usb)), int(bool(get_dc)), x.ctypes.data if x.size else None, x.size, out.ctypes.data, out.size, C.byref(n)))
        return out[:n.value].copy()

    def run_dev(self, op, d_in, n, d_out, cap, usb=True, get_dc=True, stream=None):
        m = C.c_int64(0)
        capi.check(capi.lib().b200dsp_fftfilt_run_dev(self._h, op, int(bool(usb)), int(bool(get_dc)), d_in, n, d_out, cap, C.byref(m), stream))
        return m.value
