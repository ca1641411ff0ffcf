"""ORACLE (test infrastructure): ctypes binding of oracle/_ref/libsdrref.so — the UNMODIFIED reference
DSP classes compiled from /root/reference by oracle/Makefile (driver: oracle/ref_capi.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Never imported by the product package sdrangel_b200.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MODE_INF, MODE_SUP, MODE_CEN = 0, 1, 2


def lib_path(strict=False):
    return os.path.join(_HERE, "_ref", "libsdrref_strict.so" if strict else "libsdrref.so")


def available(strict=False):
    return os.path.exists(lib_path(strict))


_libs = {}


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def load(strict=False):
    if strict in _libs:
        return _libs[strict]
    L = C.CDLL(lib_path(strict))
    vp, i32, f32, f64 = C.c_void_p, C.c_int, C.c_float, C.c_double
    pi16, pf32, pi32 = C.POINTER(C.c_int16), C.POINTER(C.c_float), C.POINTER(C.c_int32)
    sig = {
        "ref_decim_ii_create": (vp, [i32]), "ref_decim_ii_destroy": (None, [vp]),
        "ref_decim_ii_run": (i32, [vp, i32, i32, pi16, i32, pi16]),
        "ref_decim_i8_create": (vp, []), "ref_decim_i8_destroy": (None, [vp]),
        "ref_decim_i8_run": (i32, [vp, i32, i32, C.POINTER(C.c_int8), i32, pi16]),
        "ref_decim_u8_create": (vp, []), "ref_decim_u8_destroy": (None, [vp]),
        "ref_decim_u8_run": (i32, [vp, i32, i32, C.POINTER(C.c_uint8), i32, pi16]),
        "ref_iqcorr_create": (vp, []), "ref_iqcorr_destroy": (None, [vp]), "ref_iqcorr_dc": (None, [vp, pi16, i32]), "ref_iqcorr_imbalance": (None, [vp, pi16, i32]),
        "ref_decim_fi_create": (vp, []), "ref_decim_fi_destroy": (None, [vp]),
        "ref_decim_fi_run": (i32, [vp, i32, i32, pf32, i32, pi16]),
        "ref_decim_ff_create": (vp, []), "ref_decim_ff_destroy": (None, [vp]),
        "ref_decim_ff_run": (i32, [vp, i32, i32, pf32, i32, pf32]),
        "ref_decim_if_create": (vp, [i32]), "ref_decim_if_destroy": (None, [vp]),
        "ref_decim_if_run": (i32, [vp, i32, i32, pi16, i32, pf32]),
        "ref_chan_create": (vp, []), "ref_chan_destroy": (None, [vp]),
        "ref_chan_configure": (i32, [vp, i32, i32, i32, pi32, pi32, pi32, i32]),
        "ref_chan_feed": (i32, [vp, pi16, i32, pi16, i32]),
        "ref_frontend_create": (vp, [f32, f32, i32, f64, f64, f64, f32]), "ref_frontend_destroy": (None, [vp]),
        "ref_frontend_nco_increment": (i32, [vp]), "ref_frontend_ntaps": (i32, [vp]),
        "ref_frontend_taps": (None, [vp, pf32]), "ref_nco_table": (None, [pf32]),
        "ref_frontend_feed": (i32, [vp, pi16, i32, pf32, i32, pi32, pi32]),
        "ref_decim_ii_run_split": (i32, [vp, i32, i32, pi16, pi16, i32, pi16]), "ref_decim_ii_run_2u": (i32, [vp, pi16, i32, pi16]),
        "ref_interp_decimate": (i32, [vp, pf32, i32, pf32, i32]),
        "ref_interp_interpolate": (i32, [vp, pf32, i32, pf32, i32]), "ref_interp_resample": (i32, [vp, pf32, i32, pf32, i32]),
        "ref_frontend_remain": (f32, [vp]), "ref_nco_block": (None, [f32, f32, i32, pf32]),
        "ref_spectrum_create": (vp, [f32]), "ref_spectrum_destroy": (None, [vp]),
        "ref_spectrum_configure": (None, [vp, i32, i32, C.c_uint, i32, i32, i32]),
        "ref_spectrum_window": (None, [vp, pf32]), "ref_spectrum_fft": (None, [vp, pf32, pf32]),
        "ref_spectrum_feed": (i32, [vp, pi16, i32, i32, pf32, i32]),
        "ref_sdrbench_gen_s16": (None, [pi16, i32]), "ref_sdrbench_gen_f32": (None, [pf32, i32]),
        "ref_build_info": (C.c_char_p, []),
        "ref_interps_create": (vp, [i32]), "ref_interps_destroy": (None, [vp]), "ref_interps_run": (i32, [vp, i32, pi16, i32, vp, i32]),
        "ref_hb_coeffs": (i32, [i32, pi32, pi32]),
        "ref_upchan_create": (vp, []), "ref_upchan_destroy": (None, [vp]),
        "ref_upchan_configure": (i32, [vp, i32, i32, i32, pi32, pi32, pi32, i32]),
        "ref_upchan_pull": (i32, [vp, pi16, i32, pi16, i32]),
        "ref_discri_create": (vp, [f32]), "ref_discri_destroy": (None, [vp]), "ref_discri_run": (None, [vp, i32, pf32, i32, pf32, pf32, pf32]),
        "ref_fftfilt_create": (vp, [i32, f32, f32, i32]), "ref_fftfilt_destroy": (None, [vp]), "ref_fftfilt_set": (None, [vp, i32, f32, f32]),
        "ref_fftfilt_filter": (None, [vp, pf32]), "ref_fftfilt_run": (i32, [vp, i32, i32, pf32, i32, pf32, i32]),
        "ref_filerecord_write": (i32, [C.c_char_p, i32, C.c_longlong, pi16, i32, i32]),
        "ref_filerecord_read_header": (i32, [C.c_char_p, pi32, C.POINTER(C.c_ulonglong), C.POINTER(C.c_longlong), C.POINTER(C.c_uint)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _libs[strict] = L
    return L


class _Handle:
    def __init__(self, lib, h, destroy):
        self.lib, self.h, self._destroy = lib, h, destroy

    def __del__(self):
        try:
            if self.h:
                self._destroy(self.h)
                self.h = None
        except Exception:
            pass


class RefDecimators(_Handle):
    """kind: 'ii' (int16->int16), 'fi' (float->int16), 'ff' (float->float), 'if' (int16->float),
    'i8' (int8->int16, Decimators<qint32,qint8,16,8>), 'u8' (uint8->int16, DecimatorsU<qint32,quint8,16,8,127>)."""

    def __init__(self, kind="ii", input_bits=12, strict=False):
        L = load(strict)
        self.kind = kind
        if kind in ("ii", "if"):
            h = getattr(L, f"ref_decim_{kind}_create")(input_bits)
        else:
            h = getattr(L, f"ref_decim_{kind}_create")()
        super().__init__(L, h, getattr(L, f"ref_decim_{kind}_destroy"))
        self._run = getattr(L, f"ref_decim_{kind}_run")
        self.in_dt = {"i8": np.int8, "u8": np.uint8}.get(kind, np.int16 if kind[0] == "i" else np.float32)
        self.out_dt = np.int16 if kind[1] in "i8" else np.float32

    def run(self, log2, mode, buf):
        buf = np.ascontiguousarray(buf, dtype=self.in_dt)
        out = np.empty((buf.size // 2 + 8, 2), dtype=self.out_dt)
        ct_in = {np.int16: C.c_int16, np.float32: C.c_float, np.int8: C.c_int8, np.uint8: C.c_uint8}[self.in_dt]
        ct_out = C.c_int16 if self.out_dt == np.int16 else C.c_float
        n = self._run(self.h, log2, mode, _p(buf, ct_in), buf.size, _p(out, ct_out))
        if n < 0:
            raise ValueError("bad log2/mode")
        return out[:n].copy()

    def run_split(self, log2, buf_i, buf_q, u=False):
        """The split-I/Q overloads: decimate1 / decimateN_cen(it, bufI, bufQ, len), or decimate2_u with u=True ('ii', 12 bits)."""
        bi, bq = np.ascontiguousarray(buf_i, dtype=np.int16), np.ascontiguousarray(buf_q, dtype=np.int16)
        out = np.empty((bi.size + 8, 2), dtype=np.int16)
        n = self.lib.ref_decim_ii_run_split(self.h, log2, int(u), _p(bi, C.c_int16), _p(bq, C.c_int16), bi.size, _p(out, C.c_int16))
        if n < 0:
            raise ValueError("bad log2")
        return out[:n].copy()

    def run_2u(self, buf):
        buf = np.ascontiguousarray(buf, dtype=np.int16)
        out = np.empty((buf.size // 2 + 8, 2), dtype=np.int16)
        n = self.lib.ref_decim_ii_run_2u(self.h, _p(buf, C.c_int16), buf.size, _p(out, C.c_int16))
        return out[:n].copy()


class RefIQCorrections(_Handle):
    """DSPDeviceSourceEngine::iqCorrections(begin, end, false): DC correction in place."""

    def __init__(self, strict=False):
        L = load(strict)
        super().__init__(L, L.ref_iqcorr_create(), L.ref_iqcorr_destroy)

    def run(self, iq, imbalance=False):
        a = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1, 2).copy()
        (self.lib.ref_iqcorr_imbalance if imbalance else self.lib.ref_iqcorr_dc)(self.h, _p(a, C.c_int16), a.shape[0])
        return a


class RefDownChannelizer(_Handle):
    def __init__(self, strict=False):
        L = load(strict)
        super().__init__(L, L.ref_chan_create(), L.ref_chan_destroy)

    def configure(self, input_rate, requested_rate, center_offset):
        rate, ofs = C.c_int32(0), C.c_int32(0)
        modes = np.zeros(32, dtype=np.int32)
        n = self.lib.ref_chan_configure(self.h, input_rate, requested_rate, center_offset,
                                        C.byref(rate), C.byref(ofs), _p(modes, C.c_int32), 32)
        return rate.value, ofs.value, [int(m) for m in modes[:n]]

    def feed(self, iq):
        iq = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1, 2)
        out = np.empty((iq.shape[0] + 8, 2), dtype=np.int16)
        n = self.lib.ref_chan_feed(self.h, _p(iq, C.c_int16), iq.shape[0], _p(out, C.c_int16), out.shape[0])
        assert n >= 0
        return out[:n].copy()


class RefFrontEnd(_Handle):
    """NCO + Interpolator::decimate as a channel plugin wires them (nfmdemod.cpp:152-155,315,462-470)."""

    def __init__(self, nco_freq, rate, out_rate, cutoff, phase_steps=16, taps_per_phase=4.5, strict=False):
        L = load(strict)
        distance = np.float32(np.float32(rate) / np.float32(out_rate))
        h = L.ref_frontend_create(float(nco_freq), float(rate), phase_steps, float(rate), float(cutoff),
                                  float(taps_per_phase), float(distance))
        super().__init__(L, h, L.ref_frontend_destroy)
        self.phase_steps = phase_steps

    def nco_increment(self):
        return self.lib.ref_frontend_nco_increment(self.h)

    def taps(self):
        nt = self.lib.ref_frontend_ntaps(self.h)
        t = np.empty(nt * self.phase_steps, dtype=np.float32)
        self.lib.ref_frontend_taps(self.h, _p(t, C.c_float))
        return t.reshape(self.phase_steps, nt)

    def feed(self, iq, want_schedule=False):
        iq = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1, 2)
        n = iq.shape[0]
        out = np.empty((n + 8, 2), dtype=np.float32)
        idx = np.empty(n + 8, dtype=np.int32)
        ph = np.empty(n + 8, dtype=np.int32)
        m = self.lib.ref_frontend_feed(self.h, _p(iq, C.c_int16), n, _p(out, C.c_float), n + 8,
                                       _p(idx, C.c_int32), _p(ph, C.c_int32))
        assert m >= 0
        if want_schedule:
            return out[:m].copy(), idx[:m].copy(), ph[:m].copy()
        return out[:m].copy()

    def run_c64(self, mode, x, cap=None):
        """Interpolator::decimate (mode 0) / interpolate (1) / resample (2) on complex64 input in the callers' loops."""
        x = np.ascontiguousarray(x, dtype=np.complex64)
        n = x.shape[0]
        cap = int(cap or (n * 64 + 64))
        out = np.empty((cap, 2), dtype=np.float32)
        fn = (self.lib.ref_interp_decimate, self.lib.ref_interp_interpolate, self.lib.ref_interp_resample)[mode]
        m = fn(self.h, _p(x.view(np.float32), C.c_float), n, _p(out, C.c_float), cap)
        assert m >= 0
        return out[:m].copy().view(np.complex64).reshape(-1)

    def remain(self):
        return float(self.lib.ref_frontend_remain(self.h))


def nco_block(freq, rate, n, strict=False):
    out = np.empty((n, 2), dtype=np.float32)
    load(strict).ref_nco_block(float(freq), float(rate), int(n), _p(out, C.c_float))
    return out.view(np.complex64).reshape(-1)


def nco_table(strict=False):
    t = np.empty(4096, dtype=np.float32)
    load(strict).ref_nco_table(_p(t, C.c_float))
    return t


class RefSpectrumVis(_Handle):
    AVG_NONE, AVG_MOVING, AVG_FIXED = 0, 1, 2
    WINDOWS = {"bartlett": 0, "blackmanharris": 1, "flattop": 2, "hamming": 3, "hanning": 4, "rectangle": 5}

    def __init__(self, scalef=32768.0, strict=False):
        L = load(strict)
        super().__init__(L, L.ref_spectrum_create(scalef), L.ref_spectrum_destroy)
        self.fft_size = 1024

    def configure(self, fft_size, overlap_pct=0, avg_nb=0, avg_mode=0, window=1, linear=False):
        self.lib.ref_spectrum_configure(self.h, fft_size, overlap_pct, avg_nb, avg_mode, window, int(linear))
        self.fft_size = min(max(fft_size, 64), 4096)

    def window(self):
        w = np.empty(self.fft_size, dtype=np.float32)
        self.lib.ref_spectrum_window(self.h, _p(w, C.c_float))
        return w

    def fft(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        assert x.size == self.fft_size
        out = np.empty(self.fft_size, dtype=np.complex64)
        self.lib.ref_spectrum_fft(self.h, x.view(np.float32).ctypes.data_as(C.POINTER(C.c_float)),
                                  out.view(np.float32).ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def feed(self, iq, positive_only=False):
        iq = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1, 2)
        cap = iq.shape[0] // self.fft_size + 2
        frames = np.empty((cap, self.fft_size), dtype=np.float32)
        n = self.lib.ref_spectrum_feed(self.h, _p(iq, C.c_int16), iq.shape[0], int(positive_only),
                                       _p(frames, C.c_float), cap)
        assert n >= 0
        return frames[:n].copy()


def sdrbench_s16(n_samples, strict=False):
    buf = np.empty(2 * n_samples, dtype=np.int16)
    load(strict).ref_sdrbench_gen_s16(_p(buf, C.c_int16), buf.size)
    return buf


def sdrbench_f32(n_samples, strict=False):
    buf = np.empty(2 * n_samples, dtype=np.float32)
    load(strict).ref_sdrbench_gen_f32(_p(buf, C.c_float), buf.size)
    return buf


def fnv1a64_u16(a):
    """FNV-1a-64 fed one uint16 at a time (SURVEY.md Appendix D convention)."""
    a = np.ascontiguousarray(a).view(np.uint16).ravel()
    h = 1469598103934665603
    prime = 1099511628211
    mask = (1 << 64) - 1
    for v in a.tolist():
        h = ((h ^ v) * prime) & mask
    return "%016x" % h


class RefInterpolators(_Handle):
    """The reference's Interpolators<qint16,16,16> / <qint16,16,12> / <qint8,16,8> (oracle/ref_capi_tx.cpp)."""

    def __init__(self, output_bits=16, strict=False):
        L = load(strict)
        super().__init__(L, L.ref_interps_create(output_bits), L.ref_interps_destroy)
        self.dtype = np.int8 if output_bits == 8 else np.int16

    def run(self, log2, samples, length=None, fill=0):
        x = np.ascontiguousarray(samples, dtype=np.int16).reshape(-1, 2)
        if length is None:
            length = x.shape[0] * (2 << log2)
        assert length // (2 << log2) <= x.shape[0]
        buf = np.full(int(length), fill, dtype=self.dtype)
        n = self.lib.ref_interps_run(self.h, log2, _p(x, C.c_int16), x.shape[0], buf.ctypes.data, int(length))
        return buf, n


class RefUpChannelizer(_Handle):
    def __init__(self, strict=False):
        L = load(strict)
        super().__init__(L, L.ref_upchan_create(), L.ref_upchan_destroy)

    def configure(self, output_rate, requested_rate, center_offset):
        rate, ofs = C.c_int32(0), C.c_int32(0)
        modes = np.zeros(32, dtype=np.int32)
        n = self.lib.ref_upchan_configure(self.h, output_rate, requested_rate, center_offset, C.byref(rate), C.byref(ofs), _p(modes, C.c_int32), 32)
        return rate.value, ofs.value, [int(m) for m in modes[:n]]

    def pull(self, source, n_out):
        x = np.ascontiguousarray(source, dtype=np.int16).reshape(-1, 2)
        out = np.empty((n_out, 2), dtype=np.int16)
        used = self.lib.ref_upchan_pull(self.h, _p(x, C.c_int16), x.shape[0], _p(out, C.c_int16), n_out)
        return out, used


def hb_coeffs(order, strict=False):
    a = np.zeros(24, dtype=np.int32)
    sh = C.c_int32(0)
    n = load(strict).ref_hb_coeffs(order, _p(a, C.c_int32), C.byref(sh))
    return a[:n].copy(), sh.value


class RefDemod(_Handle):
    """The reference's PhaseDiscriminators (kinds 0-2) and the AM magnitude lines of AMDemod::processOneSample (kind 3)."""

    def __init__(self, kind, fm_scaling=1.0, strict=False):
        L = load(strict)
        super().__init__(L, L.ref_discri_create(fm_scaling), L.ref_discri_destroy)
        self.kind = kind

    def run(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out, a0, a1 = (np.zeros(x.size, dtype=np.float32) for _ in range(3))
        self.lib.ref_discri_run(self.h, self.kind, _p(x.view(np.float32), C.c_float), x.size, _p(out, C.c_float), _p(a0, C.c_float), _p(a1, C.c_float))
        return out, a0, a1


def filerecord_write(path, sample_rate, center_frequency, iq, n1):
    """FileRecord: DSPSignalNotification, startRecording, feed(first n1), feed(rest), stopRecording.  Returns getByteCount()."""
    iq = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1, 2)
    return load().ref_filerecord_write(path.encode(), sample_rate, center_frequency, _p(iq, C.c_int16), n1, iq.shape[0] - n1)


def filerecord_read_header(path):
    r, c, t, s = C.c_int32(), C.c_ulonglong(), C.c_longlong(), C.c_uint()
    pos = load().ref_filerecord_read_header(path.encode(), C.byref(r), C.byref(c), C.byref(t), C.byref(s))
    return {"sample_rate": r.value, "center_frequency": c.value, "timestamp": t.value, "sample_size": s.value, "data_offset": pos}


class RefFftFilt(_Handle):
    """The reference's fftfilt (sdrbase/dsp/fftfilt.cpp over gfft.h), driven one sample per call like the demodulators do."""

    def __init__(self, kind, f1, f2, length, strict=False):
        L = load(strict)
        super().__init__(L, L.ref_fftfilt_create(kind, f1, f2, length), L.ref_fftfilt_destroy)
        self.flen = length

    def set_filter(self, kind, f1, f2):
        self.lib.ref_fftfilt_set(self.h, kind, f1, f2)

    def filter(self):
        out = np.zeros(self.flen, dtype=np.complex64)
        self.lib.ref_fftfilt_filter(self.h, _p(out.view(np.float32), C.c_float))
        return out

    def run(self, op, x, usb=True, get_dc=True):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.zeros(x.size + self.flen, dtype=np.complex64)
        m = self.lib.ref_fftfilt_run(self.h, op, (1 if usb else 0) | (2 if get_dc else 0), _p(x.view(np.float32), C.c_float), x.size, _p(out.view(np.float32), C.c_float), out.size)
        assert m >= 0
        return out[:m].copy()
