"""ORACLE (test infrastructure): ctypes binding of oracle/_build/liboracle_port.so, the plain-C restatement
(oracle/port/sdr_oracle.c).  Same class surface as oracle/refbind.py so tests can run either checker.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MODE_INF, MODE_SUP, MODE_CEN = 0, 1, 2
FMT_I16, FMT_F32 = 0, 1
_lib = None


def lib_path():
    return os.path.join(_HERE, "_build", "liboracle_port.so")


def build():
    subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(_HERE, "port", "sdr_oracle.c")
    if not os.path.exists(lib_path()) or os.path.getmtime(lib_path()) < os.path.getmtime(src):
        build()
    L = C.CDLL(lib_path())
    vp, i32, f32, f64 = C.c_void_p, C.c_int, C.c_float, C.c_double
    pi16, pf32, pi32 = C.POINTER(C.c_int16), C.POINTER(C.c_float), C.POINTER(C.c_int32)
    sig = {
        "orc_sdrbench_gen_s16": (None, [pi16, i32]), "orc_sdrbench_gen_f32": (None, [pf32, i32]),
        "orc_decim_ii_create": (vp, [i32]), "orc_decim_ii_destroy": (None, [vp]),
        "orc_decim_ii_run": (i32, [vp, i32, i32, pi16, i32, pi16]),
        "orc_iqcorr_create": (vp, []), "orc_iqcorr_destroy": (None, [vp]), "orc_iqcorr_dc": (None, [vp, pi16, i32]), "orc_iqcorr_imbalance": (None, [vp, pi16, i32]),
        "orc_decim_x8_create": (vp, [i32, i32]), "orc_decim_x8_destroy": (None, [vp]),
        "orc_decim_x8_run": (i32, [vp, i32, i32, vp, i32, pi16]),
        "orc_decim_f_create": (vp, [i32, i32, i32]), "orc_decim_f_destroy": (None, [vp]),
        "orc_decim_f_run": (i32, [vp, i32, i32, vp, i32, vp]),
        "orc_chan_create": (vp, []), "orc_chan_destroy": (None, [vp]),
        "orc_chan_configure": (i32, [vp, i32, i32, i32, pi32, pi32, pi32, i32]),
        "orc_chan_feed": (i32, [vp, pi16, i32, pi16, i32]),
        "orc_nco_table": (None, [pf32]), "orc_nco_increment": (i32, [f32, f32]),
        "orc_interp_ntaps": (i32, [i32, f64]), "orc_interp_taps": (None, [i32, f64, f64, f64, pf32]),
        "orc_frontend_create": (vp, [f32, f32, i32, f64, f64, f64, f32]), "orc_frontend_destroy": (None, [vp]),
        "orc_frontend_feed": (i32, [vp, pi16, i32, pf32, i32, pi32, pi32]),
        "orc_interp_run": (i32, [vp, i32, pf32, i32, pf32, i32]), "orc_frontend_remain": (f32, [vp]),
        "orc_nco_block": (None, [f32, f32, i32, pf32]),
        "orc_fft_window": (None, [i32, i32, pf32]), "orc_kissfft_forward": (None, [i32, pf32, pf32]),
        "orc_spectrum_create": (vp, [f32]), "orc_spectrum_destroy": (None, [vp]),
        "orc_spectrum_configure": (None, [vp, i32, i32, C.c_uint, i32, i32, i32]),
        "orc_spectrum_feed": (i32, [vp, pi16, i32, i32, pf32, i32]),
        "orc_interps_create": (vp, [i32]), "orc_interps_destroy": (None, [vp]), "orc_interps_run": (i32, [vp, i32, pi16, vp, i32]),
        "orc_upchan_create": (vp, []), "orc_upchan_destroy": (None, [vp]),
        "orc_upchan_configure": (i32, [vp, i32, i32, i32, pi32, pi32, pi32, i32]),
        "orc_upchan_pull": (i32, [vp, pi16, i32, pi16, i32]), "orc_hb_interp_coeffs": (None, [i32, pi32]),
        "orc_discri_create": (vp, [f32]), "orc_discri_destroy": (None, [vp]), "orc_discri_run": (None, [vp, i32, pf32, i32, pf32, pf32, pf32]),
        "orc_fftfilt_create": (vp, [i32, f32, f32, i32]), "orc_fftfilt_destroy": (None, [vp]), "orc_fftfilt_set": (None, [vp, i32, f32, f32]),
        "orc_fftfilt_filter": (None, [vp, pf32]), "orc_fftfilt_run": (i32, [vp, i32, i32, pf32, i32, pf32, i32]),
        "orc_sdriq_header": (None, [C.c_int32, C.c_uint64, C.c_int64, C.c_uint32, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


class _Handle:
    def __init__(self, h, destroy):
        self.h, self._destroy = h, destroy

    def __del__(self):
        try:
            if self.h:
                self._destroy(self.h)
                self.h = None
        except Exception:
            pass


class PortDecimators(_Handle):
    def __init__(self, kind="ii", input_bits=12, shift=127):
        L = load()
        self.kind = kind
        self.in_dt = {"i8": np.int8, "u8": np.uint8}.get(kind, np.int16 if kind[0] == "i" else np.float32)
        self.out_dt = np.int16 if kind[1] in "i8" else np.float32
        if kind in ("i8", "u8"):            # input_bits is 8 by construction; `shift` only matters for 'u8'
            super().__init__(L.orc_decim_x8_create(int(kind == "u8"), shift), L.orc_decim_x8_destroy)
        elif kind == "ii":
            super().__init__(L.orc_decim_ii_create(input_bits), L.orc_decim_ii_destroy)
        else:
            super().__init__(L.orc_decim_f_create(FMT_I16 if kind[0] == "i" else FMT_F32,
                                                  FMT_I16 if kind[1] == "i" else FMT_F32, input_bits),
                             L.orc_decim_f_destroy)

    def run(self, log2, mode, buf):
        L = load()
        buf = np.ascontiguousarray(buf, dtype=self.in_dt)
        out = np.empty((buf.size // 2 + 8, 2), dtype=self.out_dt)
        if self.kind in ("i8", "u8"):
            n = L.orc_decim_x8_run(self.h, log2, mode, buf.ctypes.data, buf.size, _p(out, C.c_int16))
        elif self.kind == "ii":
            n = L.orc_decim_ii_run(self.h, log2, mode, _p(buf, C.c_int16), buf.size, _p(out, C.c_int16))
        else:
            n = L.orc_decim_f_run(self.h, log2, mode, buf.ctypes.data, buf.size, out.ctypes.data)
        if n < 0:
            raise ValueError("bad log2/mode")
        return out[:n].copy()


class PortIQCorrections(_Handle):
    def __init__(self):
        L = load()
        super().__init__(L.orc_iqcorr_create(), L.orc_iqcorr_destroy)

    def run(self, iq, imbalance=False):
        a = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1, 2).copy()
        (load().orc_iqcorr_imbalance if imbalance else load().orc_iqcorr_dc)(self.h, _p(a, C.c_int16), a.shape[0])
        return a


class PortDownChannelizer(_Handle):
    def __init__(self):
        L = load()
        super().__init__(L.orc_chan_create(), L.orc_chan_destroy)

    def configure(self, input_rate, requested_rate, center_offset):
        rate, ofs = C.c_int32(0), C.c_int32(0)
        modes = np.zeros(32, dtype=np.int32)
        n = load().orc_chan_configure(self.h, input_rate, requested_rate, center_offset,
                                      C.byref(rate), C.byref(ofs), _p(modes, C.c_int32), 32)
        return rate.value, ofs.value, [int(m) for m in modes[:n]]

    def feed(self, iq):
        iq = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1, 2)
        out = np.empty((iq.shape[0] + 8, 2), dtype=np.int16)
        n = load().orc_chan_feed(self.h, _p(iq, C.c_int16), iq.shape[0], _p(out, C.c_int16), out.shape[0])
        assert n >= 0
        return out[:n].copy()


class PortFrontEnd(_Handle):
    def __init__(self, nco_freq, rate, out_rate, cutoff, phase_steps=16, taps_per_phase=4.5):
        L = load()
        distance = np.float32(np.float32(rate) / np.float32(out_rate))
        self.args = (phase_steps, float(rate), float(cutoff), float(taps_per_phase))
        self.nco = (float(nco_freq), float(rate))
        h = L.orc_frontend_create(float(nco_freq), float(rate), phase_steps, float(rate), float(cutoff),
                                  float(taps_per_phase), float(distance))
        super().__init__(h, L.orc_frontend_destroy)
        self.phase_steps = phase_steps

    def nco_increment(self):
        return load().orc_nco_increment(*self.nco)

    def taps(self):
        L = load()
        nt = L.orc_interp_ntaps(self.phase_steps, self.args[3])
        t = np.empty(nt * self.phase_steps, dtype=np.float32)
        L.orc_interp_taps(self.phase_steps, self.args[1], self.args[2], self.args[3], _p(t, C.c_float))
        return t.reshape(self.phase_steps, nt)

    def feed(self, iq, want_schedule=False):
        iq = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1, 2)
        n = iq.shape[0]
        out = np.empty((n + 8, 2), dtype=np.float32)
        idx = np.empty(n + 8, dtype=np.int32)
        ph = np.empty(n + 8, dtype=np.int32)
        m = load().orc_frontend_feed(self.h, _p(iq, C.c_int16), n, _p(out, C.c_float), n + 8,
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
        m = load().orc_interp_run(self.h, int(mode), _p(x.view(np.float32), C.c_float), n, _p(out, C.c_float), cap)
        assert m >= 0
        return out[:m].copy().view(np.complex64).reshape(-1)

    def remain(self):
        return float(load().orc_frontend_remain(self.h))


def nco_table():
    t = np.empty(4096, dtype=np.float32)
    load().orc_nco_table(_p(t, C.c_float))
    return t


def nco_block(freq, rate, n):
    """NCO::setFreq(freq, rate) then n x nextIQ() as complex64."""
    out = np.empty((n, 2), dtype=np.float32)
    load().orc_nco_block(float(freq), float(rate), int(n), _p(out, C.c_float))
    return out.view(np.complex64).reshape(-1)


def fft_window(function, n):
    w = np.empty(n, dtype=np.float32)
    load().orc_fft_window(function, n, _p(w, C.c_float))
    return w


def kissfft(x):
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.empty(x.size, dtype=np.complex64)
    load().orc_kissfft_forward(x.size, x.view(np.float32).ctypes.data_as(C.POINTER(C.c_float)),
                               out.view(np.float32).ctypes.data_as(C.POINTER(C.c_float)))
    return out


class PortSpectrumVis(_Handle):
    AVG_NONE, AVG_MOVING, AVG_FIXED = 0, 1, 2

    def __init__(self, scalef=32768.0):
        L = load()
        super().__init__(L.orc_spectrum_create(scalef), L.orc_spectrum_destroy)
        self.fft_size = 1024

    def configure(self, fft_size, overlap_pct=0, avg_nb=0, avg_mode=0, window=1, linear=False):
        load().orc_spectrum_configure(self.h, fft_size, overlap_pct, avg_nb, avg_mode, window, int(linear))
        self.fft_size = min(max(fft_size, 64), 4096)

    def feed(self, iq, positive_only=False):
        iq = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1, 2)
        cap = iq.shape[0] // self.fft_size + 2
        frames = np.empty((cap, self.fft_size), dtype=np.float32)
        n = load().orc_spectrum_feed(self.h, _p(iq, C.c_int16), iq.shape[0], int(positive_only),
                                     _p(frames, C.c_float), cap)
        assert n >= 0
        return frames[:n].copy()


def sdrbench_s16(n_samples):
    buf = np.empty(2 * n_samples, dtype=np.int16)
    load().orc_sdrbench_gen_s16(_p(buf, C.c_int16), buf.size)
    return buf


def sdrbench_f32(n_samples):
    buf = np.empty(2 * n_samples, dtype=np.float32)
    load().orc_sdrbench_gen_f32(_p(buf, C.c_float), buf.size)
    return buf


class PortInterpolators(_Handle):
    """Interpolators<T,16,output_bits> restated (oracle/port/sdr_oracle.c: orc_interps_*)."""

    def __init__(self, output_bits=16):
        L = load()
        super().__init__(L.orc_interps_create(output_bits), L.orc_interps_destroy)
        self.dtype = np.int8 if output_bits == 8 else np.int16

    def run(self, log2, samples, length=None, fill=0):
        x = np.ascontiguousarray(samples, dtype=np.int16).reshape(-1, 2)
        if length is None:
            length = x.shape[0] * (2 << log2)
        assert length // (2 << log2) <= x.shape[0]
        buf = np.full(int(length), fill, dtype=self.dtype)
        n = load().orc_interps_run(self.h, log2, _p(x, C.c_int16), buf.ctypes.data, int(length))
        return buf, n


class PortUpChannelizer(_Handle):
    def __init__(self):
        L = load()
        super().__init__(L.orc_upchan_create(), L.orc_upchan_destroy)

    def configure(self, output_rate, requested_rate, center_offset):
        rate, ofs = C.c_int32(0), C.c_int32(0)
        modes = np.zeros(32, dtype=np.int32)
        n = load().orc_upchan_configure(self.h, output_rate, requested_rate, center_offset, C.byref(rate), C.byref(ofs), _p(modes, C.c_int32), 32)
        return rate.value, ofs.value, [int(m) for m in modes[:n]]

    def pull(self, source, n_out):
        x = np.ascontiguousarray(source, dtype=np.int16).reshape(-1, 2)
        out = np.empty((n_out, 2), dtype=np.int16)
        used = load().orc_upchan_pull(self.h, _p(x, C.c_int16), x.shape[0], _p(out, C.c_int16), n_out)
        return out, used


def hb_interp_coeffs(order):
    a = np.zeros(order // 4, dtype=np.int32)
    load().orc_hb_interp_coeffs(order, _p(a, C.c_int32))
    return a


class PortDemod(_Handle):
    """PhaseDiscriminators / AM magnitude restated (orc_discri_*): kind 0 atan2, 1 delta, 2 discri2, 3 AM magnitude."""

    def __init__(self, kind, fm_scaling=1.0):
        L = load()
        super().__init__(L.orc_discri_create(fm_scaling), L.orc_discri_destroy)
        self.kind = kind

    def run(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out, a0, a1 = (np.zeros(x.size, dtype=np.float32) for _ in range(3))
        load().orc_discri_run(self.h, self.kind, _p(x.view(np.float32), C.c_float), x.size, _p(out, C.c_float), _p(a0, C.c_float), _p(a1, C.c_float))
        return out, a0, a1


def sdriq_header(rate, center, ts, sample_size=16):
    b = np.zeros(24, dtype=np.uint8)
    load().orc_sdriq_header(rate, center, ts, sample_size, b.ctypes.data)
    return b.tobytes()


class PortFftFilt(_Handle):
    """fftfilt restated (orc_fftfilt_*): kind 0 fftfilt(f1, f2, len), 1 fftfilt(f2, len); op 0 runFilt, 1 runSSB, 2 runDSB."""

    def __init__(self, kind, f1, f2, length):
        L = load()
        super().__init__(L.orc_fftfilt_create(kind, f1, f2, length), L.orc_fftfilt_destroy)
        self.flen = length

    def set_filter(self, kind, f1, f2):
        load().orc_fftfilt_set(self.h, kind, f1, f2)

    def filter(self):
        out = np.zeros(self.flen, dtype=np.complex64)
        load().orc_fftfilt_filter(self.h, _p(out.view(np.float32), C.c_float))
        return out

    def run(self, op, x, usb=True, get_dc=True):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.zeros(x.size + self.flen, dtype=np.complex64)
        m = load().orc_fftfilt_run(self.h, op, (1 if usb else 0) | (2 if get_dc else 0), _p(x.view(np.float32), C.c_float), x.size, _p(out.view(np.float32), C.c_float), out.size)
        assert m >= 0
        return out[:m].copy()
