"""ctypes binding of libb200dsp.so (include/b200dsp.h).  Fails loudly when the library is not built."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200DSP_LIB") or os.path.join(_HERE, "lib", "libb200dsp.so")   # env override: kernel tuning experiments only

FMT_I16, FMT_F32, FMT_I8, FMT_U8 = 0, 1, 2, 3
MODE_INF, MODE_SUP, MODE_CEN, MODE_U = 0, 1, 2, 3
DECIM_STATE_ELEMS = 6 * 2 * 64
ENODEV = -2


class B200DspError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200dsp error %d: %s" % (code, msg))
        self.code = code


_lib = None

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_pvp, _pi32, _pi64 = C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol include/b200dsp.h declares
SIGNATURES = {
    "b200dsp_init": (_i32, [_i32]),
    "b200dsp_device_count": (_i32, []),
    "b200dsp_last_error": (C.c_char_p, []),
    "b200dsp_version": (C.c_char_p, []),
    "b200dsp_sm_count": (_i32, []),
    "b200dsp_decim_create": (_i32, [_pvp, _i32, _i32, _i32]),
    "b200dsp_decim_destroy": (_i32, [_vp]),
    "b200dsp_decim_set_shift": (_i32, [_vp, _i32]),
    "b200dsp_decim_set_exact_float": (_i32, [_vp, _i32]),
    "b200dsp_decim_run": (_i32, [_vp, _i32, _i32, _vp, _i32, _vp, _pi32]),
    "b200dsp_decim_run_split": (_i32, [_vp, _i32, _i32, _vp, _vp, _i32, _vp, _pi32]),
    "b200dsp_decim_run_dev": (_i32, [_vp, _i32, _i32, _vp, _i64, _vp, _pi64, _vp]),
    "b200dsp_decim_out_count": (_i64, [_i32, _i32, _i32, _i32, _i64]),
    "b200dsp_decim_get_state": (_i32, [_vp, _vp]),
    "b200dsp_decim_set_state": (_i32, [_vp, _vp]),
    "b200dsp_decim_reset": (_i32, [_vp]),
    "b200dsp_decim_sync": (_i32, [_vp]),
    "b200dsp_filter_chain": (_i32, [_i32, _i32, _i32, _pi32, _pi32, _pi32, _i32]),
    "b200dsp_bank_create": (_i32, [_pvp, _i32]),
    "b200dsp_bank_destroy": (_i32, [_vp]),
    "b200dsp_bank_set_chunk": (_i32, [_vp, _i64]),
    "b200dsp_bank_add_channel": (_i32, [_vp, _i32, _i32, _pi32, _pi32, _pi32]),
    "b200dsp_bank_add_channel_path": (_i32, [_vp, _pi32, _i32, _i32, _pi32]),
    "b200dsp_bank_reset": (_i32, [_vp, _vp]),
    "b200dsp_bank_channel_path": (_i32, [_vp, _i32, _pi32, _i32]),
    "b200dsp_bank_node_count": (_i32, [_vp]),
    "b200dsp_bank_set_frontend": (_i32, [_vp, _i32, _f32, _i32, C.c_double, C.c_double, _i32]),
    "b200dsp_bank_frontend_info": (_i32, [_vp, _i32, _pi32, _pi32, _vp, _i32]),
    "b200dsp_bank_feed": (_i32, [_vp, _vp, _i64]),
    "b200dsp_bank_feed_dev": (_i32, [_vp, _vp, _i64, _vp]),
    "b200dsp_bank_fetch": (_i32, [_vp, _i32, _i32, _vp, _i64, _pi64]),
    "b200dsp_bank_fetch_schedule": (_i32, [_vp, _i32, _vp, _vp, _i64, _pi64]),
    "b200dsp_bank_fetch_dev": (_i32, [_vp, _i32, _i32, _pvp, _pi64]),
    "b200dsp_bank_fetch_all": (_i32, [_vp, _i32, _vp, _i64, _vp, _vp]),
    "b200dsp_bank_gather_dev": (_i32, [_vp, _i32, _vp, _i64, _vp, _vp]),
    "b200dsp_bank_process": (_i32, [_vp, _vp, _i64, _i32, _vp, _i64, _vp]),
    "b200dsp_bank_stream": (_vp, [_vp]),
    "b200dsp_dist_shard": (_i32, [_i32, _i32, _i32, _pi32, _pi32]),
    "b200dsp_dist_unique_id": (_i32, [_vp]),
    "b200dsp_dist_create": (_i32, [_pvp, _vp, _i32, _i32]),
    "b200dsp_dist_destroy": (_i32, [_vp]),
    "b200dsp_dist_reserve": (_i32, [_vp, _i64]),
    "b200dsp_dist_bcast_begin": (_i32, [_vp, _i32, _vp, _i64, _i32, _vp]),
    "b200dsp_dist_ingest_begin": (_i32, [_vp, _i32, _vp, _i64]),
    "b200dsp_dist_feed": (_i32, [_vp, _i32, _vp, _vp]),
    "b200dsp_dist_slot": (_i32, [_vp, _i32, _pvp, _pi64]),
    "b200dsp_dist_sync": (_i32, [_vp]),
    "b200dsp_dist_p2p_export": (_i32, [_vp, _i64, _vp]),
    "b200dsp_dist_p2p_import": (_i32, [_vp, _vp]),
    "b200dsp_dist_p2p_begin": (_i32, [_vp, _i32, _vp, _i64, _vp]),
    "b200dsp_dist_p2p_feed": (_i32, [_vp, _i32, _vp, _vp]),
    "b200dsp_dist_p2p_slot": (_i32, [_vp, _i32, _pvp, _pi64]),
    "b200dsp_interp_interpolate": (_i32, [_vp, C.POINTER(_f32), _f32, _vp, _i64, _vp, _i64, _pi64]),
    "b200dsp_interp_resample": (_i32, [_vp, C.POINTER(_f32), _f32, _vp, _i64, _vp, _i64, _pi64]),
    "b200dsp_interp_step": (_i32, [_vp, _i32, C.POINTER(_f32), _vp, _vp, _pi32, _pi32]),
    "b200dsp_fifo_create": (_i32, [_pvp, C.c_uint32]),
    "b200dsp_fifo_destroy": (_i32, [_vp]),
    "b200dsp_fifo_size": (C.c_uint32, [_vp]),
    "b200dsp_fifo_fill": (C.c_uint32, [_vp]),
    "b200dsp_fifo_write": (_i32, [_vp, _vp, C.c_uint32, _i32, _vp, C.POINTER(C.c_uint32)]),
    "b200dsp_fifo_read_begin": (_i32, [_vp, C.c_uint32, _pvp, C.POINTER(C.c_uint32), _pvp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "b200dsp_fifo_read_commit": (_i32, [_vp, C.c_uint32, C.POINTER(C.c_uint32)]),
    "b200dsp_fifo_read": (_i32, [_vp, _vp, C.c_uint32, _vp, C.POINTER(C.c_uint32)]),
    "b200dsp_nco_create": (_i32, [_pvp]),
    "b200dsp_nco_destroy": (_i32, [_vp]),
    "b200dsp_nco_set_freq": (_i32, [_vp, _f32, _f32]),
    "b200dsp_nco_set_phase": (_i32, [_vp, _i32]),
    "b200dsp_nco_get": (_i32, [_vp, _pi32, _pi32]),
    "b200dsp_nco_next_iq": (_i32, [_vp, _i64, _vp]),
    "b200dsp_nco_next_iq_dev": (_i32, [_vp, _i64, _vp, _vp]),
    "b200dsp_nco_mix_dev": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "b200dsp_bank_copy_out_dev": (_i32, [_vp, _i32, _i64, _i64, _vp, _vp]),
    "b200dsp_bank_sync": (_i32, [_vp]),
    "b200dsp_bank_tree_time": (_i32, [_vp, _vp, _vp]),
    "b200dsp_iqcorr_create": (_i32, [_pvp]),
    "b200dsp_iqcorr_destroy": (_i32, [_vp]),
    "b200dsp_iqcorr_reset": (_i32, [_vp]),
    "b200dsp_iqcorr_run": (_i32, [_vp, _vp, _i64, _i32]),
    "b200dsp_iqcorr_run_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "b200dsp_interp_create": (_i32, [_pvp, _i32, C.c_double, C.c_double, C.c_double]),
    "b200dsp_interp_destroy": (_i32, [_vp]),
    "b200dsp_interp_info": (_i32, [_vp, _pi32, _vp, _i32]),
    "b200dsp_interp_decimate": (_i32, [_vp, C.POINTER(C.c_float), _f32, _vp, _i64, _vp, _i64, _pi64]),
    "b200dsp_interps_create": (_i32, [_pvp, _i32, _i32]),
    "b200dsp_interps_destroy": (_i32, [_vp]),
    "b200dsp_interps_reset": (_i32, [_vp]),
    "b200dsp_interps_in_count": (_i64, [_i32, _i64]),
    "b200dsp_interps_run": (_i32, [_vp, _i32, _vp, _vp, _i32, _pi32]),
    "b200dsp_interps_run_dev": (_i32, [_vp, _i32, _vp, _vp, _i64, _pi64, _vp]),
    "b200dsp_upchan_create": (_i32, [_pvp]),
    "b200dsp_upchan_destroy": (_i32, [_vp]),
    "b200dsp_upchan_configure": (_i32, [_vp, _i32, _i32, _i32, _pi32, _pi32]),
    "b200dsp_upchan_set_path": (_i32, [_vp, _pi32, _i32]),
    "b200dsp_upchan_path": (_i32, [_vp, _pi32, _i32]),
    "b200dsp_upchan_source_count": (_i64, [_vp, _i64]),
    "b200dsp_upchan_pull": (_i32, [_vp, _vp, _i64, _vp, _i64]),
    "b200dsp_upchan_pull_dev": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp]),
    "b200dsp_demod_create": (_i32, [_pvp, _i32, _f32, _i32]),
    "b200dsp_demod_destroy": (_i32, [_vp]),
    "b200dsp_demod_reset": (_i32, [_vp]),
    "b200dsp_demod_set_fm_scaling": (_i32, [_vp, _f32]),
    "b200dsp_demod_run": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "b200dsp_demod_run_pool_dev": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp, _i64, _vp, _vp, _vp]),
    "b200dsp_fftfilt_create": (_i32, [_pvp, _i32, _f32, _f32, _i32]),
    "b200dsp_fftfilt_destroy": (_i32, [_vp]),
    "b200dsp_fftfilt_set_filter": (_i32, [_vp, _i32, _f32, _f32]),
    "b200dsp_fftfilt_filter": (_i32, [_vp, _vp, _i32]),
    "b200dsp_fftfilt_out_count": (_i64, [_vp, _i64]),
    "b200dsp_fftfilt_run": (_i32, [_vp, _i32, _i32, _i32, _vp, _i64, _vp, _i64, _pi64]),
    "b200dsp_fftfilt_run_dev": (_i32, [_vp, _i32, _i32, _i32, _vp, _i64, _vp, _i64, _pi64, _vp]),
    "b200dsp_sdriq_header_encode": (_i32, [_i32, C.c_uint64, _i64, C.c_uint32, _vp]),
    "b200dsp_sdriq_header_decode": (_i32, [_vp, _pi32, C.POINTER(C.c_uint64), _pi64, C.POINTER(C.c_uint32)]),
    "b200dsp_sdriq_open": (_i32, [_pvp, C.c_char_p, _pi32, C.POINTER(C.c_uint64), _pi64, C.POINTER(C.c_uint32), _pi64]),
    "b200dsp_sdriq_read": (_i32, [_vp, _vp, _i64, _pi64]),
    "b200dsp_sdriq_create": (_i32, [_pvp, C.c_char_p, _i32, C.c_uint64, _i64]),
    "b200dsp_sdriq_write": (_i32, [_vp, _vp, _i64]),
    "b200dsp_sdriq_close": (_i32, [_vp]),
    "b200dsp_spectrum_create": (_i32, [_pvp, _f32]),
    "b200dsp_spectrum_destroy": (_i32, [_vp]),
    "b200dsp_spectrum_configure": (_i32, [_vp, _i32, _i32, C.c_uint, _i32, _i32, _i32]),
    "b200dsp_spectrum_frames_for": (_i64, [_vp, _i64]),
    "b200dsp_spectrum_feed": (_i32, [_vp, _vp, _i64, _i32, _vp, _i64, _pi64]),
    "b200dsp_spectrum_feed_dev": (_i32, [_vp, _vp, _i64, _i32, _vp, _i64, _pi64, _vp]),
}
STAGE_CHANNELIZER, STAGE_FRONTEND = 0, 1


def lib():
    """The loaded library.  Raises ImportError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libb200dsp.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C sdrangel_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise B200DspError(rc, lib().b200dsp_last_error().decode("utf-8", "replace"))
    return rc


def device_count():
    return lib().b200dsp_device_count()


def init(device=0):
    check(lib().b200dsp_init(device))
