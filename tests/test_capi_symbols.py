"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol include/*.h
declares, reports errors (never falls back) when no device is present, and the pure-host block arithmetic matches
the reference's loop bounds."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

from conftest import ROOT, MODES


def _declared_symbols():
    syms = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        syms |= set(re.findall(r"\b(b200dsp_[a-z0-9_]+)\s*\(", src))
    return sorted(syms)


def test_library_loads_and_exports_every_declared_symbol():
    from sdrangel_b200 import capi
    L = capi.lib()
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), "libb200dsp.so does not export %s" % s
    assert set(syms) == set(capi.SIGNATURES), "ctypes SIGNATURES and include/b200dsp.h disagree"
    assert b"sm_100a" in L.b200dsp_version()


def test_no_cpu_fallback_without_device():
    from sdrangel_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = capi.lib().b200dsp_decim_create(C.byref(h), 0, 0, 12)
    assert rc == capi.ENODEV and not h.value
    assert b"no CPU fallback" in capi.lib().b200dsp_last_error()
    with pytest.raises(capi.B200DspError):
        import sdrangel_b200
        sdrangel_b200.Decimators(12)


def test_out_count_matches_oracle_block_rounding(port):
    from sdrangel_b200 import capi
    L = capi.lib()
    rs = np.random.RandomState(3)
    lens = [0, 1, 2, 3, 7, 8, 15, 16, 31, 32, 127, 128, 129, 255, 256, 257, 1000, 4097] + rs.randint(0, 20000, 20).tolist()
    for kind, fi, fo in (("ii", 0, 0), ("fi", 1, 0), ("ff", 1, 1), ("if", 0, 1)):
        dt = np.int16 if kind[0] == "i" else np.float32
        for log2 in range(7):
            for mode in MODES.values():
                for n in lens:
                    want = port.PortDecimators(kind, 12).run(log2, mode, np.zeros(n, dtype=dt)).shape[0]
                    got = L.b200dsp_decim_out_count(fi, fo, log2, mode, n)
                    assert got == want, (kind, log2, mode, n, got, want)
    assert L.b200dsp_decim_out_count(0, 0, 7, 0, 100) == -1
    assert L.b200dsp_decim_out_count(0, 0, 3, 5, 100) == -1


def test_filter_chain_host_arithmetic_matches_reference_plans(golden_meta):
    """b200dsp_filter_chain is pure host arithmetic (createFilterChain's float32/double mix, downchannelizer.cpp:250-287): every
    golden plan of the reference -- the 64- and 1024-channel plans and 200 random (rate, request, offset) triples -- without
    a device."""
    from sdrangel_b200.sharding import filter_chain
    plans = golden_meta["chan_plans"]
    n = 0
    for name in ("bank64", "bank1024"):
        fs = plans[name]["input_rate"]
        for fc, rate, ofs, path in plans[name]["channels"]:
            assert filter_chain(fs, 48000, fc) == (rate, ofs, path), (name, fc)
            n += 1
    for fs, req, fc, rate, ofs, path in plans["random"]:
        assert filter_chain(fs, req, fc) == (rate, ofs, path), (fs, req, fc)
        n += 1
    assert n >= 1288


def test_out_count_8bit_formats_and_bad_arguments(port):
    from sdrangel_b200 import capi
    L = capi.lib()
    for kind, fmt, dt in (("i8", capi.FMT_I8, np.int8), ("u8", capi.FMT_U8, np.uint8)):
        for log2 in range(7):
            for mode in MODES.values():
                for n in (0, 1, 7, 8, 31, 32, 255, 256, 257, 1000, 4097, 12345):
                    want = port.PortDecimators(kind).run(log2, mode, np.zeros(n, dtype=dt)).shape[0]
                    assert L.b200dsp_decim_out_count(fmt, capi.FMT_I16, log2, mode, n) == want, (kind, log2, mode, n)
        assert L.b200dsp_decim_out_count(fmt, capi.FMT_F32, 4, 2, 1024) == -1        # 8-bit -> float does not exist in the reference
    assert L.b200dsp_decim_out_count(0, 0, 7, 2, 1024) == -1 and L.b200dsp_decim_out_count(0, 0, 4, 5, 1024) == -1
    assert L.b200dsp_decim_out_count(9, 0, 4, 2, 1024) == -1


def test_dist_shard_is_a_partition_and_matches_the_python_helper():
    """b200dsp_dist_shard (pure host arithmetic): contiguous, disjoint, complete, sizes within one of each other."""
    import ctypes as C
    from sdrangel_b200 import capi
    from sdrangel_b200.sharding import shard_channels
    L = capi.lib()
    for n in (0, 1, 7, 64, 1024, 1031):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = C.c_int32(), C.c_int32()
                assert L.b200dsp_dist_shard(n, world, r, C.byref(lo), C.byref(hi)) == 0
                assert (lo.value, hi.value) == shard_channels(n, world, r)
                assert lo.value == prev and hi.value - lo.value in (n // world, n // world + 1)
                prev = hi.value
            assert prev == n
    lo, hi = C.c_int32(), C.c_int32()
    assert L.b200dsp_dist_shard(8, 2, 2, C.byref(lo), C.byref(hi)) < 0
