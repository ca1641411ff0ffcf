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
