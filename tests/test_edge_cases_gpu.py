"""Edge cases through the C ABI on the GPU: empty and shorter-than-a-block inputs, one-sample feeds, stage-less channels,
partial spectrum frames, invalid arguments.  Every case is compared with the oracle (same calls, same order) or with the
error behaviour include/b200dsp.h documents; none may fall back, skip work silently or disturb the carried state."""
import ctypes as C

import numpy as np
import pytest

from conftest import MODES, rel_rms

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,bits", [("ii", 12), ("fi", 12), ("if", 16), ("u8", 8)])
def test_decimators_empty_and_sub_block_calls_leave_state_alone(gpu_lib, port, kind, bits):
    """The reference's loops simply do not run for len < one block (decimators.h:2862 `pos < len - 31`): no output, no
    state change.  Interleave such calls with real ones and compare every call with the oracle."""
    import sdrangel_b200 as S
    cls = {"ii": S.Decimators, "fi": S.DecimatorsFI, "if": S.DecimatorsIF, "u8": S.DecimatorsU}[kind]
    rs = np.random.RandomState(5)
    if kind == "u8":
        x = rs.randint(0, 256, size=40_000).astype(np.uint8)
        d, o = cls(), port.PortDecimators("u8")
    elif kind[0] == "i":
        x = rs.randint(-32768, 32768, size=40_000).astype(np.int16)
        d, o = cls(bits), port.PortDecimators(kind, bits)
    else:
        x = rs.uniform(-1, 1, size=40_000).astype(np.float32)
        d, o = cls(bits), port.PortDecimators(kind, bits)
    if kind in ("fi", "if"):
        d.set_exact_float(True)                     # the reference's rounding order: bit-identical to its strict build
    for log2, mname in ((4, "cen"), (6, "inf"), (3, "sup"), (0, "cen"), (1, "cen")):
        pos = 0
        for n in (0, 1, 7, 2, 4096, 0, 31, 10_000, 1, 3):
            got, want = d.run(log2, MODES[mname], x[pos:pos + n]), o.run(log2, MODES[mname], x[pos:pos + n])
            assert got.shape == want.shape, (log2, mname, n)
            if kind[1] == "i" or kind == "u8":
                assert np.array_equal(got, want), (log2, mname, n)
            else:
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (log2, mname, n)
            pos += n
    d.close()


def test_decimators_invalid_arguments_are_loud(gpu_lib):
    from sdrangel_b200 import Decimators, capi
    L = capi.lib()
    d = Decimators(12)
    buf = np.zeros(64, dtype=np.int16)
    out = np.zeros((64, 2), dtype=np.int16)
    n = C.c_int32(0)
    assert L.b200dsp_decim_run(d._h, 7, 2, buf.ctypes.data, 64, out.ctypes.data, C.byref(n)) == -1          # log2 > 6
    assert L.b200dsp_decim_run(d._h, 4, 3, buf.ctypes.data, 64, out.ctypes.data, C.byref(n)) == -1          # bad fc position
    assert L.b200dsp_decim_run(d._h, 4, 2, None, 64, out.ctypes.data, C.byref(n)) == -1                     # null input
    assert L.b200dsp_decim_run(None, 4, 2, buf.ctypes.data, 64, out.ctypes.data, C.byref(n)) == -1          # null handle
    assert b"null" in L.b200dsp_last_error()
    h = C.c_void_p()
    assert L.b200dsp_decim_create(C.byref(h), capi.FMT_I16, capi.FMT_I16, 10) == -1                          # input_bits
    assert L.b200dsp_decim_create(C.byref(h), capi.FMT_U8, capi.FMT_F32, 8) == -1                            # 8-bit -> float does not exist
    assert L.b200dsp_decim_out_count(capi.FMT_I16, capi.FMT_I16, 4, 2, -5) == -1
    d.close()


def test_bank_tiny_and_empty_feeds_vs_oracle(gpu_lib, port, golden_meta):
    """Feeds of 0, 1, 2, 3 ... samples: odd lengths leave a pending sample at every level (downchannelizer.cpp:63-75:
    a stage emits on every second sample it receives), empty feeds do nothing; per-call output counts and samples must
    equal the reference chains' for a shallow (S=3), a deep (S=7) and a stage-less (S=0) channel, front-ends included."""
    from sdrangel_b200 import DownChannelizerBank, capi
    fs = 10_000_000
    rs = np.random.RandomState(77)
    x = rs.randint(-30000, 30000, size=(60_000, 2)).astype(np.int16)
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    specs = [(1_000_000, 600_000), (48_000, 1_234_567), (48_000, -3_000_000), (20_000_000, 0)]    # (requested rate, offset)
    b = DownChannelizerBank(fs)
    chans, refs = [], []
    for req, fc in specs:
        cid, rate, ofs, path = b.add_channel(req, fc)
        o = port.PortDownChannelizer()
        assert o.configure(fs, req, fc)[:2] == (rate, ofs)
        fe = None
        if req == 48_000:
            b.set_frontend(cid, -ofs, cutoff, 48000)
            fe = port.PortFrontEnd(-ofs, rate, 48000, cutoff)
        chans.append((cid, path))
        refs.append((o, fe))
    assert chans[3][1] == ""                                  # requested >= input rate: no stage, samples forwarded unchanged
    pos = 0
    for n in (0, 1, 1, 2, 3, 0, 5, 127, 1, 128, 1000, 0, 7, 20_001, 2, 30_000):
        seg = x[pos:pos + n]
        b.feed(seg)
        for (cid, path), (o, fe) in zip(chans, refs):
            want = o.feed(seg)
            got = b.fetch(cid)
            assert got.shape == want.shape, (n, path)
            assert np.array_equal(got, want), (n, path)
            if fe is not None:
                wf, gf = fe.feed(want), b.fetch(cid, capi.STAGE_FRONTEND)
                assert gf.shape == wf.shape, (n, path)
                if wf.size:
                    assert np.max(np.abs(gf - wf)) <= 1e-5 * max(1.0, float(np.max(np.abs(wf)))) * 4, (n, path)
        pos += n
    b.close()


def test_spectrum_partial_frames_and_empty_feeds(gpu_lib, port):
    """SpectrumVis::feed keeps a partial FFT buffer between calls (spectrumvis.cpp:98-233): feeds shorter than a frame emit
    nothing until the frame completes; frame counts per call must match the reference glue, values within the K5 tolerance."""
    from sdrangel_b200 import SpectrumVis
    rs = np.random.RandomState(8)
    x = rs.randint(-2048, 2048, size=(1024 * 12, 2)).astype(np.int16)
    s, o = SpectrumVis(), port.PortSpectrumVis()
    s.configure(1024, 0, 0, SpectrumVis.AvgModeNone, 4, True)
    o.configure(1024, 0, 0, 0, 4, True)
    pos = 0
    for n in (0, 1, 1022, 1, 0, 500, 524, 3000, 72, 1024 * 2, 1):
        got, want = s.feed(x[pos:pos + n]), o.feed(x[pos:pos + n])
        assert got.shape[0] == want.shape[0], (n, got.shape, want.shape)
        if want.size:
            assert got.shape == want.shape and rel_rms(got, want) <= 1e-5, n
        pos += n
    s.close()


def test_interpolator_and_iqcorr_empty_inputs(gpu_lib, port):
    from sdrangel_b200 import Interpolator, IQCorrections
    it = Interpolator(16, 156250, 12500 / 2.2)
    out, rem = it.decimate(0.25, 3.25, np.zeros(0, dtype=np.complex64))
    assert out.size == 0 and rem == 0.25                      # no input: the caller's distance is untouched
    q, o = IQCorrections(), port.PortIQCorrections()
    rs = np.random.RandomState(3)
    x = (rs.randint(-100, 100, size=(5000, 2)) + 900).astype(np.int16)
    pos = 0
    for n in (0, 1, 0, 2, 1021, 1, 2000):
        got, want = q.iqCorrections(x[pos:pos + n].copy()), o.run(x[pos:pos + n])
        assert got.shape == want.shape and np.array_equal(got, want), n
        pos += n
