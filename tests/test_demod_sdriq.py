"""demodulator back-ends (PhaseDiscriminators, AM magnitude) and the .sdriq record format.
CPU: the C oracle against the reference's golden vectors (bit-identical to the build without -ffast-math, <= 1e-6 of the
-ffast-math build), the host-side .sdriq code of the C ABI against a file the reference's FileRecord wrote.
GPU: b200dsp_demod_* against the goldens and the oracle, single stream and the pooled bank layout."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR


@pytest.fixture(scope="module")
def gdm():
    z = np.load(os.path.join(GOLDEN_DIR, "golden_demod.npz"))
    with open(os.path.join(GOLDEN_DIR, "golden_demod.json")) as f:
        return {k: z[k] for k in z.files}, json.load(f)


def demod_input(seed, n):
    rs = np.random.RandomState(seed)
    x = ((rs.randn(n) + 1j * rs.randn(n)) * 8000).astype(np.complex64)
    x[100] = 0; x[200] = 5 + 0j; x[201] = 0 + 7j; x[202] = -3 + 0j; x[203] = 0 - 2j; x[204] = 4 + 4j; x[205] = -4 + 4j; x[300:310] = 0
    return x


def _check_demod(make, gdm, exact_kinds):
    arrays, meta = gdm
    cuts = meta["cuts"]
    x = demod_input(meta["seed"], cuts[-1])
    for kind in range(4):
        d = make(kind, meta["scaling"])
        outs = [d.run(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
        for j, name in enumerate(("out", "aux0", "aux1")):
            got = np.concatenate([o[j] for o in outs])
            strict, fast = arrays["demod/strict/%d/%s" % (kind, name)], arrays["demod/fast/%d/%s" % (kind, name)]
            if kind in exact_kinds:
                assert np.array_equal(got, strict), (kind, name)
            else:       # atan2f of another libm: <= 2 ulp of a value in [-pi, pi], scaled by fm_scaling / pi
                assert np.max(np.abs(got - strict)) <= 1e-6, (kind, name)
            scale = max(1.0, float(np.max(np.abs(fast))))
            assert np.max(np.abs(got - fast)) <= 1e-6 * scale, (kind, name)


def test_port_demod_equals_strict_reference_goldens(port, gdm):
    _check_demod(port.PortDemod, gdm, exact_kinds=(0, 1, 2, 3))


def test_sdriq_host_code_against_reference_file(port, gdm, tmp_path):
    """Header layout, lazy header, reader: the bytes FileRecord wrote (timestamp masked) and what FileRecord::readHeader read."""
    from sdrangel_b200 import SdriqFile, capi
    arrays, meta = gdm
    m = meta["sdriq"]
    raw = arrays["sdriq/file"].tobytes()
    iq = np.random.RandomState(m["seed"]).randint(-32768, 32768, size=(m["n_samples"], 2)).astype(np.int16)
    assert len(raw) == 24 + 4 * m["n_samples"] and raw[24:] == iq.tobytes()
    assert m["header_read"] == {"sample_rate": m["sample_rate"], "center_frequency": m["center_frequency"], "sample_size": 16, "data_offset": 24}
    # encode == the reference's header (and the oracle's), decode round trip, garbage sample size -> 16 (filerecord.cpp:145-147)
    hdr = (C.c_ubyte * 24)()
    assert capi.lib().b200dsp_sdriq_header_encode(m["sample_rate"], m["center_frequency"], 0, 16, hdr) == 0
    assert bytes(hdr) == raw[:24] == port.sdriq_header(m["sample_rate"], m["center_frequency"], 0)
    bad = bytearray(raw[:24]); bad[20:24] = (77).to_bytes(4, "little")
    r, c, t, s = C.c_int32(), C.c_uint64(), C.c_int64(), C.c_uint32()
    assert capi.lib().b200dsp_sdriq_header_decode(bytes(bad), C.byref(r), C.byref(c), C.byref(t), C.byref(s)) == 0
    assert (r.value, c.value, s.value) == (m["sample_rate"], m["center_frequency"], 16)
    # writer: two feeds like the golden recording, an empty feed first (no header yet), same bytes
    p = str(tmp_path / "w.sdriq")
    w = SdriqFile(p, "w", m["sample_rate"], m["center_frequency"], 0)
    w.write(iq[:0])
    assert os.path.getsize(p) == 0 or open(p, "rb").read() == b""
    w.write(iq[:m["first_feed"]]); w.write(iq[m["first_feed"]:]); w.close()
    assert open(p, "rb").read() == raw
    # reader on the reference's file
    g = str(tmp_path / "g.sdriq")
    open(g, "wb").write(raw)
    f = SdriqFile(g)
    assert (f.sample_rate, f.center_frequency, f.sample_size, f.n_samples) == (m["sample_rate"], m["center_frequency"], 16, m["n_samples"])
    a, b, c2 = f.read(300), f.read(5000), f.read(10)
    assert np.array_equal(np.concatenate([a, b]), iq) and c2.shape[0] == 0
    f.close()
    h = C.c_void_p()
    assert capi.lib().b200dsp_sdriq_open(C.byref(h), b"/nonexistent/x.sdriq", None, None, None, None, None) == -1


@pytest.mark.gpu
def test_demod_equals_reference_goldens(gpu_lib, gdm):
    from sdrangel_b200 import Demod
    _check_demod(Demod, gdm, exact_kinds=(1, 2, 3))


@pytest.mark.gpu
def test_demod_pooled_bank_layout_vs_oracle(gpu_lib, port):
    """NFM back-end of a bank: DownChannelizer tree + NCO + Interpolator::decimate (the bank), pooled with gather_dev, then
    phaseDiscriminatorDelta on every channel in one launch; per-channel state across two feeds; vs the oracle chain."""
    torch = pytest.importorskip("torch")
    from sdrangel_b200 import DownChannelizerBank, Demod, capi
    fs, n = 10_000_000, 400_000
    x = np.random.RandomState(12).randint(-20000, 20000, size=(n, 2)).astype(np.int16)
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    offs = [1_234_567, -3_000_000, 156_250, -40_000, 4_000_000]
    bank = DownChannelizerBank(fs)
    refs = []
    for fc in offs:
        cid, rate, ofs, _ = bank.add_channel(48000, fc)
        bank.set_frontend(cid, -ofs, cutoff, 48000)
        oc = port.PortDownChannelizer(); oc.configure(fs, 48000, fc)
        refs.append((oc, port.PortFrontEnd(-ofs, rate, 48000, cutoff), port.PortDemod(1, 0.25)))
    nc = len(offs)
    dm = Demod(Demod.FM_DELTA, 0.25, n_channels=nc)
    stride = 4096
    pool = torch.zeros((nc, stride, 2), dtype=torch.float32, device="cuda")
    counts = torch.zeros(nc, dtype=torch.int64, device="cuda")
    out = torch.zeros((3, nc, stride), dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()              # an explicit stream: a NULL stream argument would mean "each handle's own stream"
    st = stream.cuda_stream
    for part in (x[:250_001], x[250_001:]):
        dx = torch.from_numpy(part).cuda()
        torch.cuda.synchronize()
        bank.feed_dev(dx.data_ptr(), part.shape[0], stream=st)
        bank.gather_dev(capi.STAGE_FRONTEND, pool.data_ptr(), stride, counts.data_ptr(), stream=st)
        dm.run_pool_dev(pool.data_ptr(), stride, counts.data_ptr(), out[0].data_ptr(), stride, out[1].data_ptr(), out[2].data_ptr(), stream=st)
        torch.cuda.synchronize()
        cn, o = counts.cpu().numpy(), out.cpu().numpy()
        for c, (oc, fe, od) in enumerate(refs):
            z = fe.feed(oc.feed(part))
            zc = np.ascontiguousarray(z).view(np.complex64).ravel() if z.dtype != np.complex64 else z.ravel()
            assert cn[c] == zc.size and zc.size > 100
            # the oracle discriminator on the GPU front-end's own output isolates the demod kernel: bit-exact
            gz = pool[c, :cn[c]].cpu().numpy().view(np.complex64).ravel()
            w = od.run(gz)
            for j in range(3):
                assert np.array_equal(o[j, c, :cn[c]], w[j]), (c, j)
            # end to end against the oracle chain: the front-end's 1e-5 budget, seen through atan2 (small signals excluded)
            we = port.PortDemod(1, 0.25).run(zc)[1]
            assert np.max(np.abs(o[1, c, :cn[c]] - we)) <= 1e-4 * max(1.0, float(np.max(we)))


def _check_fftfilt(make, gdm, tol):
    """Every golden fftfilt case through `make(kind, f1, f2, len)`: frequency response and outputs within tol of the block maximum
    of both reference builds (their own difference is ~2e-7)."""
    arrays, meta = gdm
    m = meta["fftfilt"]
    x = demod_input(m["seed"], m["n"])
    for ci, (kind, f1, f2, flen) in enumerate(m["cases"]):
        for tag in ("fast", "strict"):
            want = arrays["fftfilt/%s/%d/filter" % (tag, ci)]
            assert np.max(np.abs(make(kind, f1, f2, flen).filter() - want)) <= tol * np.max(np.abs(want)), (ci, tag)
        for oi, (op, usb, dc) in enumerate(m["ops"]):
            f = make(kind, f1, f2, flen)
            got = np.concatenate([f.run(op, x[:m["split"]], usb, dc), f.run(op, x[m["split"]:], usb, dc)])
            for tag in ("fast", "strict"):
                want = arrays["fftfilt/%s/%d/%d" % (tag, ci, oi)]
                assert got.shape == want.shape and want.size >= flen, (ci, oi)
                assert np.max(np.abs(got - want)) <= tol * np.max(np.abs(want)), (ci, oi, tag)


def test_port_fftfilt_equals_reference_goldens(port, gdm):
    _check_fftfilt(port.PortFftFilt, gdm, 2e-6)


@pytest.mark.gpu
def test_fftfilt_equals_reference_goldens(gpu_lib, gdm):
    from sdrangel_b200 import FftFilt
    _check_fftfilt(FftFilt, gdm, 1e-5)


@pytest.mark.gpu
def test_fftfilt_long_run_many_ctas_and_filter_change(gpu_lib, port):
    """Ranges of blocks on many CTAs (each range recomputes the block before it), ragged call splits (inptr carried),
    create_filter on a live object, the device form; against the oracle."""
    torch = pytest.importorskip("torch")
    from sdrangel_b200 import FftFilt
    rs = np.random.RandomState(21)
    n = 300_000
    x = ((rs.randn(n) + 1j * rs.randn(n)) * 3000).astype(np.complex64)
    g, o = FftFilt(0, 300 / 48000.0, 3000 / 48000.0, 1024), port.PortFftFilt(0, 300 / 48000.0, 3000 / 48000.0, 1024)
    pos = 0
    for k, cnt in enumerate((1, 510, 1, 100_003, 0, 511, 199_000 - 26)):
        if k == 4:
            g.set_filter(0, 0.01, 0.1); o.set_filter(0, 0.01, 0.1)
        a, b = g.run(1, x[pos:pos + cnt], usb=(k % 2 == 0), get_dc=False), o.run(1, x[pos:pos + cnt], usb=(k % 2 == 0), get_dc=False)
        assert a.shape == b.shape, (k, cnt)
        if b.size:
            assert np.max(np.abs(a - b)) <= 1e-5 * np.max(np.abs(b)), (k, cnt)
        pos += cnt
    g2, o2 = FftFilt(1, 0.0, 0.125, 2048), port.PortFftFilt(1, 0.0, 0.125, 2048)
    dx = torch.from_numpy(x.view(np.float32).reshape(-1, 2)).cuda()
    dy = torch.zeros((n, 2), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    m = g2.run_dev(2, dx.data_ptr(), n, dy.data_ptr(), n, get_dc=True, stream=st.cuda_stream)
    torch.cuda.synchronize()
    want = o2.run(2, x, get_dc=True)
    assert m == want.size
    got = dy[:m].cpu().numpy().view(np.complex64).ravel()
    assert np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want))
