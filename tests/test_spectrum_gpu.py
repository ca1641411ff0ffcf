"""GPU parity tests of K5 (SpectrumVis: window + FFT + power + averaging + log) through the C ABI, against the golden
frames generated from the reference (KissFFT).  Frame counts identical.  Tolerances: linear power rel. RMS <= 1e-5
(north_star bar; dB frames are converted back to linear power for it); dB output abs. diff <= 1e-3 dB on
every bin within 60 dB of its frame's maximum, and <= 1e-2 dB on the weaker bins, where two float32 FFTs with different
operation orders (KissFFT's radix-4 recursion vs the kernel's radix-16 passes) differ by their own rounding noise."""
import numpy as np
import pytest

from conftest import rel_rms

pytestmark = pytest.mark.gpu


def spectrum_input(meta):
    m = meta["spectrum"]
    rs = np.random.RandomState(m["seed"])
    n = m["n"]
    sx = rs.randint(-2048, 2048, size=(n, 2)).astype(np.int16)
    t = np.arange(n)
    tone = m["tone"][0] * np.exp(2j * np.pi * m["tone"][1] * t)
    sx[:, 0] += tone.real.astype(np.int16)
    sx[:, 1] += tone.imag.astype(np.int16)
    return m, sx


def check_frames(got, want, linear):
    assert got.shape == want.shape
    if linear:
        assert rel_rms(got, want) <= 1e-5
    else:
        ok = np.isfinite(want) & np.isfinite(got)
        assert ok.mean() > 0.99
        strong = ok & (want >= want.max(axis=1, keepdims=True) - 60.0)
        assert np.max(np.abs(got[strong] - want[strong])) <= 1e-3
        assert np.max(np.abs(got[ok] - want[ok])) <= 1e-2
        assert rel_rms(10.0 ** (got[ok] / 10.0), 10.0 ** (want[ok] / 10.0)) <= 1e-5


def test_spectrum_golden_all_modes(gpu_lib, golden, golden_meta):
    from sdrangel_b200 import SpectrumVis
    m, sx = spectrum_input(golden_meta)
    for fft, mode, nb, linear, posonly in m["cases(fft,avg_mode,avg_nb,linear,positive_only)"]:
        s = SpectrumVis()
        s.configure(fft, 0, nb, mode, 1, linear)
        fr = np.concatenate([s.feed(sx[a:b], posonly) for a, b in zip(m["cuts"][:-1], m["cuts"][1:])])
        for tag in ("strict", "fast"):
            key = f"spectrum/{tag}/{fft}_{mode}_{nb}_{int(linear)}_{int(posonly)}"
            assert fr.shape[0] == int(golden[key + "/nframes"][0]), key
            keep = np.concatenate([fr[:3], fr[-3:]]) if fr.shape[0] > 6 else fr
            check_frames(keep, golden[key], linear)
        s.close()


@pytest.mark.parametrize("fft,mode,nb,linear", [(4096, 2, 10, False), (4096, 1, 10, False), (2048, 0, 0, True), (512, 2, 3, True), (128, 1, 5, False)])
def test_spectrum_vs_oracle_ragged_feeds(gpu_lib, port, fft, mode, nb, linear):
    """Config-4 style stream in ragged feeds (partial-frame and averager carry) vs the oracle port, every frame."""
    from sdrangel_b200 import SpectrumVis
    rs = np.random.RandomState(7 + fft)
    n = fft * 57 + 333
    x = rs.randint(-2048, 2048, size=(n, 2)).astype(np.int16)
    t = np.arange(n)
    x[:, 0] += (1200 * np.cos(2 * np.pi * 0.171 * t)).astype(np.int16)
    x[:, 1] += (1200 * np.sin(2 * np.pi * 0.171 * t)).astype(np.int16)
    s, o = SpectrumVis(), port.PortSpectrumVis()
    s.configure(fft, 0, nb, mode, 1, linear)
    o.configure(fft, 0, nb, mode, 1, linear)
    cuts = [0, 1, fft - 1, fft, 3 * fft + 5, 20 * fft + 17, 20 * fft + 18, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        got, want = s.feed(x[a:b]), o.feed(x[a:b])
        assert got.shape == want.shape, (a, b)
        if got.shape[0]:
            check_frames(got, want, linear)


def test_spectrum_window_and_tone_bin(gpu_lib):
    """Size-independent property: a pure tone at bin k of an N-point rectangular-window FFT puts (A*N/scalef)^2 / N^2 of
    linear power in the DC-centred bin k + N/2 and ~nothing elsewhere."""
    from sdrangel_b200 import SpectrumVis
    n, k, amp = 4096, 300, 8192.0
    t = np.arange(4 * n)
    x = np.stack([np.round(amp * np.cos(2 * np.pi * k * t / n)), np.round(amp * np.sin(2 * np.pi * k * t / n))], axis=1).astype(np.int16)
    s = SpectrumVis()
    s.configure(n, 0, 0, 0, 5, True)
    fr = s.feed(x)
    assert fr.shape == (4, n)
    peak = (amp / 32768.0) ** 2
    assert np.all(np.argmax(fr, axis=1) == k + n // 2)
    assert np.allclose(fr[:, k + n // 2], peak, rtol=1e-3)
    assert np.all(np.delete(fr[0], k + n // 2) < peak * 1e-6)
