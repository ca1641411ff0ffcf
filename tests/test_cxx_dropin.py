"""The reference-facing C++ boundary: include/sdrangel_b200/dsp/*.h keep the reference's class names and method
signatures over the C ABI.  CPU: the headers compile and link; GPU: sdrbench's decimateII code path run through them
reproduces the reference's output hash."""
import os
import subprocess

import numpy as np

import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "build", "cxx_dropin")


def build_exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), "-o", EXE,
                           os.path.join(ROOT, "tests", "cxx_dropin.cpp"), "-L" + os.path.join(ROOT, "sdrangel_b200", "lib"), "-lb200dsp",
                           "-Wl,-rpath," + os.path.join(ROOT, "sdrangel_b200", "lib")])


def test_wrapper_headers_compile_and_link():
    from sdrangel_b200 import capi
    capi.lib()
    build_exe()
    if capi.device_count() == 0:      # without a device the program must fail loudly, not fall back
        r = subprocess.run([EXE], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_sdrbench_decimateii_through_cxx_wrappers(gpu_lib, port, golden_meta, golden_x8):
    build_exe()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = dict(l.split(" ", 1) for l in r.stdout.strip().splitlines())
    g = golden_meta["decim_ii_sdrbench"]["12/4/cen"]
    assert "n_out=%d" % g["n_out"] in lines["decimate16_cen"]
    assert "in=" + golden_meta["sdrbench_s16"]["fnv"] in lines["decimate16_cen"]      # same libstdc++ generator as the reference run
    assert "out=" + g["fnv"] in lines["decimate16_cen"]
    rows = golden_x8[1]["long"]["rows"]
    assert lines["decimatorsu16_cen"] == "n_out=%d out=%s" % (rows["u8/4/cen"]["n_out"], rows["u8/4/cen"]["fnv"])
    assert lines["decimators8_64_inf"] == "n_out=%d out=%s" % (rows["i8/6/inf"]["n_out"], rows["i8/6/inf"]["fnv"])
    from conftest import fnv1a64_u16
    q = port.PortIQCorrections()
    xs = port.sdrbench_s16(1 << 20)[: 2 << 16].reshape(-1, 2)      # the C++ program corrects the first 2^16 samples of its 2^20 buffer
    want = np.concatenate([q.run(xs[:40000]), q.run(xs[40000:])])
    assert lines["iqcorrections"] == "out=" + fnv1a64_u16(want)
    assert lines["iqcorrections_imbalance"] == "silent"          # the imbalance branch exists since round 2 (no exception)
    assert lines["downchannelizer"].startswith("rate=156250 ofs=-15433 n_out=937")
    assert lines["spectrumvis"] == "frames=2"
    assert lines["interpolator"] == "n_out=6145"          # 20 000 inputs at 156 250 -> 48 000
    # the reference's per-sample method signatures equal the block forms; NCO increment of SURVEY.md Appendix D
    assert lines["interpolator_per_sample_decimate"].startswith("same n=36")
    assert lines["interpolator_per_sample_interpolate"].startswith("same")
    assert lines["nco"] == "inc=404 same"
    # SampleSinkFifo (samplesinkfifo.cpp:113-231): 700 in, 500 out, 700 in (wraps), then only 100 fit; read of 900 = two spans
    assert lines["samplesinkfifo"] == "w=700,700,100 r=500 begin=900 spans=500+400 first=500 second=300 fill=1000"
    assert lines["devicesamplesinkfifo"] == "w=700,700,100 r=500 begin=900 spans=500+400 fill=1000 r0=499"
    assert lines["frequencyshift"] == "-1250000 625000 434375000"
    assert lines["device_fifo_route"].startswith("same n=")
    assert lines["split_iq_overload"] == "same n=256"
    # Tx mirror: per-sample pull == block pull == the oracle; Interpolators<qint16,16,12>::interpolate8_cen of that stream
    k = np.arange(700, dtype=np.int64)
    src = np.stack([((k * 37 - 20000 + 32768) % 65536 - 32768), ((15000 - k * 91 + 32768) % 65536 - 32768)], axis=1).astype(np.int16)
    u = port.PortUpChannelizer()
    rate, ofs, _ = u.configure(4000000, 100000, 1200000)
    out, used = u.pull(src, 10000)
    assert lines["upchannelizer"] == "same rate=%d ofs=%d pulled=%d out=%s" % (rate, ofs, used, fnv1a64_u16(out))
    dev, n = port.PortInterpolators(12).run(3, out, 10000 * 16 + 5, fill=77)
    assert lines["interpolators8_cen"] == "consumed=%d tail=77 out=%s" % (n, fnv1a64_u16(dev))
    assert lines["phasediscri"] == "same"
    assert lines["filerecord"] == "rate=2400000 fc=434000000 size=16 count=1000 bytes=4024"
    assert lines["fftfilt"] == "same n=1536"
