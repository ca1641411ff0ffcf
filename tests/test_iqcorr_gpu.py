"""GPU parity of the engine-side DC correction through the C ABI: bit-exact against the golden vectors
generated from the reference's MovingAverageUtil loop and against the oracle port on large ragged streams."""
import numpy as np
import pytest

from conftest import fnv1a64_u16

pytestmark = pytest.mark.gpu


def test_iqcorrections_dc_golden_bit_exact(gpu_lib, golden_x8):
    from sdrangel_b200 import IQCorrections
    arrays, meta = golden_x8
    x, cuts = arrays["iqcorr/in"], meta["iqcorr"]["cuts"]
    q = IQCorrections()
    out = np.concatenate([q.iqCorrections(x[a:b].copy()) for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.array_equal(out, arrays["iqcorr/out"]), int(np.argmax(np.any(out != arrays["iqcorr/out"], axis=1)))
    assert fnv1a64_u16(out) == meta["iqcorr"]["fnv"]
    q.reset()                                               # a reset object behaves like a new one
    again = np.concatenate([q.iqCorrections(x[a:b].copy()) for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.array_equal(again, out)


def test_iqcorrections_dc_large_ragged_vs_oracle_and_device_forms(gpu_lib, port):
    import torch
    from sdrangel_b200 import IQCorrections
    rs = np.random.RandomState(11)
    n = 3_000_001
    x = (rs.randint(-20000, 20000, size=(n, 2)) + np.array([5000, -7000])).clip(-32768, 32767).astype(np.int16)
    x[1_000_000:1_002_000] = -32768
    cuts = [0, 1, 1023, 1025, 500_000, 500_001, 2_000_000, n]
    q, o = IQCorrections(), port.PortIQCorrections()
    for a, b in zip(cuts[:-1], cuts[1:]):
        got, want = q.iqCorrections(x[a:b].copy()), o.run(x[a:b])
        assert np.array_equal(got, want), (a, b, int(np.argmax(np.any(got != want, axis=1))))
    # device forms: out of place and in place give the same samples as a fresh oracle
    dev = torch.device("cuda:0")
    dx = torch.from_numpy(x).to(dev)
    dy = torch.empty_like(dx)
    q2, q3, o2 = IQCorrections(), IQCorrections(), port.PortIQCorrections()
    q2.run_dev(dx.data_ptr(), dy.data_ptr(), n)
    torch.cuda.synchronize()                                # q3 overwrites what q2 reads
    q3.run_dev(dx.data_ptr(), dx.data_ptr(), n)
    torch.cuda.synchronize()
    want = o2.run(x)
    assert np.array_equal(dy.cpu().numpy(), want)
    assert np.array_equal(dx.cpu().numpy(), want)


def test_iqcorrections_imbalance_golden(gpu_lib, golden_x8):
    """The I/Q imbalance branch (dspdevicesourceengine.cpp:219-252, floating-point flavour) against the vectors of both reference
    builds, fed in the fixture's ragged calls: equal to the build without -ffast-math (the kernel rounds every operation
    separately, in the reference's order), within 1 LSB of the -ffast-math build."""
    from sdrangel_b200 import IQCorrections
    arrays, meta = golden_x8
    x, cuts = arrays["iqcorr_imb/in"], meta["iqcorr_imb"]["cuts"]
    q = IQCorrections()
    out = np.concatenate([q.iqCorrections(x[a:b].copy(), True) for a, b in zip(cuts[:-1], cuts[1:])])
    strict = arrays["iqcorr_imb/out_strict"]
    diff = np.abs(out.astype(np.int32) - strict.astype(np.int32))
    assert diff.max() <= 1 and np.count_nonzero(diff) <= 1e-3 * diff.size, (int(diff.max()), int(np.count_nonzero(diff)))
    assert np.max(np.abs(out.astype(np.int32) - arrays["iqcorr_imb/out"].astype(np.int32))) <= 1


def test_iqcorrections_imbalance_long_stream_vs_oracle(gpu_lib, port):
    """2^20 samples of an imbalanced tone + noise, ragged calls (segments of 2048 samples replayed by one thread each after a
    2048-sample warm-up): within 1 LSB of the sequential oracle, differing samples below 0.1 % (the reference's running
    totals carry rounding drift from before a segment's warm-up), and the DC branch keeps working on the same
    object afterwards (both branches share the DC averages)."""
    from sdrangel_b200 import IQCorrections
    rs = np.random.RandomState(5)
    n = 1 << 20
    t = np.arange(n)
    i = 9000 * np.cos(2 * np.pi * 0.01 * t) + 300 + rs.normal(0, 500, n)
    qd = 7000 * np.sin(2 * np.pi * 0.01 * t + 0.2) - 200 + rs.normal(0, 500, n)
    x = np.stack([i, qd], axis=1).clip(-32768, 32767).astype(np.int16)
    x[400_000:400_600] = 0
    g, o = IQCorrections(), port.PortIQCorrections()
    cuts = [0, 1, 5000, 5001, 300_000, 700_001, n]
    worst, bad, tot = 0, 0, 0
    for a, b in zip(cuts[:-1], cuts[1:]):
        got, want = g.iqCorrections(x[a:b].copy(), True), o.run(x[a:b], True)
        d = np.abs(got.astype(np.int32) - want.astype(np.int32))
        worst, bad, tot = max(worst, int(d.max())), bad + int(np.count_nonzero(d)), tot + d.size
    assert worst <= 1 and bad <= 1e-3 * tot, (worst, bad, tot)
    y = rs.randint(-3000, 3000, size=(5000, 2)).astype(np.int16)
    assert np.array_equal(g.iqCorrections(y.copy(), False), o.run(y, False))
