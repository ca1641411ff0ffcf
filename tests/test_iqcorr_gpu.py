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
    with pytest.raises(RuntimeError):
        q.iqCorrections(x[:10].copy(), imbalanceCorrection=True)      # not implemented: loud, no silent DC-only fallback


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
