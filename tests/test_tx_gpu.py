"""GPU parity of the Tx mirror through the C ABI: K8 b200dsp_interps_* == Interpolators<T,16,OutputBits>
(interpolators.h:104-617) and K9 b200dsp_upchan_* == UpChannelizer::pull (upchannelizer.cpp:51-104), bit for bit against the
golden vectors of the unmodified reference and against the C oracle on sizes no fixture holds."""
import ctypes as C

import numpy as np
import pytest

from test_tx_oracle import gtx, run_interps_cases, run_upchan_cases, sha, tx_inputs  # noqa: F401

pytestmark = pytest.mark.gpu


def test_interpolators_equal_reference_goldens(gpu_lib, gtx):  # noqa: F811
    from sdrangel_b200 import Interpolators
    run_interps_cases(Interpolators, gtx)


def test_upchannelizer_equals_reference_goldens(gpu_lib, gtx):  # noqa: F811
    from sdrangel_b200 import UpChannelizer

    class U(UpChannelizer):
        def configure(self, *a):
            r, o, p = super().configure(*a)
            return r, o, p
    run_upchan_cases(U, gtx)


def test_big_cases_hash_equal(gpu_lib, gtx):  # noqa: F811
    """Many CTAs: ranges that start inside the stream warm up from the input instead of the carried state."""
    from sdrangel_b200 import Interpolators, UpChannelizer
    _, meta = gtx
    seed = meta["seed"]
    for bits, log2, n in meta["big"]["interps"]:
        buf, used = Interpolators(bits).run(log2, tx_inputs(seed + 2, n), None, fill=meta["fill"])
        assert used == n and sha(buf) == meta["big_sha"]["interps/%d/%d/%d" % (bits, log2, n)], (bits, log2)
    for pi, n_out in meta["big"]["upchan"]:
        u = UpChannelizer()
        u.configure(*meta["up_plans"][pi])
        out, used = u.pull(tx_inputs(seed + 30 + pi, n_out), n_out)
        want = meta["big_sha"]["upchan/%d/%d" % (pi, n_out)]
        assert used == want["used"] and sha(out) == want["sha"], pi


def test_interpolators_vs_oracle_random_splits(gpu_lib, port):
    """Arbitrary call splits and factor changes: the six stage rings persist exactly like the reference objects'."""
    from sdrangel_b200 import Interpolators
    rs = np.random.RandomState(2024)
    for bits in (16, 12, 8):
        g, o = Interpolators(bits), port.PortInterpolators(bits)
        for it in range(24):
            log2 = int(rs.randint(0, 7))
            n = int(rs.choice([1, 2, 3, 63, 64, 65, 1000, 5000, 40_000]))
            x = rs.randint(-32768, 32768, size=(n, 2)).astype(np.int16)
            if it % 5 == 0:
                x[:] = rs.choice([-32768, 32767])
            length = n * (2 << log2) + int(rs.randint(0, 2 << log2))
            a, na = g.run(log2, x, length, fill=33)
            b, nb = o.run(log2, x, length, fill=33)
            assert na == nb and np.array_equal(a, b), (bits, it, log2, n)
        g.reset()
        o2 = port.PortInterpolators(bits)
        x = rs.randint(-32768, 32768, size=(100, 2)).astype(np.int16)
        assert np.array_equal(g.run(3, x)[0], o2.run(3, x)[0])


def test_upchannelizer_vs_oracle_random_pulls_and_reconfiguration(gpu_lib, port):
    from sdrangel_b200 import UpChannelizer
    rs = np.random.RandomState(77)
    plans = [(10_000_000, 48_000, 1_234_567), (2_400_000, 300_000, -700_000), (122_880_000, 48_000, -61_380_000), (48_000, 48_000, 0)]
    g, o = UpChannelizer(), port.PortUpChannelizer()
    for plan in plans:
        cg, co = g.configure(*plan), o.configure(*plan)
        assert cg == co, plan
        for it in range(14):
            n_out = int(rs.choice([1, 2, 3, 4, 5, 511, 1024, 1025, 2049, 50_000]))
            need = g.source_count(n_out)
            src = rs.randint(-32768, 32768, size=(need + 3, 2)).astype(np.int16)
            if it == 3:
                src[:] = -32768
            a, ua = g.pull(src, n_out)
            b, ub = o.pull(src, n_out)
            assert ua == ub == need and np.array_equal(a, b), (plan, it, n_out)


def test_device_forms_and_bad_arguments(gpu_lib, port):
    torch = pytest.importorskip("torch")
    from sdrangel_b200 import Interpolators, UpChannelizer, capi
    L = capi.lib()
    rs = np.random.RandomState(5)
    n = 30_000
    x = rs.randint(-32768, 32768, size=(n, 2)).astype(np.int16)
    dx = torch.from_numpy(x).cuda()
    stream = torch.cuda.Stream()              # an explicit stream: a NULL stream argument would mean "the handle's own stream"
    sp = stream.cuda_stream
    for bits, log2 in ((16, 3), (8, 6), (12, 5)):
        g = Interpolators(bits)
        dt = torch.int8 if bits == 8 else torch.int16
        dbuf = torch.full((n * (2 << log2),), 21, dtype=dt, device="cuda")
        torch.cuda.synchronize()              # the fill above ran on torch's stream
        used = g.run_dev(log2, dx.data_ptr(), dbuf.data_ptr(), dbuf.numel(), stream=sp)
        torch.cuda.synchronize()
        want, nw = port.PortInterpolators(bits).run(log2, x, None, fill=21)
        assert used == nw and np.array_equal(dbuf.cpu().numpy(), want), (bits, log2)
    u, o = UpChannelizer(), port.PortUpChannelizer()
    u.configure(4_000_000, 100_000, 1_200_000)
    o.configure(4_000_000, 100_000, 1_200_000)
    n_out = 100_001
    need = u.source_count(n_out)
    dout = torch.zeros((n_out, 2), dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    u.pull_dev(dx.data_ptr(), need, dout.data_ptr(), n_out, stream=sp)
    torch.cuda.synchronize()
    want, used = o.pull(x, n_out)
    assert used == need and np.array_equal(dout.cpu().numpy(), want)
    # errors are reported, nothing is computed
    h = C.c_void_p()
    assert L.b200dsp_interps_create(C.byref(h), capi.FMT_I8, 16) == -1 and L.b200dsp_interps_create(C.byref(h), capi.FMT_F32, 16) == -1
    g = Interpolators(16)
    with pytest.raises(capi.B200DspError):
        g.run_dev(7, dx.data_ptr(), dout.data_ptr(), 1024)
    with pytest.raises(capi.B200DspError):
        u.pull_dev(dx.data_ptr(), 1, dout.data_ptr(), 100_000)          # too few modulator samples for that many pulls
    with pytest.raises(capi.B200DspError):
        u.set_path([0, 3])
    buf, used = g.run(4, x[:0], 0)
    assert used == 0 and buf.size == 0
