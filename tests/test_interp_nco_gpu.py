"""GPU parity of the stand-alone Interpolator (decimate / interpolate / resample block forms) and NCO through the C ABI
against vectors from the reference built in place (tests/golden/golden_interp.npz, oracle/gen_golden_interp.py) and the
C restatement: identical output counts per call and carried distance (the float32 recurrence is replayed exactly), values
within 1e-5 relative RMS (north_star tolerance); NCO samples bit-identical (table values)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, rel_rms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden_interp():
    z = np.load(os.path.join(GOLDEN_DIR, "golden_interp.npz"))
    with open(os.path.join(GOLDEN_DIR, "golden_interp.json")) as f:
        return {k: z[k] for k in z.files}, json.load(f)


def test_interpolate_and_resample_golden(gpu_lib, golden_interp):
    from sdrangel_b200 import Interpolator
    g, meta = golden_interp
    rs = np.random.RandomState(meta["seed"])
    n = meta["n"]
    x = (rs.randint(-20000, 20000, size=n) + 1j * rs.randint(-20000, 20000, size=n)).astype(np.complex64)
    cuts = meta["cuts"]
    for rin, rout in meta["cases"]:
        cutoff = float(np.float32(min(rin, rout) / 2.2))
        dist = float(np.float32(np.float32(rin) / np.float32(rout)))
        for name in ("interpolate", "resample"):
            key = "%s/strict/%d_%d" % (name, rin, rout)
            if key + "/out" not in g:
                continue
            it = Interpolator(16, max(rin, rout), cutoff)
            remain, outs = 0.0, []
            for a, b in zip(cuts[:-1], cuts[1:]):
                o, remain = getattr(it, name)(remain, dist, x[a:b])
                outs.append(o.copy())
            assert [o.shape[0] for o in outs] == g[key + "/counts"].tolist(), key
            assert np.float32(remain) == g[key + "/remain"][0], key
            out = np.concatenate(outs)
            assert rel_rms(out.view(np.float32), g[key + "/out"].view(np.float32)) <= 1e-5, key
            assert rel_rms(out.view(np.float32), g["%s/fast/%d_%d/out" % (name, rin, rout)].view(np.float32)) <= 1e-5, key
            it.close()


def test_interpolate_with_no_input_and_large_ratio(gpu_lib, port):
    """Edge cases: a call with zero inputs still emits the outputs the Tx loop would (distance_remain < 1); an
    interpolation by ~21 (2250 -> 48000); everything against the C restatement call by call."""
    from sdrangel_b200 import Interpolator
    rs = np.random.RandomState(11)
    x = (rs.randint(-3000, 3000, size=900) + 1j * rs.randint(-3000, 3000, size=900)).astype(np.complex64)
    rin, rout = 2250, 48000
    cutoff = float(np.float32(rin / 2.2))
    dist = float(np.float32(np.float32(rin) / np.float32(rout)))
    for name, mode in (("interpolate", 1), ("resample", 2)):
        it = Interpolator(16, rout, cutoff)
        fe = port.PortFrontEnd(0, rout, rout, cutoff)
        port.load().orc_frontend_destroy(fe.h)
        fe.h = port.load().orc_frontend_create(0.0, float(rout), 16, float(rout), cutoff, 4.5, dist)
        remain = 0.0
        for a, b in ((0, 0), (0, 1), (1, 400), (400, 400), (400, 900)):
            got, remain = getattr(it, name)(remain, dist, x[a:b])
            want = fe.run_c64(mode, x[a:b])
            assert got.shape == want.shape, (name, a, b, got.shape, want.shape)
            assert np.float32(remain) == np.float32(fe.remain()), (name, a, b)
            if want.size:
                assert rel_rms(got.view(np.float32), want.view(np.float32)) <= 1e-5, (name, a, b)
        it.close()


def test_nco_block_golden(gpu_lib, golden_interp):
    from sdrangel_b200 import NCO
    g, meta = golden_interp
    for freq, rate in meta["nco_cases"]:
        nco = NCO()
        nco.setFreq(freq, rate)
        got = np.concatenate([nco.nextIQ(k) for k in (1, 7, 2000, 2992)])
        assert np.array_equal(got, g["nco/%g_%g" % (freq, rate)]), (freq, rate)
        p, inc = nco.state()
        assert p == (5000 * inc) % 4096
        nco.close()


@pytest.mark.parametrize("ratio", [156250.0 / 48000.0, 78125.0 / 48000.0, 1.0000001, 1.9999, 3.0000002, 7.3, 15.99, 31.9, 63.5])
def test_parallel_schedule_equals_serial_replay(gpu_lib, ratio):
    """K4's exact schedule for non-lattice ratios runs as a warp-parallel scan (closed-form residues + a 4-state automaton for
    the float32 tie-rounding corrections); B200DSP_SERIAL_SCHEDULE=1 forces the one-lane serial replay of the same recurrence.
    Both must give the same outputs bit for bit (same schedule, same arithmetic), call after call, with the carried distance
    identical -- 2^21 inputs per ratio, ragged calls, a caller-set off-lattice starting distance."""
    from sdrangel_b200 import Interpolator
    rs = np.random.RandomState(int(ratio * 1000))
    n = 1 << 21
    x = (rs.randint(-20000, 20000, size=n) + 1j * rs.randint(-20000, 20000, size=n)).astype(np.complex64)
    a, b = Interpolator(16, 48000.0 * ratio, 12500 / 2.2), Interpolator(16, 48000.0 * ratio, 12500 / 2.2)
    step = float(np.float32(ratio))
    ra = rb = 0.1
    cuts = [0, 5, 1000, 1001, 700_000, 700_129, n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        os.environ.pop("B200DSP_SERIAL_SCHEDULE", None)
        ya, ra = a.decimate(ra, step, x[lo:hi])
        os.environ["B200DSP_SERIAL_SCHEDULE"] = "1"
        try:
            yb, rb = b.decimate(rb, step, x[lo:hi])
        finally:
            os.environ.pop("B200DSP_SERIAL_SCHEDULE", None)
        assert ya.shape == yb.shape, (ratio, lo, hi, ya.shape, yb.shape)
        assert np.float32(ra) == np.float32(rb), (ratio, lo, hi)
        assert np.array_equal(ya, yb), (ratio, lo, hi)
    a.close()
    b.close()
