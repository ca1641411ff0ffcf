"""CPU tests of the Tx-mirror oracle: the C restatement (oracle/port: orc_interps_*, orc_upchan_*) against the
golden vectors the unmodified reference produced (oracle/gen_golden_tx.py), the host arithmetic of the C ABI without a device,
and -- where the in-place reference build is present (this container) -- the reference itself on fresh random input."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR


@pytest.fixture(scope="module")
def gtx():
    z = np.load(os.path.join(GOLDEN_DIR, "golden_tx.npz"))
    with open(os.path.join(GOLDEN_DIR, "golden_tx.json")) as f:
        return {k: z[k] for k in z.files}, json.load(f)


def tx_inputs(seed, n):
    return np.random.RandomState(seed).randint(-32768, 32768, size=(n, 2)).astype(np.int16)


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_interps_cases(make, gtx):
    """Shared by the CPU (port) and GPU (CUDA) tests: every golden Interpolators case through `make(bits)`."""
    arrays, meta = gtx
    cuts, seed, fill = meta["interp_cuts"], meta["seed"], meta["fill"]
    x = tx_inputs(seed, cuts[-1])
    for bits in (16, 12, 8):
        for log2 in range(7):
            d = make(bits)
            for k, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
                length = (b - a) * (2 << log2) + (k * 3) % (2 << log2)
                buf, n = d.run(log2, x[a:b], length, fill=fill)
                want = arrays["interps/%d/%d/%d" % (bits, log2, k)]
                assert n == b - a, (bits, log2, k)
                assert buf.dtype == want.dtype and np.array_equal(buf, want), (bits, log2, k)
        d = make(bits)
        xs = tx_inputs(seed + 1, 40 * len(meta["switch"]))
        for k, log2 in enumerate(meta["switch"]):
            buf, n = d.run(log2, xs[40 * k:40 * (k + 1)], None, fill=fill)
            assert np.array_equal(buf, arrays["interps_switch/%d/%d" % (bits, k)]), (bits, k, log2)


def run_upchan_cases(make, gtx):
    arrays, meta = gtx
    seed = meta["seed"]
    for pi, (orate, req, fc) in enumerate(meta["up_plans"]):
        u = make()
        rate, ofs, modes = u.configure(orate, req, fc)
        cfg = meta["up_cfg"][pi]
        assert (rate, ofs, list(modes)) == (cfg["in_rate"], cfg["ofs"], cfg["modes"]), pi
        src = tx_inputs(seed + 10 + pi, sum(meta["up_pulls"]) + 16)
        pos = 0
        for k, n_out in enumerate(meta["up_pulls"]):
            s = src[pos:].copy()
            if n_out == 1000:
                s[:] = -32768
            out, used = u.pull(s, n_out)
            assert used == int(arrays["upchan/%d/used" % pi][k]), (pi, k)
            assert np.array_equal(out, arrays["upchan/%d/%d" % (pi, k)]), (pi, k, n_out)
            pos += used


def test_port_interpolators_equal_reference_goldens(port, gtx):
    run_interps_cases(port.PortInterpolators, gtx)


def test_port_upchannelizer_equals_reference_goldens(port, gtx):
    run_upchan_cases(port.PortUpChannelizer, gtx)


def test_port_big_cases_hash_equal(port, gtx):
    _, meta = gtx
    seed = meta["seed"]
    for bits, log2, n in meta["big"]["interps"]:
        buf, used = port.PortInterpolators(bits).run(log2, tx_inputs(seed + 2, n), None, fill=meta["fill"])
        assert used == n and sha(buf) == meta["big_sha"]["interps/%d/%d/%d" % (bits, log2, n)]
    for pi, n_out in meta["big"]["upchan"]:
        u = port.PortUpChannelizer()
        u.configure(*meta["up_plans"][pi])
        out, used = u.pull(tx_inputs(seed + 30 + pi, n_out), n_out)
        want = meta["big_sha"]["upchan/%d/%d" % (pi, n_out)]
        assert used == want["used"] and sha(out) == want["sha"]


def test_port_coefficients_are_the_reference_tables(port, gtx):
    _, meta = gtx
    for o in (16, 32, 64, 96):
        assert list(port.hb_interp_coeffs(o)) == meta["coeffs"][str(o)]["h"]
    assert [meta["coeffs"][str(o)]["shift"] for o in (16, 32, 64, 96)] == [12, 12, 12, 16]


def test_interps_in_count_is_the_reference_loop_bound(port):
    """b200dsp_interps_in_count: pure host arithmetic, == iterations of `for (pos = 0; pos < len - (2N-1); pos += 2N)`."""
    from sdrangel_b200 import capi
    L = capi.lib()
    for log2 in range(7):
        for length in (0, 1, 2, 3, 127, 128, 129, 255, 256, 1000, 4097, 65536 + 5):
            x = np.zeros((length // 2 + 1, 2), dtype=np.int16)
            _, n = port.PortInterpolators(16).run(log2, x, length)
            assert L.b200dsp_interps_in_count(log2, length) == n, (log2, length)
    assert L.b200dsp_interps_in_count(7, 100) == -1 and L.b200dsp_interps_in_count(2, -1) == -1


def test_tx_calls_report_no_device_instead_of_falling_back():
    from sdrangel_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert capi.lib().b200dsp_interps_create(C.byref(h), capi.FMT_I16, 16) == capi.ENODEV and not h.value
    assert capi.lib().b200dsp_upchan_create(C.byref(h)) == capi.ENODEV and not h.value


def test_port_equals_reference_on_fresh_input(port):
    """Only where oracle/_ref was built (the authoring container): port vs the reference on input no fixture holds."""
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built here")
    try:
        refbind.load().ref_interps_create
    except AttributeError:
        pytest.skip("oracle/_ref predates the Tx driver")
    rs = np.random.RandomState(99)
    for bits in (16, 12, 8):
        r, p = refbind.RefInterpolators(bits), port.PortInterpolators(bits)
        for log2 in (2, 6, 0, 5, 1, 3, 4):
            x = rs.randint(-32768, 32768, size=(257, 2)).astype(np.int16)
            a, b = r.run(log2, x, None, fill=9), p.run(log2, x, None, fill=9)
            assert a[1] == b[1] and np.array_equal(a[0], b[0]), (bits, log2)
    for plan in ((10_000_000, 48_000, 1_234_567), (2_400_000, 300_000, -700_000)):
        r, p = refbind.RefUpChannelizer(), port.PortUpChannelizer()
        assert r.configure(*plan) == p.configure(*plan)
        for n_out in (3, 1001, 2, 20_000):
            src = rs.randint(-32768, 32768, size=(n_out + 4, 2)).astype(np.int16)
            a, b = r.pull(src, n_out), p.pull(src, n_out)
            assert a[1] == b[1] and np.array_equal(a[0], b[0]), (plan, n_out)
