"""GPU parity tests of K3 (shared-prefix HB48 tree) and K4 (NCO + polyphase Interpolator front-end) through the
C ABI.  Channelizer outputs: bit-exact (int16).  Front-end: identical output count/schedule, values within
1e-5 relative RMS (north_star tolerance)."""
import os

import numpy as np
import pytest

from conftest import rel_rms

pytestmark = pytest.mark.gpu


def chan_input(meta):
    m = meta["chan_feed"]
    rs = np.random.RandomState(m["seed"])
    cx = rs.randint(-32768, 32768, size=(m["n"], 2)).astype(np.int16)
    cx[m["min_run"][0]:m["min_run"][1]] = -32768
    return m, cx


def test_filter_chain_plans_match_reference(gpu_lib, golden_meta):
    from sdrangel_b200 import DownChannelizerBank
    plans = golden_meta["chan_plans"]
    for name in ("bank64", "bank1024"):
        b = DownChannelizerBank(plans[name]["input_rate"])
        for fc, rate, ofs, path in plans[name]["channels"]:
            cid, r, o, p = b.add_channel(48000, fc)
            assert (r, o, p) == (rate, ofs, path), (name, fc)
        b.close()
    for fs, req, fc, rate, ofs, path in plans["random"]:
        b = DownChannelizerBank(fs)
        assert b.add_channel(req, fc)[1:] == (rate, ofs, path)
        b.close()


def test_bank_feed_golden_bit_exact(gpu_lib, golden, golden_meta):
    """6 channels sharing one tree, awkward feed splits (odd lengths), a run of -32768 (int16 negation wrap)."""
    from sdrangel_b200 import DownChannelizerBank
    m, cx = chan_input(golden_meta)
    b = DownChannelizerBank(m["input_rate"])
    ids = [b.add_channel(m["requested_rate"], fc)[0] for fc in m["offsets"]]
    outs = {c: [] for c in ids}
    for a, e in zip(m["cuts"][:-1], m["cuts"][1:]):
        b.feed(cx[a:e])
        for c in ids:
            outs[c].append(b.fetch(c).copy())
    for c, fc in zip(ids, m["offsets"]):
        got, want = np.concatenate(outs[c]), golden[f"chan_feed/{fc}"]
        assert got.shape == want.shape, (fc, got.shape, want.shape)
        assert np.array_equal(got, want), (fc, int(np.argmax(np.any(got != want, axis=1))))


@pytest.mark.parametrize("chunk", [768, 3 << 18])
def test_bank64_vs_oracle_per_call_counts(gpu_lib, port, golden_meta, chunk):
    """BASELINE config 3 plan (64 NFM channels @ 10 MS/s): every channel, every call, vs one oracle chain per channel."""
    from sdrangel_b200 import DownChannelizerBank
    plan = golden_meta["chan_plans"]["bank64"]
    rs = np.random.RandomState(33)
    n = 300_000
    x = rs.randint(-32768, 32768, size=(n, 2)).astype(np.int16)
    b = DownChannelizerBank(plan["input_rate"])
    b.set_chunk(chunk)
    chans = plan["channels"][::5] + plan["channels"][-1:]
    ids = [b.add_channel(48000, fc)[0] for fc, _, _, _ in chans]
    assert b.node_count() > 0
    oracles = []
    for fc, _, _, _ in chans:
        o = port.PortDownChannelizer()
        o.configure(plan["input_rate"], 48000, fc)
        oracles.append(o)
    cuts = [0, 1, 4, 1001, 65536 + 1001, 65536 + 1002, 200_001, n]
    for a, e in zip(cuts[:-1], cuts[1:]):
        b.feed(x[a:e])
        for cid, o in zip(ids, oracles):
            got, want = b.fetch(cid), o.feed(x[a:e])
            assert got.shape == want.shape, (a, e, cid, got.shape, want.shape)
            assert np.array_equal(got, want), (a, e, cid)


def test_bank_shared_tree_is_smaller_than_independent_chains(gpu_lib, golden_meta):
    from sdrangel_b200 import DownChannelizerBank
    plan = golden_meta["chan_plans"]["bank64"]
    b = DownChannelizerBank(plan["input_rate"])
    total = 0
    for fc, _, _, path in plan["channels"]:
        b.add_channel(48000, fc)
        total += len(path)
    assert b.node_count() < total
    assert b.node_count() == len({p[:k] for _, _, _, p in plan["channels"] for k in range(1, len(p) + 1)})


def test_frontend_schedule_and_values_golden(gpu_lib, golden, golden_meta):
    """NCO + Interpolator::decimate on a stage-less channel (requested rate = input rate => DownChannelizer forwards
    samples unchanged), same inputs/cuts as the golden fixtures."""
    from sdrangel_b200 import DownChannelizerBank, capi
    m = golden_meta["frontend"]
    rs = np.random.RandomState(m["seed"])
    fx = rs.randint(-20000, 20000, size=(m["n"], 2)).astype(np.int16)
    cutoff = np.float32(np.float32(12500) / np.float32(2.2))
    for freq, rate, outr in m["cases"]:
        b = DownChannelizerBank(rate)
        cid, r, ofs, path = b.add_channel(rate, 0)
        assert path == "" and r == rate
        b.set_frontend(cid, freq, float(cutoff), outr)
        inc, nt, taps = b.frontend_info(cid)
        key = f"frontend/strict/{freq}_{rate}"
        assert inc == golden_meta["frontend_inc"][f"{freq}_{rate}"]
        assert nt == 72 and np.array_equal(taps[: 16 * 72].reshape(16, 72), golden[key + "/taps"])
        outs = []
        for a, e in ((0, 7), (7, 9000), (9000, 20000)):
            b.feed(fx[a:e])
            assert np.array_equal(b.fetch(cid), fx[a:e])
            outs.append(b.fetch(cid, capi.STAGE_FRONTEND).copy())
        out = np.concatenate(outs)
        want = golden[key + "/out"]
        assert out.shape == want.shape          # same schedule length: the float32 distance recurrence is replayed exactly
        assert rel_rms(out, want) <= 1e-5
        assert rel_rms(out, golden[f"frontend/fast/{freq}_{rate}/out"]) <= 1e-5
        b.close()


def test_bank_with_frontends_vs_oracle(gpu_lib, port, golden_meta):
    """Config-3 style channels end to end: HB48 tree -> NCO(-residual) -> Interpolator to 48 kS/s, vs oracle chains."""
    from sdrangel_b200 import DownChannelizerBank, capi
    plan = golden_meta["chan_plans"]["bank64"]
    rs = np.random.RandomState(44)
    n = 2_000_000
    x = (rs.randint(-6000, 6000, size=(n, 2))).astype(np.int16)
    b = DownChannelizerBank(plan["input_rate"])
    chans = plan["channels"][3::16]
    cutoff = np.float32(np.float32(12500) / np.float32(2.2))
    ids = []
    for fc, rate, ofs, path in chans:
        cid = b.add_channel(48000, fc)[0]
        b.set_frontend(cid, -ofs, float(cutoff), 48000)
        ids.append(cid)
    refs = []
    for fc, rate, ofs, path in chans:
        o = port.PortDownChannelizer()
        o.configure(plan["input_rate"], 48000, fc)
        refs.append((o, port.PortFrontEnd(-ofs, rate, 48000, cutoff)))
    for a, e in ((0, 1_000_001), (1_000_001, n)):
        b.feed(x[a:e])
        for cid, (o, fe) in zip(ids, refs):
            ch = o.feed(x[a:e])
            assert np.array_equal(b.fetch(cid), ch)
            want = fe.feed(ch)
            got = b.fetch(cid, capi.STAGE_FRONTEND)
            assert got.shape == want.shape
            assert rel_rms(got, want) <= 1e-5
        # the pooled fetch hands out exactly what the per-channel fetch does, for both stages
        ch_all, ch_cnt = b.fetch_all(capi.STAGE_CHANNELIZER)
        fe_all, fe_cnt = b.fetch_all(capi.STAGE_FRONTEND, stride=ch_all.shape[1])
        for cid in ids:
            one = b.fetch(cid)
            assert ch_cnt[cid] == one.shape[0] and np.array_equal(ch_all[cid, :ch_cnt[cid]], one)
            onef = b.fetch(cid, capi.STAGE_FRONTEND)
            assert fe_cnt[cid] == onef.shape[0] and np.array_equal(fe_all[cid, :fe_cnt[cid]], onef)
    with pytest.raises(RuntimeError):
        b.fetch_all(capi.STAGE_FRONTEND, stride=1)             # stride too small: loud, not truncated


@pytest.mark.parametrize("rate,outr", [(156250, 48000), (60000, 48000), (78125, 48000), (48000, 48000)])
def test_standalone_interpolator_vs_oracle(gpu_lib, port, golden, rate, outr):
    """b200dsp_interp_decimate == the plugin loop around Interpolator::decimate, fed in ragged blocks with the caller-owned
    distance carried between calls.  Oracle: the reference front-end with a zero-frequency NCO (mix by exactly 1)."""
    from sdrangel_b200 import Interpolator
    rs = np.random.RandomState(rate)
    n = 50_000
    xi = rs.randint(-20000, 20000, size=(n, 2)).astype(np.int16)
    x = (xi[:, 0].astype(np.float32) + 1j * xi[:, 1].astype(np.float32)).astype(np.complex64)
    cutoff = np.float32(np.float32(12500) / np.float32(2.2))
    it = Interpolator(16, rate, float(cutoff))
    fe = port.PortFrontEnd(0, rate, outr, cutoff)
    assert np.array_equal(it.taps(), fe.taps())
    step = float(np.float32(np.float32(rate) / np.float32(outr)))
    remain = 0.0
    cuts = [0, 1, 2, 1000, 1001, 30_000, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        got, remain = it.decimate(remain, step, x[a:b])
        want = fe.feed(xi[a:b])
        assert got.shape[0] == want.shape[0], (a, b)
        if got.size:
            assert rel_rms(got.view(np.float32), want) <= 1e-5


def _plain_bank(input_rate):
    """A bank that never takes the fused multi-level kernel (every pass through hb48_level_kernel)."""
    import os
    from sdrangel_b200 import DownChannelizerBank
    os.environ["B200DSP_NO_FUSED_TREE"] = "1"
    try:
        return DownChannelizerBank(input_rate)
    finally:
        del os.environ["B200DSP_NO_FUSED_TREE"]


def test_bank_fused_tree_aligned_and_mixed_feeds(gpu_lib, port, golden_meta):
    """Passes that start aligned at every level and are a multiple of 2^depth long take the fused multi-level kernel
    (hb48_fused_kernel: levels of a group in shared memory), anything else the one-level kernel; both share the carried
    tails.  A stream that alternates between them, with runs of -32768 and full-scale noise, must stay bit-exact against
    one oracle chain per channel, call by call, and identical to a bank that never fuses."""
    from sdrangel_b200 import DownChannelizerBank
    plan = golden_meta["chan_plans"]["bank64"]
    rs = np.random.RandomState(77)
    unit = 256 * 768                     # a multiple of 2^(depth+1) for the 7-level tree; 64 tiles of the top group
    sizes = [unit * 3, 256, unit, 1000, unit * 2, 7, unit * 2 - 7 - 1000, unit * 4 + 128, 128, unit]
    n = sum(sizes)
    x = rs.randint(-32768, 32768, size=(n, 2)).astype(np.int16)
    x[5000:5100] = -32768
    x[unit * 3 + 500: unit * 3 + 600, 0] = -32768
    x[unit * 5: unit * 5 + 3000] = -32768
    x[n - unit - 40: n - unit + 40, 1] = -32768            # across a pass boundary: the carried tails hold -32768
    chans = plan["channels"][::7]
    plain = _plain_bank(plan["input_rate"])
    fused = DownChannelizerBank(plan["input_rate"])
    ids = [fused.add_channel(48000, fc)[0] for fc, _, _, _ in chans]
    for fc, _, _, _ in chans:
        plain.add_channel(48000, fc)
    oracles = []
    for fc, _, _, _ in chans:
        o = port.PortDownChannelizer()
        o.configure(plan["input_rate"], 48000, fc)
        oracles.append(o)
    pos = 0
    for sz in sizes:
        blk = x[pos:pos + sz]
        pos += sz
        fused.feed(blk)
        plain.feed(blk)
        for cid, o in zip(ids, oracles):
            want = o.feed(blk)
            got = fused.fetch(cid)
            assert got.shape == want.shape, (sz, cid)
            assert np.array_equal(got, want), (sz, cid, int(np.argmax(np.any(got != want, axis=1))))
            assert np.array_equal(plain.fetch(cid), want), (sz, cid)
    fused.close()
    plain.close()


@pytest.mark.parametrize("chunk", [3 << 22, 768 * 64])
def test_bank1024_every_channel_vs_oracle(gpu_lib, port, golden_meta, chunk):
    """The whole 1024-channel plan (BASELINE config 5), every channel: two feeds (2^18 then 2^17 + 2^12 samples) against one
    oracle DownChannelizer per channel (downchannelizer.cpp:50-91), tree bit-exact; the front-end of every 16th channel
    within 1e-5.  With the small chunk the second feed runs as several internal passes."""
    from sdrangel_b200 import DownChannelizerBank, capi
    plan = golden_meta["chan_plans"]["bank1024"]
    fs = plan["input_rate"]
    rows = plan["channels"]
    assert len(rows) == 1024
    rs = np.random.RandomState(1024)
    n1, n2 = 1 << 18, (1 << 17) + (1 << 12)
    x = rs.randint(-32768, 32768, size=(n1 + n2, 2)).astype(np.int16)
    x[100_000:100_300] = -32768
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    b = DownChannelizerBank(fs)
    b.set_chunk(chunk)
    for fc, rate, ofs, path in rows:
        cid = b.add_channel(48000, fc)[0]
        b.set_frontend(cid, -ofs, cutoff, 48000)
    assert b.node_count() == 3070
    oracles = []
    for fc, rate, ofs, path in rows:
        o = port.PortDownChannelizer()
        o.configure(fs, 48000, fc)
        oracles.append(o)
    fes = {i: port.PortFrontEnd(-rows[i][2], rows[i][1], 48000, cutoff) for i in range(0, 1024, 16)}
    for a, e in ((0, n1), (n1, n1 + n2)):
        b.feed(x[a:e])
        ch_all, ch_cnt = b.fetch_all(capi.STAGE_CHANNELIZER)
        fe_all, fe_cnt = b.fetch_all(capi.STAGE_FRONTEND, stride=ch_all.shape[1])
        bad = []
        for i, o in enumerate(oracles):
            want = o.feed(x[a:e])
            if ch_cnt[i] != want.shape[0] or not np.array_equal(ch_all[i, :ch_cnt[i]], want):
                bad.append(i)
            if i in fes:
                wf = fes[i].feed(want)
                assert fe_cnt[i] == wf.shape[0], (a, i)
                assert rel_rms(fe_all[i, :fe_cnt[i]], wf) <= 1e-5, (a, i)
        assert not bad, (a, len(bad), bad[:8])
    b.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_cooperative_bank_emulated_ranks_bit_exact(gpu_lib, port, golden_meta, world):
    """The cooperative multi-GPU scheme (time-sliced top levels with halos -> all-to-all of node streams -> per-rank banks by
    path suffix), with the ranks emulated one after the other on one GPU and the exchange done by array copies: every
    channel's output must equal one oracle chain over the whole stream, step by step (tree bit-exact, front-end <= 1e-5)."""
    from sdrangel_b200 import capi
    from sdrangel_b200.coop import CoopPlan, CoopRank, HALO
    plan1024 = golden_meta["chan_plans"]["bank1024"]
    fs = plan1024["input_rate"]
    rows = plan1024["channels"][3::32]                      # 32 channels spread over the band
    fcs = [r[0] for r in rows]
    n = 8 * 768 * 8 * 4
    steps = 3
    rs = np.random.RandomState(world)
    X = rs.randint(-32768, 32768, size=(steps * n, 2)).astype(np.int16)
    X[n - 50:n + 50] = -32768                               # a -32768 run across a step boundary
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    plan = CoopPlan(fs, fcs, 48000, world, n)
    assert [c[2] for c in plan.chains] == [r[3] for r in rows]
    ranks = [CoopRank(plan, r, frontend=(cutoff, 48000)) for r in range(world)]
    oracles = []
    for fc, rate, ofs, path in rows:
        o = port.PortDownChannelizer()
        o.configure(fs, 48000, fc)
        oracles.append((o, port.PortFrontEnd(-ofs, rate, 48000, cutoff)))
    Xpad = np.concatenate([np.zeros((HALO, 2), np.int16), X])
    for t in range(steps):
        node_out = []
        for r, rk in enumerate(ranks):
            a = t * n + r * plan.m                           # slice start in X; Xpad is shifted by HALO
            rk.top.reset()
            rk.top.feed(Xpad[a:a + HALO + plan.m])
            outs = {}
            for v, cid in rk.top_ids.items():
                o = rk.top.fetch(cid)
                assert o.shape[0] == (HALO + plan.m) >> plan.k
                outs[v] = o[plan.skip:].copy()
            node_out.append(outs)
        for q, rk in enumerate(ranks):
            for v, bank in rk.subs.items():
                bank.feed(np.concatenate([node_out[r][v] for r in range(world)]))
        for i, (o, fe) in enumerate(oracles):
            q = next(r for r, (lo, hi) in enumerate(plan.ranges) if lo <= i < hi)
            node, cid = ranks[q].chan[i]
            want = o.feed(X[t * n:(t + 1) * n])
            got = ranks[q].subs[node].fetch(cid)
            assert got.shape == want.shape, (t, i)
            assert np.array_equal(got, want), (t, i, int(np.argmax(np.any(got != want, axis=1))))
            wf = fe.feed(want)
            gf = ranks[q].subs[node].fetch(cid, capi.STAGE_FRONTEND)
            assert gf.shape == wf.shape and rel_rms(gf, wf) <= 1e-5, (t, i)
    for rk in ranks:
        rk.close()


@pytest.mark.gpu
def test_cooperative_bank_device_path_wrapped_halo(gpu_lib, port, golden_meta):
    """The same scheme through the device-pointer calls bench.py uses at N > 1 (reset / feed_dev / copy_out_dev on a caller
    stream), two emulated ranks, the stream being one buffer repeated (slice 0's halo is the buffer's end): the second
    step must equal oracle chains fed the buffer twice."""
    import torch
    from sdrangel_b200 import capi
    from sdrangel_b200.coop import CoopPlan, CoopRank, HALO
    world = 2
    plan1024 = golden_meta["chan_plans"]["bank1024"]
    fs = plan1024["input_rate"]
    rows = plan1024["channels"][5::64]
    fcs = [r[0] for r in rows]
    n = world * 768 * 8 * 16
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    x = torch.randint(-2048, 2048, (2 * n,), dtype=torch.int16, device=dev, generator=g)
    hx = x.cpu().numpy().reshape(-1, 2)
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    p = CoopPlan(fs, fcs, 48000, world, n)
    ranks = [CoopRank(p, r, frontend=(cutoff, 48000)) for r in range(world)]
    xi = x.view(torch.int32)
    xcat = torch.cat([xi[-HALO:], xi])
    stream = torch.cuda.Stream(device=dev)
    sptr = stream.cuda_stream
    torch.cuda.synchronize()
    node_buf = {v: torch.empty((world, p.mk), dtype=torch.int32, device=dev) for v in p.nodes}
    with torch.cuda.stream(stream):
        for step in range(2):
            # first step: nothing precedes the stream (zero history, as the oracle); then the buffer's end precedes its start
            xcat[:HALO] = 0 if step == 0 else xi[-HALO:]
            for r, rk in enumerate(ranks):
                sl = xcat[r * p.m: r * p.m + HALO + p.m].clone()
                rk.top.reset(sptr)
                rk.top.feed_dev(sl.data_ptr(), HALO + p.m, sptr)
                for v in p.nodes:
                    rk.top.copy_out_dev(rk.top_ids[v], p.skip, p.mk, node_buf[v][r].data_ptr(), sptr)
            for rk in ranks:
                for v, bank in rk.subs.items():
                    bank.feed_dev(node_buf[v].data_ptr(), p.n >> p.k, sptr)
    stream.synchronize()
    for i, (fc, rate, ofs, path) in enumerate(rows):
        o = port.PortDownChannelizer()
        o.configure(fs, 48000, fc)
        fe = port.PortFrontEnd(-ofs, rate, 48000, cutoff)
        fe.feed(o.feed(hx))
        ch = o.feed(hx)
        want = fe.feed(ch)
        q = next(r for r, (lo, hi) in enumerate(p.ranges) if lo <= i < hi)
        node, cid = ranks[q].chan[i]
        got = ranks[q].subs[node].fetch(cid)
        assert got.shape == ch.shape, i
        assert np.array_equal(got, ch), (i, int(np.argmax(np.any(got != ch, axis=1))))
        gf = ranks[q].subs[node].fetch(cid, capi.STAGE_FRONTEND)
        assert gf.shape == want.shape and rel_rms(gf, want) <= 1e-5, i
    for rk in ranks:
        rk.close()


@pytest.mark.gpu
@pytest.mark.parametrize("plan_name,chunk", [("bank64", 2304), ("bank1024", 768 * 5), ("bank64", 100_000)])
def test_bank_frontends_many_internal_passes(gpu_lib, port, golden_meta, plan_name, chunk):
    """A feed larger than the bank's chunk runs as several internal passes: the front-end's per-feed output count, its
    schedule state (exact replay for bank64's ratios, closed form for bank1024's 1.25), the NCO phase and the 128-sample
    history must carry from pass to pass exactly as they do from feed to feed."""
    from sdrangel_b200 import DownChannelizerBank, capi
    plan = golden_meta["chan_plans"][plan_name]
    fs = plan["input_rate"]
    rows = plan["channels"][2::max(1, len(plan["channels"]) // 6)][:6]
    rs = np.random.RandomState(chunk)
    n = 260_000
    x = rs.randint(-20000, 20000, size=(n, 2)).astype(np.int16)
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    b = DownChannelizerBank(fs)
    b.set_chunk(chunk)
    ids, refs = [], []
    for fc, rate, ofs, path in rows:
        cid = b.add_channel(48000, fc)[0]
        b.set_frontend(cid, -ofs, cutoff, 48000)
        ids.append(cid)
        o = port.PortDownChannelizer()
        o.configure(fs, 48000, fc)
        refs.append((o, port.PortFrontEnd(-ofs, rate, 48000, cutoff)))
    for a, e in ((0, 130_001), (130_001, 130_002), (130_002, n)):
        b.feed(x[a:e])
        for cid, (o, fe) in zip(ids, refs):
            ch = o.feed(x[a:e])
            assert np.array_equal(b.fetch(cid), ch), (a, cid)
            want = fe.feed(ch)
            got = b.fetch(cid, capi.STAGE_FRONTEND)
            assert got.shape == want.shape, (a, cid, got.shape, want.shape)
            if want.size:
                assert rel_rms(got, want) <= 1e-5, (a, cid)
    b.close()


def test_bank_process_equals_feed_then_fetch_all(gpu_lib, port, golden_meta):
    """b200dsp_bank_process (host buffer in, every channel's outputs out, H2D / kernels / D2H of finished columns overlapped
    over several internal passes) hands out exactly what b200dsp_bank_feed + b200dsp_bank_fetch_all do, for both stages, and
    the tree outputs equal the oracle's -- ragged and aligned block sizes, mixed channel depths (bank64: S = 6 and 7)."""
    import torch
    from sdrangel_b200 import DownChannelizerBank, capi
    plan = golden_meta["chan_plans"]["bank64"]
    fs = plan["input_rate"]
    rows = plan["channels"][1::6]
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    rs = np.random.RandomState(5)
    sizes = [1 << 18, 100_003, 3 * (1 << 16)]
    x = rs.randint(-32768, 32768, size=(sum(sizes), 2)).astype(np.int16)
    a, b = DownChannelizerBank(fs), DownChannelizerBank(fs)
    for bk in (a, b):
        bk.set_chunk(768 * 32)
        for fc, rate, ofs, path in rows:
            cid = bk.add_channel(48000, fc)[0]
            bk.set_frontend(cid, -ofs, cutoff, 48000)
    oracles = []
    for fc, rate, ofs, path in rows:
        o = port.PortDownChannelizer()
        o.configure(fs, 48000, fc)
        oracles.append(o)
    pos = 0
    for sz in sizes:
        blk = np.ascontiguousarray(x[pos:pos + sz])
        pos += sz
        hx = torch.from_numpy(blk.reshape(-1)).pin_memory()
        a.feed(blk)
        ch_all, ch_cnt = a.fetch_all(capi.STAGE_CHANNELIZER)
        fe_all, fe_cnt = a.fetch_all(capi.STAGE_FRONTEND, stride=ch_all.shape[1])
        stride = ch_all.shape[1] + 8
        for stage, want, wcnt, dt in ((capi.STAGE_FRONTEND, fe_all, fe_cnt, torch.float32),):
            out = torch.zeros((len(rows), stride, 2), dtype=dt).pin_memory()
            cnt = b.process(hx.data_ptr(), sz, stage, out.data_ptr(), stride)
            assert np.array_equal(cnt, wcnt), (sz, stage)
            got = out.numpy()
            for i in range(len(rows)):
                assert np.array_equal(got[i, :cnt[i]], want[i, :wcnt[i]]), (sz, stage, i)
        for i, o in enumerate(oracles):
            w = o.feed(blk)
            assert np.array_equal(b.fetch(i), w), (sz, i)
    # the channelizer stage through process() on a fresh pair of banks
    c2, d2 = DownChannelizerBank(fs), DownChannelizerBank(fs)
    for bk in (c2, d2):
        bk.set_chunk(768 * 16)
        for fc, rate, ofs, path in rows:
            bk.add_channel(48000, fc)
    blk = np.ascontiguousarray(x[:200_001])
    hx = torch.from_numpy(blk.reshape(-1)).pin_memory()
    c2.feed(blk)
    want, wcnt = c2.fetch_all(capi.STAGE_CHANNELIZER)
    out = torch.zeros((len(rows), want.shape[1] + 4, 2), dtype=torch.int16).pin_memory()
    cnt = d2.process(hx.data_ptr(), blk.shape[0], capi.STAGE_CHANNELIZER, out.data_ptr(), want.shape[1] + 4)
    assert np.array_equal(cnt, wcnt)
    for i in range(len(rows)):
        assert np.array_equal(out.numpy()[i, :cnt[i]], want[i, :wcnt[i]]), i
    for bk in (a, b, c2, d2):
        bk.close()


def test_one_gpu_equals_two_gpus(gpu_lib, golden_meta):
    """SURVEY.md 8(d) gate '1-GPU == G-GPU': the same feed through one bank holding every channel on device 0 and through two
    banks holding the two frequency blocks on devices 0 and 1 (b200dsp_dist_shard) gives byte-identical channel outputs and
    identical front-end outputs.  Needs two devices (the driver's 1-GPU run skips it; run under `gpurun --gpus 2`)."""
    import ctypes as C
    from sdrangel_b200 import DownChannelizerBank, capi
    if capi.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    plan = golden_meta["chan_plans"]["bank1024"]
    fs = plan["input_rate"]
    rows = plan["channels"][::16]
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    rs = np.random.RandomState(8)
    x = rs.randint(-32768, 32768, size=((1 << 18) + 4096, 2)).astype(np.int16)
    whole = DownChannelizerBank(fs, device=0)
    for fc, rate, ofs, path in rows:
        cid = whole.add_channel(48000, fc)[0]
        whole.set_frontend(cid, -ofs, cutoff, 48000)
    parts = []
    for r in range(2):
        lo, hi = C.c_int32(), C.c_int32()
        capi.check(capi.lib().b200dsp_dist_shard(len(rows), 2, r, C.byref(lo), C.byref(hi)))
        bk = DownChannelizerBank(fs, device=r)
        for fc, rate, ofs, path in rows[lo.value:hi.value]:
            cid = bk.add_channel(48000, fc)[0]
            bk.set_frontend(cid, -ofs, cutoff, 48000)
        parts.append((bk, lo.value, hi.value))
    for a, e in ((0, 1 << 18), (1 << 18, x.shape[0])):
        whole.feed(x[a:e])
        for bk, lo, hi in parts:
            bk.feed(x[a:e])
            for i in range(lo, hi):
                assert np.array_equal(bk.fetch(i - lo), whole.fetch(i)), (a, i)
                assert np.array_equal(bk.fetch(i - lo, capi.STAGE_FRONTEND), whole.fetch(i, capi.STAGE_FRONTEND)), (a, i)
    capi.init(0)
    whole.close()
    for bk, _, _ in parts:
        bk.close()


def _sharded_worker(rank, world, nccl_id, fs, rows, q, qs):
    import torch
    sys_path_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import sys
    sys.path.insert(0, sys_path_root)
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(sys_path_root, "oracle"), "port"])
    from oracle import portbind
    from sdrangel_b200 import ShardedBank, capi
    capi.init(rank)
    torch.cuda.set_device(rank)
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    n = 1 << 18
    rs = np.random.RandomState(3)
    X = rs.randint(-32768, 32768, size=(4 * n, 2)).astype(np.int16)          # every rank can build the expected input itself
    sb = ShardedBank(fs, [r[0] for r in rows], 48000, rank, world, nccl_id, frontend=(cutoff, 48000))
    oracles = []
    for fc, rate, ofs, path in rows[sb.lo:sb.hi]:
        o = portbind.PortDownChannelizer()
        o.configure(fs, 48000, fc)
        oracles.append((o, portbind.PortFrontEnd(-ofs, rate, 48000, cutoff)))
    ok = True
    # block 0: NCCL broadcast of rank 0's device buffer; block 1: sliced host ingest + all-gather
    xd = torch.from_numpy(X[:n].reshape(-1)).cuda() if rank == 0 else None
    sb.bcast_begin(0, xd.data_ptr() if rank == 0 else 0, n, 0, None)
    sb.feed(0)
    sb.bank.sync()
    for blk, lo_s in ((0, 0), (1, n)):
        if blk == 1:
            cnt = n // world
            hs = torch.from_numpy(np.ascontiguousarray(X[n + rank * cnt: n + (rank + 1) * cnt]).reshape(-1)).pin_memory()
            sb.ingest_begin(1, hs.data_ptr(), n)
            sb.feed(1)
            sb.bank.sync()
        for k, (o, fe) in enumerate(oracles):
            ch = o.feed(X[lo_s:lo_s + n])
            ok = ok and np.array_equal(sb.bank.fetch(k), ch)
            wf = fe.feed(ch)
            gf = sb.bank.fetch(k, capi.STAGE_FRONTEND)
            ok = ok and gf.shape == wf.shape and rel_rms(gf, wf) <= 1e-5
    # blocks 2 and 3: the copy-engine chain (IPC slots, stream-ordered counters); the blobs travel through the parent's queues
    blob = sb.p2p_export(n)
    q.put(("blob", rank, blob))
    blobs = qs[rank].get(timeout=120)
    sb.p2p_import(blobs)
    for blk in (2, 3):
        xd = torch.from_numpy(X[blk * n:(blk + 1) * n].reshape(-1)).cuda() if rank == 0 else None
        torch.cuda.synchronize()
        sb.p2p_begin(blk % 3, xd.data_ptr() if rank == 0 else 0, n)
        sb.p2p_feed(blk % 3)
        sb.bank.sync()
        sb.sync()
        for k, (o, fe) in enumerate(oracles):
            ch = o.feed(X[blk * n:(blk + 1) * n])
            ok = ok and np.array_equal(sb.bank.fetch(k), ch)
            wf = fe.feed(ch)
            gf = sb.bank.fetch(k, capi.STAGE_FRONTEND)
            ok = ok and gf.shape == wf.shape and rel_rms(gf, wf) <= 1e-5
    q.put(("done", rank, bool(ok), sb.lo, sb.hi))
    sb.close()


def test_sharded_bank_two_ranks_nccl(gpu_lib, golden_meta):
    """K6 through the library object: two processes, one GPU each, b200dsp_dist_* (NCCL broadcast of a device-resident block,
    then a host-fed block by per-rank slices + all-gather), every rank's channels against oracle chains fed the same stream.
    Needs two devices (run under `gpurun --gpus 2`)."""
    import torch.multiprocessing as mp
    from sdrangel_b200 import ShardedBank, capi
    if capi.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    plan = golden_meta["chan_plans"]["bank1024"]
    rows = plan["channels"][5::32]
    nccl_id = ShardedBank.unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    qs = [ctx.Queue() for _ in range(2)]
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, nccl_id, plan["input_rate"], rows, q, qs)) for r in range(2)]
    for p in procs:
        p.start()
    blobs, res = {}, []
    while len(res) < 2:
        msg = q.get(timeout=300)
        if msg[0] == "blob":
            blobs[msg[1]] = msg[2]
            if len(blobs) == 2:
                for r in range(2):
                    qs[r].put([blobs[0], blobs[1]])
        else:
            res.append(msg[1:])
    res.sort()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True], res
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == len(rows)


@pytest.mark.parametrize("lo,hi", [(0, 128), (512, 1024), (960, 992)])
def test_bank_shard_chain_cascade_vs_oracle(gpu_lib, port, golden_meta, lo, hi):
    """A rank's frequency block of the sharded 1024-channel bank: the top levels of its tree are single-child (3, 1 and 5
    of them here) and run as one warp-private cascade (hb48_chain_kernel) ahead of the fused pyramid launches.  Aligned feeds
    with runs of -32768 (also across a feed boundary, so the carried tails hold them) and a ragged one in between, every
    16th channel of the block against its oracle chain and against a bank that only uses the one-level kernel."""
    from sdrangel_b200 import DownChannelizerBank
    plan = golden_meta["chan_plans"]["bank1024"]
    fs = plan["input_rate"]
    rows = plan["channels"][lo:hi]
    rs = np.random.RandomState(lo + 1)
    unit = 3 << 17
    sizes = [unit, unit + 4096, 1234, unit, 4096]
    n = sum(sizes)
    x = rs.randint(-32768, 32768, size=(n, 2)).astype(np.int16)
    x[7000:7100] = -32768
    x[unit - 30:unit + 30, 0] = -32768
    x[2 * unit + 5000:2 * unit + 9000] = -32768
    fused, plain = DownChannelizerBank(fs), _plain_bank(fs)
    for fc, rate, ofs, path in rows:
        fused.add_channel(48000, fc)
        plain.add_channel(48000, fc)
    picks = list(range(0, len(rows), 16)) + [len(rows) - 1]
    oracles = {}
    for i in picks:
        o = port.PortDownChannelizer()
        o.configure(fs, 48000, rows[i][0])
        oracles[i] = o
    pos = 0
    for sz in sizes:
        blk = x[pos:pos + sz]
        pos += sz
        fused.feed(blk)
        plain.feed(blk)
        for i, o in oracles.items():
            want = o.feed(blk)
            got = fused.fetch(i)
            assert got.shape == want.shape, (sz, i)
            assert np.array_equal(got, want), (sz, i, int(np.argmax(np.any(got != want, axis=1))))
            assert np.array_equal(plain.fetch(i), want), (sz, i)
    fused.close()
    plain.close()


@pytest.mark.gpu
def test_frontend_schedule_indices_and_phases_equal_oracle(gpu_lib, port, golden_meta):
    """SURVEY.md 8d's K4 gate checked directly: for every front-end output the index of the channel sample that emitted it and
    the polyphase phase equal what the reference's float32 distance recurrence decides -- closed-form (1.25), exact parallel scan
    (3.2552, 1.6276) and a ratio below 2^21 ulps -- over several feeds (the distance carries), incl. feeds without new samples."""
    from sdrangel_b200 import DownChannelizerBank
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    rs = np.random.RandomState(31)
    for fs, fc, outr in ((122_880_000, 61_381_250, 48000), (10_000_000, 1_234_567, 48000), (10_000_000, -3_000_000, 48000), (2_000_000, 100_000, 44100)):
        b = DownChannelizerBank(fs)
        cid, rate, ofs, path = b.add_channel(48000, fc)
        b.set_frontend(cid, -ofs, cutoff, outr)
        oc = port.PortDownChannelizer(); oc.configure(fs, 48000, fc)
        fe = port.PortFrontEnd(-ofs, rate, outr, cutoff)
        S = len(path)
        # (every feed is one internal pass: an aligned feed of a whole number of 2^(S+1) blocks, then ragged ones below the chunk)
        for n in (4 << S, 1, (2_000 << S) + 5, 0, 7, (5_001 << S)):
            x = rs.randint(-8000, 8000, size=(n, 2)).astype(np.int16)
            b.feed(x)
            _, widx, wph = fe.feed(oc.feed(x), want_schedule=True)
            gidx, gph = b.fetch_schedule(cid)
            assert gidx.shape == widx.shape, (fs, fc, n)
            assert np.array_equal(gidx, widx) and np.array_equal(gph, wph), (fs, fc, n)
        b.close()
