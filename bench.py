#!/usr/bin/env python
"""bench.py — input MS/s of the SDRangel baseband-to-channel hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

Workloads (BASELINE.json configs):
  decimateii   sdrbench decimateii: int16 IQ, log2 decimation 4, centred, 12 bit          (config 1; default at N=1)
  decimatefi   sdrbench decimatefi: float IQ, log2 decimation 6, centred                  (config 2)
  bank64       64 NFM channels off a 10 MS/s baseband: HB48 tree + NCO + Interpolator     (config 3)
  spectrum     SpectrumVis 4096-pt Blackman-Harris FFT, log power, fixed average of 10 frames (config 4)
  bank1024     1024 channels over a 122.88 MS/s stream, channels sharded over the ranks,
               baseband NCCL-broadcast from rank 0 every step                             (config 5; default at N>1)
A "step" is one pass of the hot path over one batch of synthetic IQ that is larger than L2 (every step streams from
HBM).  `value` is device-resident throughput (inputs already in HBM), `e2e` goes through the host-pointer C-ABI call
with H2D/D2H inside the timed region.  The single-stream decimators do not shard ("replicas only"): with
N > 1 every rank runs an independent replica (weak scaling); the bank shards by channel (strong scaling: the same
stream, 1024/N channels per GPU).  --impl reference times the reference's own CPU code (oracle/_ref, compiled from
the reference sources) on the host cores, on the same metric.  At N=1 the default line also carries the other
workloads' numbers under "also".
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "input MS/s: sdrbench decimators + N-channel DownChannelizer bank, 1/2/4/8 GPU"


def plan64():
    return 10_000_000, [((c - 32) * 150000 + 6250) for c in range(64)]


def plan1024():
    return 122_880_000, [((c - 512) * 120000 + 60000 + 1250 * ((c % 9) - 4)) for c in range(1024)]


WORKLOADS = {
    "decimateii": dict(type="decim", kind="ii", log2=4, mode=2, in_dtype="int16", in_bytes=4, out_bytes=4 / 16, n=1 << 28,
                       desc="sdrbench decimateii: Decimators<qint32,qint16,16,12>::decimate16_cen, synthetic int16 IQ"),
    "decimatefi": dict(type="decim", kind="fi", log2=6, mode=2, in_dtype="float32", in_bytes=8, out_bytes=4 / 64, n=1 << 27,
                       desc="sdrbench decimatefi: DecimatorsFI::decimate64_cen, synthetic float IQ"),
    "bank64": dict(type="bank", plan=plan64, n=1 << 26,        # 2^26 samples (6.7 s of baseband; 256 MiB > L2)
                   desc="64 NFM 12.5 kHz channels off a synthetic 10 MS/s int16 baseband: DownChannelizer tree + NCO + Interpolator to 48 kS/s"),
    "spectrum": dict(type="spectrum", n=1 << 26, fft=4096, avg_nb=10, avg_mode=2,
                     desc="SpectrumVis: 4096-pt Blackman-Harris windowed FFT, log power, fixed averaging over 10 frames, synthetic int16 IQ (61.44 MS/s LimeSDR-rate stream)"),
    "iqcorr": dict(type="iqcorr", n=1 << 27,
                   desc="DSPDeviceSourceEngine::iqCorrections, DC branch (1024-sample moving average removed per component), synthetic int16 IQ"),
    "interps": dict(type="tx", kind="interps", bits=12, log2=5, n=1 << 22,
                    desc="Tx: Interpolators<qint16,16,12>::interpolate32_cen, SampleVector -> device buffer (LimeSDR / BladeRF sink), synthetic int16 IQ"),
    "upchan": dict(type="tx", kind="upchan", plan=(3_072_000, 48_000, 300_000), n=1 << 27,
                   desc="Tx: UpChannelizer::pull, 48 kS/s modulator -> 3.072 MS/s (6 interpolating half-bands of order 96), n = output samples"),
    "ssbfilt": dict(type="tx", kind="fftfilt", flen=1024, n=1 << 25,
                    desc="SSB back-end: fftfilt::runSSB (1024-point overlap-add FFT filter, 300-3000 Hz at 48 kS/s) on one complex64 stream"),
    "demod": dict(type="tx", kind="demod", channels=1024, n=1 << 26,
                  desc="NFM back-end: PhaseDiscriminators::phaseDiscriminatorDelta on the pooled front-end outputs of 1024 channels, n = channel samples"),
    "bank1024": dict(type="bank", plan=plan1024, n=3 << 25,       # 100.7 M samples (384 MiB) per step; SURVEY.md 8(d) runs 2^27
                     desc="1024 channels over a synthetic 122.88 MS/s int16 stream: DownChannelizer tree + NCO + Interpolator to 48 kS/s, channels sharded"),
}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index, interval=0.001, enabled=True):
        self.samples, self.reasons, self._stop, self._t = [], set(), threading.Event(), None
        self.max_mhz = None
        self.interval = interval
        self.nv = None
        if not enabled:
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.interval)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------ CPU reference
def _oracle_mod():
    from oracle import refbind, portbind
    if refbind.available():
        return "reference", refbind
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    return "port", portbind


def _run_threads(work, threads):
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0


def cpu_reference_decim(wl, seconds, threads=None):
    """The reference's own C++ (oracle/_ref) timed the way sdrbench does (mainbench.cpp:83-104): a 2^20-sample buffer
    fed repeatedly with the filter state carried; one independent decimator object per host thread."""
    kind, mod = _oracle_mod()
    n = 1 << 20
    buf = mod.sdrbench_s16(n) if wl["kind"][0] == "i" else mod.sdrbench_f32(n)
    mk = (lambda: mod.RefDecimators(wl["kind"], 12)) if kind == "reference" else (lambda: mod.PortDecimators(wl["kind"], 12))
    threads = threads or (os.cpu_count() or 1)
    objs = [mk() for _ in range(threads)]
    objs[0].run(wl["log2"], wl["mode"], buf)
    counts = [0] * threads
    t_end = time.perf_counter() + seconds

    def work(i):
        while time.perf_counter() < t_end:
            objs[i].run(wl["log2"], wl["mode"], buf)
            counts[i] += 1

    wall = _run_threads(work, threads)
    total = sum(counts) * n
    return {"value": total / wall / 1e6, "unit": "input MS/s", "cores": threads, "kind": kind,
            "sample": "%d x 2^20-sample sdrbench buffer per thread, state carried (sdrbench -r), %.1f s wall" % (max(counts), wall)}


def cpu_reference_bank(wl, seconds, threads=None):
    """One reference (DownChannelizer + NCO + Interpolator) chain per channel, channels spread over the host threads
    (the reference's thread-per-channel model, threadedbasebandsamplesink.cpp:74-78), on a subset of the plan's
    channels; the bank figure is the subset's rate scaled by subset/plan channel count (stated in `sample`)."""
    kind, mod = _oracle_mod()
    fs, fcs = wl["plan"]()
    threads = threads or (os.cpu_count() or 1)
    per_thread = 2
    sub = fcs[:: max(1, len(fcs) // (threads * per_thread))][: threads * per_thread]
    rs = np.random.RandomState(1)
    n = 1 << 18
    x = rs.randint(-2048, 2048, size=(n, 2)).astype(np.int16)
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    Chan = mod.RefDownChannelizer if kind == "reference" else mod.PortDownChannelizer
    FE = mod.RefFrontEnd if kind == "reference" else mod.PortFrontEnd
    chains = []
    for fc in sub:
        c = Chan()
        rate, ofs, _ = c.configure(fs, 48000, fc)
        chains.append((c, FE(-ofs, rate, 48000, cutoff)))
    counts = [0] * threads
    t_end = time.perf_counter() + seconds

    def work(i):
        mine = chains[i::threads]
        while time.perf_counter() < t_end:
            for c, fe in mine:
                fe.feed(c.feed(x))
            counts[i] += len(mine)

    wall = _run_threads(work, threads)
    chan_samples = sum(counts) * n                       # (channel, input sample) pairs processed
    return {"value": chan_samples / len(fcs) / wall / 1e6, "unit": "input MS/s", "cores": threads, "kind": kind,
            "sample": "%d-channel subset of the %d-channel plan, 2^18-sample feeds for %.1f s; bank rate = subset channel-samples/s / %d channels"
                      % (len(sub), len(fcs), wall, len(fcs))}


def cpu_reference_spectrum(wl, seconds, threads=None):
    """The reference's FFTWindow + KissFFT + averaging glue (oracle/_ref), one independent SpectrumVis per host thread."""
    kind, mod = _oracle_mod()
    threads = threads or (os.cpu_count() or 1)
    rs = np.random.RandomState(2)
    n = wl["fft"] * 64
    x = rs.randint(-2048, 2048, size=(n, 2)).astype(np.int16)
    Spec = mod.RefSpectrumVis if kind == "reference" else mod.PortSpectrumVis
    objs = []
    for _ in range(threads):
        s = Spec()
        s.configure(wl["fft"], 0, wl["avg_nb"], wl["avg_mode"], 1, False)
        objs.append(s)
    counts = [0] * threads
    t_end = time.perf_counter() + seconds

    def work(i):
        while time.perf_counter() < t_end:
            objs[i].feed(x)
            counts[i] += 1

    wall = _run_threads(work, threads)
    return {"value": sum(counts) * n / wall / 1e6, "unit": "input MS/s", "cores": threads, "kind": kind,
            "sample": "%d x 64-frame feeds per thread, one stream per thread, %.1f s wall" % (max(counts), wall)}


def cpu_reference_iqcorr(wl, seconds, threads=None):
    """The reference's iqCorrections loop around its own MovingAverageUtil (oracle/_ref), one engine object per host thread."""
    kind, mod = _oracle_mod()
    threads = threads or (os.cpu_count() or 1)
    n = 1 << 20
    x = np.random.RandomState(3).randint(-2048, 2048, size=(n, 2)).astype(np.int16)
    objs = [(mod.RefIQCorrections() if kind == "reference" else mod.PortIQCorrections()) for _ in range(threads)]
    counts = [0] * threads
    t_end = time.perf_counter() + seconds

    def work(i):
        while time.perf_counter() < t_end:
            objs[i].run(x)
            counts[i] += 1

    wall = _run_threads(work, threads)
    return {"value": sum(counts) * n / wall / 1e6, "unit": "input MS/s", "cores": threads, "kind": kind,
            "sample": "%d x 2^20-sample buffer per thread, state carried, %.1f s wall" % (max(counts), wall)}


def cpu_reference_tx(wl, seconds, threads=None):
    """The reference's Tx-side / demodulator classes (oracle/_ref), one object per host thread, 2^16-sample blocks."""
    kind, mod = _oracle_mod()
    threads = threads or (os.cpu_count() or 1)
    ref = (kind == "reference")
    rs = np.random.RandomState(4)
    if wl["kind"] == "interps":
        n = 1 << 14
        x = rs.randint(-32768, 32768, size=(n, 2)).astype(np.int16)
        objs = [(mod.RefInterpolators(wl["bits"]) if ref else mod.PortInterpolators(wl["bits"])) for _ in range(threads)]
        run = lambda o: o.run(wl["log2"], x)                                  # noqa: E731
    elif wl["kind"] == "upchan":
        n = 1 << 18
        objs = [(mod.RefUpChannelizer() if ref else mod.PortUpChannelizer()) for _ in range(threads)]
        for o in objs:
            o.configure(*wl["plan"])
        x = rs.randint(-32768, 32768, size=(n // 32, 2)).astype(np.int16)
        run = lambda o: o.pull(x, n)                                          # noqa: E731
    elif wl["kind"] == "fftfilt":
        n = 1 << 17
        x = ((rs.randn(n) + 1j * rs.randn(n)) * 8000).astype(np.complex64)
        mk = mod.RefFftFilt if ref else mod.PortFftFilt
        objs = [mk(0, 300 / 48000.0, 3000 / 48000.0, wl["flen"]) for _ in range(threads)]
        run = lambda o: o.run(1, x, True, False)                              # noqa: E731
    else:
        n = 1 << 18
        x = ((rs.randn(n) + 1j * rs.randn(n)) * 8000).astype(np.complex64)
        objs = [(mod.RefDemod(1, 0.25) if ref else mod.PortDemod(1, 0.25)) for _ in range(threads)]
        run = lambda o: o.run(x)                                              # noqa: E731
    counts = [0] * threads
    t_end = time.perf_counter() + seconds

    def work(i):
        while time.perf_counter() < t_end:
            run(objs[i])
            counts[i] += 1

    wall = _run_threads(work, threads)
    return {"value": sum(counts) * n / wall / 1e6, "unit": "input MS/s", "cores": threads, "kind": kind,
            "sample": "%d blocks of %d samples per thread, state carried, %.1f s wall" % (max(counts), n, wall)}


def cpu_reference(wl, seconds):
    if wl["type"] == "tx":
        return cpu_reference_tx(wl, seconds)
    if wl["type"] == "iqcorr":
        return cpu_reference_iqcorr(wl, seconds)
    if wl["type"] == "decim":
        return cpu_reference_decim(wl, seconds)
    if wl["type"] == "spectrum":
        return cpu_reference_spectrum(wl, seconds)
    return cpu_reference_bank(wl, seconds)


def run_reference(args, wl_name, wl):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    t0 = time.perf_counter()
    per = max(2.0, min(10.0, 60.0 / max(1, args.steps + args.warmup)))
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference(wl, per)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    line = {"metric": METRIC, "value": v, "unit": "input MS/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "strong" if wl["type"] == "bank" else "weak", "vs_baseline": None,
            "dtype": "f32" if (wl.get("kind") in ("fi", "ff", "if") or wl["type"] == "spectrum") else "s32", "data": "synthetic", "impl": "reference",
            "config": {"workload": wl_name, "desc": wl["desc"]},
            "cpu_baseline": {"value": v, "unit": "input MS/s", "cores": vals[-1]["cores"], "kind": vals[-1]["kind"], "sample": vals[-1]["sample"]},
            "e2e": {"value": v, "unit": "input MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ ours
class Ctx:
    pass


def setup():
    import torch
    import torch.distributed as dist
    from sdrangel_b200 import capi
    c = Ctx()
    c.torch, c.dist, c.capi = torch, dist, capi
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    if c.world > 1:
        # NCCL's kernels on a high-priority stream: when a broadcast overlaps the FIR kernels its ring CTAs are scheduled as
        # soon as a block slot frees instead of queueing behind a whole wave (a late CTA stalls the ring on every rank)
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=not os.environ.get("B200_BENCH_NCCL_NORMAL_PRIO"))
        dist.init_process_group("nccl", device_id=c.dev, pg_options=opts)
    capi.init(c.local)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    c.hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    c.peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    c.sm_count = capi.lib().b200dsp_sm_count()
    return c


def measured_traffic(name, samples_per_step, per_launch_scale=1.0):
    """roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu capture
    (profiles/r02_traffic.json for the bank: every kernel of one step; profiles/r01_traffic.json for the kernels unchanged since
    round 1), scaled from the captured sample count to this run's."""
    try:
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))[name]
        except Exception:
            t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[name]
        return (t["dram_read"] + t["dram_write"]) * (samples_per_step / t["samples"]) * per_launch_scale
    except Exception:
        return None


def barrier(c):
    if c.world > 1:
        c.dist.barrier()
    c.torch.cuda.synchronize()


def max_over_ranks(c, v):
    if c.world > 1:
        t = c.torch.tensor([v], device=c.dev, dtype=c.torch.float64)
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
        return float(t.item())
    return v


def timed_steps(c, stream, step, steps, warmup):
    """W warm-up steps, then K steps timed with CUDA events on the launching stream, barrier + synchronize on both
    sides, max over ranks.  Returns (total_ms, per-step kernel ms list, clocks).

    N = 1: NVML samples SM clock / throttle reasons every millisecond inside the timed region.  N > 1: NVML queries serialise
    on the driver across ranks and stall the querying rank's kernel launches (8 ranks sampling: 0.58 -> 2.3 ms per step;
    rank 0 alone: 0.58 -> 0.81), so the K steps are timed unperturbed and the clocks are sampled by rank 0 during an immediate
    repeat of the same K steps, whose own time is reported next to them."""
    torch = c.torch

    def region(sampler):
        barrier(c)
        if sampler:
            sampler.start()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for a, b in evs:
            a.record(stream)
            step()
            b.record(stream)
        e1.record(stream)
        barrier(c)
        clocks = sampler.stop() if sampler else None
        return max_over_ranks(c, e0.elapsed_time(e1)), [a.elapsed_time(b) for a, b in evs], clocks

    with torch.cuda.stream(stream):
        for _ in range(max(warmup, 3)):
            step()
        if c.world == 1:
            return region(ClockSampler(c.local))
        total_ms, kern_ms, _ = region(None)
        rep_ms, _, clocks = region(ClockSampler(c.local, interval=0.002, enabled=(c.rank == 0)))
        clocks["sampled"] = "by rank 0 during an immediate repeat of the K timed steps (NVML queries stall kernel launches at N > 1)"
        clocks["ms_per_step_while_sampling"] = rep_ms / steps
    return total_ms, kern_ms, clocks


def bench_decim(c, args, wl_name, wl, steps, warmup, want_e2e=True, want_parity=True):
    import sdrangel_b200 as S
    torch, capi = c.torch, c.capi
    n = args.samples or wl["n"]
    cls = {"ii": S.Decimators, "fi": S.DecimatorsFI, "ff": S.DecimatorsFF, "if": S.DecimatorsIF}[wl["kind"]]
    dec = cls(12)
    g = torch.Generator(device=c.dev)
    g.manual_seed(5489 + c.rank)
    if wl["in_dtype"] == "int16":
        x = torch.randint(-2048, 2048, (2 * n,), dtype=torch.int16, device=c.dev, generator=g)
    else:
        x = torch.rand((2 * n,), dtype=torch.float32, device=c.dev, generator=g) * 2 - 1
    n_out = dec.out_count(wl["log2"], wl["mode"], 2 * n)
    out_dt = torch.int16 if wl["kind"][1] == "i" else torch.float32
    y = torch.empty((n_out, 2), dtype=out_dt, device=c.dev)
    stream = torch.cuda.Stream(device=c.dev)
    sptr = stream.cuda_stream

    parity = None
    if c.rank == 0 and want_parity:       # oracle = checker only: first 2^20 samples, exact arithmetic flavour
        _, mod = _oracle_mod()
        from oracle import portbind
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
        m = 1 << 20
        chk = cls(12)
        if wl["kind"] != "ii":
            chk.set_exact_float(True)
        yy = torch.empty((chk.out_count(wl["log2"], wl["mode"], 2 * m), 2), dtype=out_dt, device=c.dev)
        chk.run_dev(wl["log2"], wl["mode"], x.data_ptr(), 2 * m, yy.data_ptr(), sptr)
        stream.synchronize()
        want = portbind.PortDecimators(wl["kind"], 12).run(wl["log2"], wl["mode"], x[: 2 * m].cpu().numpy())
        parity = bool(np.array_equal(yy.cpu().numpy(), want))
        chk.close()

    def step():
        dec.run_dev(wl["log2"], wl["mode"], x.data_ptr(), 2 * n, y.data_ptr(), sptr)

    total_ms, kern_ms, clocks = timed_steps(c, stream, step, steps, warmup)
    value = c.world * n * steps / (total_ms * 1e-3) / 1e6
    kern_avg_ms = float(np.mean(kern_ms))
    alg_bytes = n * (wl["in_bytes"] + wl["out_bytes"])
    achieved = alg_bytes / (kern_avg_ms * 1e-3) / 1e9
    L = wl["log2"]
    instr_per_sample = 2 * 33 * (1 - 2.0 ** -L)
    f_clk = (clocks.get("sm_mhz") or 1965) * 1e6
    issue_roof = c.sm_count * 128 * f_clk / instr_per_sample / 1e6
    res = {"value": value, "ms_per_step": total_ms / steps, "clocks": clocks, "parity": parity, "launches": steps,
           "config": {"workload": wl_name, "desc": wl["desc"], "samples_per_step": n, "log2_decim": L, "fc_pos": "cen", "input_bits": 12,
                      "l2": "input %.0f MiB per step > 126 MB L2, streamed from HBM every step" % (n * wl["in_bytes"] / 2 ** 20),
                      "parallelism": "replicas x%d (single-stream decimator does not shard)" % c.world},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": c.hbm_peak, "unit": "GB/s", "frac": achieved / c.hbm_peak,
                        "traffic": measured_traffic(wl_name, n), "peak_source": c.peak_src, "kernel": "hb64_cascade_kernel", "kernel_ms": kern_avg_ms,
                        "algorithmic_bytes_per_sample": wl["in_bytes"] + wl["out_bytes"],
                        "issue": {"instr_per_sample": instr_per_sample, "roof_MSps_at_sampled_clk": issue_roof,
                                  "frac": (n / (kern_avg_ms * 1e-3) / 1e6) / issue_roof,
                                  "note": "binding roof (DESIGN.md): 33 instructions per real stage output, 15-16 IMAD on the FMA-heavy pipe "
                                          "(64 lanes/clk/SM) + 16-17 IADD3/shift on the ALU pipe (64 lanes/clk/SM)"}},
           "dtype": "s32" if wl["kind"] == "ii" else "f32", "scaling": "weak"}
    if want_e2e:
        n_e = min(n, args.e2e_samples)
        hx = torch.empty((2 * n_e,), dtype=x.dtype, pin_memory=True)
        hx.copy_(x[: 2 * n_e])
        n_out_e = dec.out_count(wl["log2"], wl["mode"], 2 * n_e)
        hy = torch.empty((n_out_e, 2), dtype=out_dt, pin_memory=True)
        d2 = cls(12)
        L_ = capi.lib()
        nout = C.c_int32(0)

        def e2e_step():
            capi.check(L_.b200dsp_decim_run(d2._h, wl["log2"], wl["mode"], hx.data_ptr(), 2 * n_e, hy.data_ptr(), C.byref(nout)))

        for _ in range(2):
            e2e_step()
        barrier(c)
        t0 = time.perf_counter()
        ksteps = max(3, min(steps, 10))
        for _ in range(ksteps):
            e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks(c, time.perf_counter() - t0)
        res["e2e"] = {"value": c.world * n_e * ksteps / dt / 1e6, "unit": "input MS/s", "h2d_bytes_per_step": int(n_e * wl["in_bytes"]),
                      "d2h_bytes_per_step": int(n_out_e * (4 if out_dt == torch.int16 else 8)), "steps": ksteps,
                      "api": "b200dsp_decim_run (host pointers, pinned; chunked H2D/compute overlap)", "samples_per_step": n_e}
        d2.close()
    dec.close()
    return res


def bench_iqcorr(c, args, wl_name, wl, steps, warmup, want_e2e=True, want_parity=True):
    """the engine's DC correction on a device-resident int16 IQ stream (replicas at N > 1)."""
    import sdrangel_b200 as S
    torch, capi = c.torch, c.capi
    n = args.samples or wl["n"]
    g = torch.Generator(device=c.dev)
    g.manual_seed(3)
    x = torch.randint(-2048, 2048, (n, 2), dtype=torch.int16, device=c.dev, generator=g)
    y = torch.empty_like(x)
    stream = torch.cuda.Stream(device=c.dev)
    sptr = stream.cuda_stream
    q = S.IQCorrections()
    parity = None
    if c.rank == 0 and want_parity:
        _oracle_mod()
        from oracle import portbind
        m = 1 << 20
        chk = S.IQCorrections()
        chk.run_dev(x.data_ptr(), y.data_ptr(), m, sptr)
        stream.synchronize()
        parity = bool(np.array_equal(y[:m].cpu().numpy(), portbind.PortIQCorrections().run(x[:m].cpu().numpy())))
        chk.close()

    def step():
        q.run_dev(x.data_ptr(), y.data_ptr(), n, sptr)

    total_ms, kern_ms, clocks = timed_steps(c, stream, step, steps, warmup)
    k_ms = float(np.mean(kern_ms))
    achieved = n * 8.0 / (k_ms * 1e-3) / 1e9
    res = {"value": c.world * n * steps / (total_ms * 1e-3) / 1e6, "ms_per_step": total_ms / steps, "clocks": clocks, "parity": parity, "launches": steps,
           "config": {"workload": wl_name, "desc": wl["desc"], "samples_per_step": n,
                      "l2": "input %.0f MiB per step > 126 MB L2" % (n * 4 / 2 ** 20), "parallelism": "replicas x%d" % c.world},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": c.hbm_peak, "unit": "GB/s", "frac": achieved / c.hbm_peak, "traffic": None,
                        "peak_source": c.peak_src, "kernel": "dc_correct_kernel", "kernel_ms": k_ms, "algorithmic_bytes_per_sample": 8.0,
                        "issue": {"instr_per_sample": None, "roof_MSps_at_sampled_clk": None, "frac": None, "note": "HBM-bound by design: 4 B in + 4 B out per sample"}},
           "dtype": "s32", "scaling": "weak"}
    if want_e2e:
        hx = torch.empty((n, 2), dtype=torch.int16, pin_memory=True)
        hx.copy_(x)
        q2 = S.IQCorrections()
        L_ = capi.lib()

        def e2e_step():
            capi.check(L_.b200dsp_iqcorr_run(q2._h, hx.data_ptr(), n, 0))

        e2e_step()
        barrier(c)
        t0 = time.perf_counter()
        for _ in range(3):
            e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks(c, time.perf_counter() - t0)
        res["e2e"] = {"value": c.world * n * 3 / dt / 1e6, "unit": "input MS/s", "h2d_bytes_per_step": int(n * 4), "d2h_bytes_per_step": int(n * 4),
                      "steps": 3, "api": "b200dsp_iqcorr_run (pinned host buffer, corrected in place)", "samples_per_step": n}
        q2.close()
    q.close()
    return res


def bench_tx(c, args, wl_name, wl, steps, warmup, want_e2e=True, want_parity=True):
    """SURVEY.md 8f-3 / 8f-4 kernels on device-resident data (replicas at N > 1): the fused Interpolators<> cascade (value = samples
    consumed per second), UpChannelizer::pull (value = output samples per second), the pooled discriminator (channel samples)."""
    import sdrangel_b200 as S
    torch = c.torch
    n = args.samples or wl["n"]
    stream = torch.cuda.Stream(device=c.dev)
    sptr = stream.cuda_stream
    g = torch.Generator(device=c.dev)
    g.manual_seed(5)
    parity = None
    _oracle_mod()
    from oracle import portbind
    if wl["kind"] == "interps":
        L2 = wl["log2"]
        x = torch.randint(-32768, 32768, (n, 2), dtype=torch.int16, device=c.dev, generator=g)
        y = torch.empty((n << L2, 2), dtype=torch.int16, device=c.dev)
        obj = S.Interpolators(wl["bits"])
        if c.rank == 0 and want_parity:
            m = 1 << 15
            chk = S.Interpolators(wl["bits"])
            chk.run_dev(L2, x.data_ptr(), y.data_ptr(), m * (2 << L2), sptr)
            stream.synchronize()
            want, _ = portbind.PortInterpolators(wl["bits"]).run(L2, x[:m].cpu().numpy())
            parity = bool(np.array_equal(y[:m << L2].cpu().numpy().ravel(), want))
            chk.close()

        def step():
            obj.run_dev(L2, x.data_ptr(), y.data_ptr(), n * (2 << L2), sptr)
        bytes_per, kernel = 4.0 + 4.0 * (1 << L2), "interps_cascade_kernel"
        # per consumed sample: stage s produces 2^(s-1) FIR outputs of order/4 taps x (add + mad) x 2 components (+ shift, centre copy, pack)
        instr = sum((1 << s) * t * 4 for s, t in zip(range(L2), (16, 8, 4, 4, 4, 4))) + 6.0 * (1 << L2)
    elif wl["kind"] == "upchan":
        obj = S.UpChannelizer()
        obj.configure(*wl["plan"])
        need = obj.source_count(n) + 64
        x = torch.randint(-32768, 32768, (need, 2), dtype=torch.int16, device=c.dev, generator=g)
        y = torch.empty((n, 2), dtype=torch.int16, device=c.dev)
        if c.rank == 0 and want_parity:
            m = 1 << 18
            chk = S.UpChannelizer(); chk.configure(*wl["plan"])
            o = portbind.PortUpChannelizer(); o.configure(*wl["plan"])
            chk.pull_dev(x.data_ptr(), chk.source_count(m), y.data_ptr(), m, sptr)
            stream.synchronize()
            parity = bool(np.array_equal(y[:m].cpu().numpy(), o.pull(x[:m].cpu().numpy(), m)[0]))
            chk.close()

        def step():
            obj.pull_dev(x.data_ptr(), obj.source_count(n), y.data_ptr(), n, sptr)
        S_ = len(obj.path())
        bytes_per, kernel = 4.0 + 4.0 / (1 << S_), "upchan_stage_kernel x%d" % S_
        instr = sum(0.5 ** s for s in range(S_)) * (24 * 2 * 2 / 2.0 + 6.0)      # per output sample: half the calls run the 24-tap FIR on 2 components
    elif wl["kind"] == "fftfilt":
        x = torch.randn((n, 2), dtype=torch.float32, device=c.dev, generator=g) * 8000
        y = torch.empty((n, 2), dtype=torch.float32, device=c.dev)
        obj = S.FftFilt(0, 300 / 48000.0, 3000 / 48000.0, wl["flen"])
        if c.rank == 0 and want_parity:
            m = 1 << 16
            chk = S.FftFilt(0, 300 / 48000.0, 3000 / 48000.0, wl["flen"])
            got_n = chk.run_dev(1, x.data_ptr(), m, y.data_ptr(), n, usb=True, get_dc=False, stream=sptr)
            stream.synchronize()
            want = portbind.PortFftFilt(0, 300 / 48000.0, 3000 / 48000.0, wl["flen"]).run(1, x[:m].cpu().numpy().view(np.complex64).ravel(), True, False)
            got = y[:got_n].cpu().numpy().view(np.complex64).ravel()
            parity = bool(got.shape == want.shape and np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want)))
            chk.close()

        def step():
            obj.run_dev(1, x.data_ptr(), n, y.data_ptr(), n, usb=True, get_dc=False, stream=sptr)
        bytes_per, kernel, instr = 16.0, "fftfilt_kernel", None
    else:
        nc = wl["channels"]
        per = n // nc
        x = torch.randn((nc, per, 2), dtype=torch.float32, device=c.dev, generator=g) * 8000
        cnt = torch.full((nc,), per, dtype=torch.int64, device=c.dev)
        y = torch.empty((3, nc, per), dtype=torch.float32, device=c.dev)
        obj = S.Demod(S.Demod.FM_DELTA, 0.25, n_channels=nc)
        if c.rank == 0 and want_parity:
            chk = S.Demod(S.Demod.FM_DELTA, 0.25, n_channels=nc)
            chk.run_pool_dev(x.data_ptr(), per, cnt.data_ptr(), y[0].data_ptr(), per, y[1].data_ptr(), y[2].data_ptr(), sptr)
            stream.synchronize()
            ch = nc // 3
            want = portbind.PortDemod(1, 0.25).run(x[ch].cpu().numpy().view(np.complex64).ravel())
            parity = all(bool(np.array_equal(y[j, ch].cpu().numpy(), want[j])) for j in range(3))
            chk.close()

        def step():
            obj.run_pool_dev(x.data_ptr(), per, cnt.data_ptr(), y[0].data_ptr(), per, y[1].data_ptr(), y[2].data_ptr(), sptr)
        bytes_per, kernel, instr = 8.0 + 12.0, "demod_kernel", None
    total_ms, kern_ms, clocks = timed_steps(c, stream, step, steps, warmup)
    k_ms = float(np.mean(kern_ms))
    achieved = n * bytes_per / (k_ms * 1e-3) / 1e9
    clk = (clocks or {}).get("sm_mhz") or 1965.0
    issue = None
    if instr:
        roof = c.sm_count * 128 * clk * 1e6 / instr / 1e6
        issue = {"instr_per_sample": instr, "roof_MSps_at_sampled_clk": roof, "frac": (n / (k_ms * 1e-3) / 1e6) / roof}
    else:
        issue = {"instr_per_sample": None, "roof_MSps_at_sampled_clk": None, "frac": None,
                 "note": "HBM view only: %g algorithmic bytes per sample" % bytes_per}
    res = {"value": c.world * n * steps / (total_ms * 1e-3) / 1e6, "ms_per_step": total_ms / steps, "clocks": clocks, "parity": parity, "launches": steps,
           "config": {"workload": wl_name, "desc": wl["desc"], "samples_per_step": n, "l2": "outputs > 126 MB L2 per step", "parallelism": "replicas x%d" % c.world},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": c.hbm_peak, "unit": "GB/s", "frac": achieved / c.hbm_peak, "traffic": None,
                        "peak_source": c.peak_src, "kernel": kernel, "kernel_ms": k_ms, "algorithmic_bytes_per_sample": bytes_per, "issue": issue},
           "dtype": "f32" if wl["kind"] in ("demod", "fftfilt") else "s32", "scaling": "weak"}
    obj.close()
    return res


def _bank_parity(c, bank, info, mine, fs, cutoff, d_ptr, m, sptr, stream):
    """Oracle check (checker only) of THIS rank's bank on the first m samples of the device buffer at d_ptr -- at N > 1 the
    block this rank actually received: tree outputs bit for bit, front-end within 1e-5 relative RMS, for the first, middle
    and last of the rank's channels (downchannelizer.cpp:50-91, nfmdemod.cpp:150-155,315)."""
    torch, capi = c.torch, c.capi
    _oracle_mod()
    from oracle import portbind
    xs = _tensor_from_ptr(c, d_ptr, 2 * m).cpu().numpy().reshape(-1, 2)
    bank.reset(sptr)
    bank.feed_dev(d_ptr, m, sptr)
    stream.synchronize()
    ok = True
    for k in sorted({0, len(info) // 2, len(info) - 1}):
        cid, rate, ofs, path = info[k]
        o = portbind.PortDownChannelizer()
        o.configure(fs, 48000, mine[k])
        ch = o.feed(xs)
        ok = ok and np.array_equal(bank.fetch(cid), ch)
        fe = portbind.PortFrontEnd(-ofs, rate, 48000, cutoff).feed(ch)
        got = bank.fetch(cid, capi.STAGE_FRONTEND)
        ok = ok and got.shape == fe.shape and float(np.sqrt(np.mean((got - fe) ** 2)) / np.sqrt(np.mean(fe ** 2))) <= 1e-5
    return bool(ok)


def _tensor_from_ptr(c, ptr, n_int16):
    """A torch int16 view of n_int16 scalars of device memory owned by the library (no copy)."""
    class _Arr:
        pass
    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (int(n_int16),), "typestr": "<i2", "data": (int(ptr), False), "version": 3}
    return c.torch.as_tensor(a, device=c.dev)


def bench_bank(c, args, wl_name, wl, steps, warmup, want_e2e=True, want_parity=True):
    """The channel bank (configs 3 and 5).  N = 1: one bank, device-resident baseband.  N > 1: channels sharded by frequency
    block (b200dsp_dist_shard); every step the library object b200dsp_dist NCCL-broadcasts the block from rank 0's device
    buffer into a receive slot (the transfer of step k+1 under the kernels of step k) and feeds this rank's bank from it."""
    import sdrangel_b200 as S
    torch, capi, dist = c.torch, c.capi, c.dist
    fs, fcs = wl["plan"]()
    n = args.samples or wl["n"]
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    sb = None
    if c.world > 1:
        ids = [S.ShardedBank.unique_id() if c.rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sb = S.ShardedBank(fs, fcs, 48000, c.rank, c.world, ids[0], frontend=(cutoff, 48000), chunk=n)
        bank, info, lo_ch, hi_ch = sb.bank, sb.info, sb.lo, sb.hi
        sb.reserve(n)
        # transports: the copy-engine chain (IPC slots, no SMs) when the driver allows it, NCCL broadcast otherwise
        p2p_ok = not os.environ.get("B200_BENCH_NO_P2P")
        blob = None
        if p2p_ok:
            try:
                blob = sb.p2p_export(n)
            except Exception:
                blob = None
        blobs = [None] * c.world
        dist.all_gather_object(blobs, blob)
        p2p_ok = all(b is not None for b in blobs)
        if p2p_ok:
            sb.p2p_import(blobs)
    else:
        bank = S.DownChannelizerBank(fs)
        bank.set_chunk(n)                     # one pass over the tree per step: launch latencies amortised over the whole batch
        info = []
        for fc in fcs:
            cid, rate, ofs, path = bank.add_channel(48000, fc)
            bank.set_frontend(cid, -ofs, cutoff, 48000)
            info.append((cid, rate, ofs, path))
        lo_ch, hi_ch = 0, len(fcs)
    mine = fcs[lo_ch:hi_ch]
    nodes = bank.node_count()
    stage_inputs = sum(2.0 ** -(k - 1) for k in _node_depths([p for _, _, _, p in info]))
    g = torch.Generator(device=c.dev)
    g.manual_seed(1)
    x = torch.randint(-2048, 2048, (2 * n,), dtype=torch.int16, device=c.dev, generator=g) if c.rank == 0 else None
    stream = torch.cuda.Stream(device=c.dev)
    sptr = stream.cuda_stream
    torch.cuda.synchronize()
    barrier(c)
    state = {"i": 0, "overlap": True, "begun": set(), "mode": "nccl", "pbegun": 0}

    def step():
        i = state["i"]
        if sb is None:
            bank.feed_dev(x.data_ptr(), n, sptr)
        elif state["mode"] == "p2p":
            # blocks i+1 and i+2 are already on their way down the chain while block i is computed (three slots)
            xp = x.data_ptr() if c.rank == 0 else 0
            while state["pbegun"] <= i + 2:
                sb.p2p_begin(state["pbegun"] % 3, xp, n, None)
                state["pbegun"] += 1
            sb.p2p_feed(i % 3, sptr)
        else:
            cur = i & 1
            xp = x.data_ptr() if c.rank == 0 else 0
            if cur not in state["begun"]:
                sb.bcast_begin(cur, xp, n, 0, None)
                state["begun"].add(cur)
            if state["overlap"]:
                # the next step's block starts moving now (collective stream), under this step's kernels; its slot was last
                # read by step i-1, which the begin waits for
                sb.bcast_begin(cur ^ 1, xp, n, 0, None)
                state["begun"].add(cur ^ 1)
            sb.feed(cur, sptr)
            state["begun"].discard(cur)
        state["i"] = i + 1

    bcast_mode, bcast_trials = None, None
    if sb is not None:
        # overlapped or in stream order: whichever is faster at this N (all ranks agree through a MAX all-reduce)
        trial = {}
        for mode in (False, True):
            state["overlap"] = mode
            with torch.cuda.stream(stream):
                for _ in range(3):
                    step()
                barrier(c)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(8):
                    step()
                e1.record(stream)
                barrier(c)
                trial[mode] = max_over_ranks(c, e0.elapsed_time(e1) / 8)
        state["overlap"] = trial[True] <= trial[False]
        bcast_mode = "NCCL broadcast per step, " + ("overlapped with the previous step's kernels" if state["overlap"] else "in stream order before the step's kernels")
        bcast_trials = {"nccl_serial": round(trial[False], 4), "nccl_overlap": round(trial[True], 4)}
        if p2p_ok:
            state["mode"] = "p2p"
            state["i"] = 0
            with torch.cuda.stream(stream):
                for _ in range(3):
                    step()
                barrier(c)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(8):
                    step()
                e1.record(stream)
                barrier(c)
                t_p2p = max_over_ranks(c, e0.elapsed_time(e1) / 8)
            bcast_trials["p2p_chain"] = round(t_p2p, 4)
            if t_p2p <= min(trial.values()):
                bcast_mode = "copy-engine chain rank 0 -> ... -> N-1 (b200dsp_dist_p2p_*: IPC slots, stream-ordered counters), two blocks ahead of the kernels"
            else:
                state["mode"] = "nccl"

    total_ms, kern_ms, clocks = timed_steps(c, stream, step, steps, warmup)
    value = n * steps / (total_ms * 1e-3) / 1e6
    step_ms = total_ms / steps
    tree_ms, tree_launches = bank.tree_time()         # the tree launches of the last timed step, between events on its stream

    parity = None
    if want_parity:
        # every rank checks ITS channels on the block IT holds (N > 1: the slot the last broadcast landed in), and that this
        # block is rank 0's, byte for byte (a checksum all-reduced as min and max)
        m = min(n, 1 << 19)
        if sb is None:
            dptr = x.data_ptr()
        else:
            sb.sync()
            stream.synchronize()
            dptr = sb.p2p_slot((state["i"] - 1) % 3)[0] if state["mode"] == "p2p" else sb.slot((state["i"] - 1) & 1)[0]
        ok = _bank_parity(c, bank, info, mine, fs, cutoff, dptr, m, sptr, stream)
        if sb is not None:
            cs = _tensor_from_ptr(c, dptr, 2 * n).to(torch.int64).sum().to(torch.float64)
            lo_, hi_ = cs.clone(), cs.clone()
            dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
            ok = ok and float(lo_.item()) == float(hi_.item())
            ok = max_over_ranks(c, 0.0 if ok else 1.0) == 0.0
        parity = bool(ok)

    paths = [p for _, _, _, p in info]
    out_bytes = sum(48000.0 / fs * 8 for _ in fcs)                  # whole job: every channel's 48 kS/s complex64 output
    tree_instr = _tree_instr_per_sample(paths)
    fe_instr = len(mine) * (48000.0 / fs * 144 + (fs / 2.0 ** len(paths[0]) if paths else 0) / fs * 12)
    instr_per_sample = tree_instr + fe_instr
    f_clk = (clocks.get("sm_mhz") or 1965) * 1e6
    issue_peak = c.sm_count * 128 * f_clk                            # thread-instructions per second per GPU
    issue_roof = issue_peak / instr_per_sample / 1e6                 # input MS/s per GPU for this rank's share of the work
    rank_rate = n / (step_ms * 1e-3) / 1e6
    alg_bytes = 4 + out_bytes                                        # 7.2 B per input sample at 1024 channels
    traffic = measured_traffic(wl_name, n) if c.world == 1 else None
    roofline = {
        "bound": "issue", "kernel": "hb48_fused_kernel (%d launches per step) + frontend54_kernel" % tree_launches,
        "achieved": rank_rate * instr_per_sample / 1e3, "peak": issue_peak / 1e9, "unit": "G thread-instr/s", "frac": rank_rate / issue_roof,
        "instr_per_sample": instr_per_sample,
        "note": "binding roof (SURVEY.md 8d): algorithmic instructions of this rank's step -- per parent sample 27 per child of the shared-prefix tree, "
                "31 for a lower/upper pair (one shared tap sum); 144 FFMA per front-end output + 12 per mixed channel sample -- over 128 lanes x SMs x "
                "sampled clock",
        "tree_share_of_step": tree_ms / step_ms, "tree_ms": tree_ms,
        "traffic": traffic,
        "traffic_note": "dram__bytes_read+write of every kernel of one step (ncu, profiles/r02_traffic.json), scaled to this step's samples",
        "hbm": {"algorithmic_bytes_per_sample": alg_bytes, "achieved": alg_bytes * n / (step_ms * 1e-3) / 1e9, "peak": c.hbm_peak, "unit": "GB/s",
                "frac": alg_bytes * n / (step_ms * 1e-3) / 1e9 / c.hbm_peak, "peak_source": c.peak_src,
                "traffic_over_algorithmic": (traffic / (alg_bytes * n)) if traffic else None},
        "issue": {"instr_per_sample": instr_per_sample, "roof_MSps_at_sampled_clk": issue_roof, "frac": rank_rate / issue_roof},
    }
    if c.world > 1:
        roofline["nvlink"] = {"bytes_per_sample_per_gpu": 4, "peak_GBps": 770.0, "roof_MSps": 770e9 / 4 / 1e6, "frac": value / (770e9 / 4 / 1e6),
                              "note": "every GPU ingests the whole baseband: 4 B per input sample against the measured 770 GB/s peer bandwidth (B200_PROFILING.md)"}
    res = {"value": value, "ms_per_step": step_ms, "clocks": clocks, "parity": parity,
           "launches": steps * (tree_launches + 2),
           "config": {"workload": wl_name, "desc": wl["desc"], "samples_per_step": n, "input_rate": fs, "channels": len(fcs),
                      "channels_this_rank": len(mine), "tree_nodes_this_rank": nodes, "stage_inputs_per_sample_this_rank": stage_inputs,
                      "l2": "input %.0f MiB per step > 126 MB L2, streamed from HBM every step" % (n * 4 / 2 ** 20),
                      "parallelism": "channels sharded x%d (contiguous frequency blocks), baseband %s" % (
                          c.world, ("b200dsp_dist (library object): " + str(bcast_mode)) if c.world > 1 else "local"),
                      "broadcast_trials_ms_per_step": bcast_trials},
           "roofline": roofline, "dtype": "s32", "scaling": "strong"}
    if want_e2e:
        with _NearGpu(c) as near:
            res["e2e"] = _bank_e2e(c, args, fs, fcs, cutoff, n, x, sb)
            res["e2e"]["host_cpus"] = ("GPU-local NUMA node: %d CPUs" % len(near.cpus)) if near.cpus else "no affinity change"
    if sb is not None:
        sb.close()
    else:
        bank.close()
    return res


class _NearGpu:
    """Host side of an end-to-end measurement on the CPUs of the GPU's own NUMA node (what a device plugin's thread would be
    pinned to): pinned staging buffers are first touched there and the PCIe copies do not cross the socket interconnect.
    B200_BENCH_NO_AFFINITY=1 disables it.  The CPU-baseline legs run outside, on every core."""

    def __init__(self, c):
        self.c, self.old, self.cpus = c, None, None

    def __enter__(self):
        if os.environ.get("B200_BENCH_NO_AFFINITY"):
            return self
        try:
            pr = self.c.torch.cuda.get_device_properties(self.c.local)
            bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            txt = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
            cpus = set()
            for part in txt.split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cpus.update(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
            old = os.sched_getaffinity(0)
            cpus &= old
            if cpus and cpus != old:
                os.sched_setaffinity(0, cpus)
                self.old, self.cpus = old, cpus
        except Exception:
            pass
        return self

    def __exit__(self, *a):
        if self.old is not None:
            try:
                os.sched_setaffinity(0, self.old)
            except Exception:
                pass
        return False


def _bank_e2e(c, args, fs, fcs, cutoff, n, x, sb):
    """End to end through the plugin-facing calls, host buffers both sides, K sub-blocks per step (16, or 32 for steps of 100 M samples and more).
    N = 1: b200dsp_bank_process (one call per step: pinned host baseband in, every channel's 48 kS/s complex64 output to pinned
    host memory; H2D / kernels / D2H of finished columns overlapped inside the library).
    N > 1: b200dsp_dist_ingest_begin per sub-block -- every rank copies ITS 1/N time slice over its own PCIe link, an in-place
    NCCL all-gather completes the block on every GPU -- then b200dsp_dist_feed and the pooled D2H of this rank's channels."""
    import sdrangel_b200 as S
    torch, capi, dist = c.torch, c.capi, c.dist
    # passes of ~3.1 M samples (12.6 MB of H2D each: measured -- half that size is launch-bound, 9.95e3 against 1.13e4 MS/s)
    K = int(os.environ.get("B200_BENCH_E2E_PASSES", "32" if n >= (3 << 25) else "16"))
    nb_ = n // K
    ksteps = 10 if sb is None else 5           # a step is ~5 ms of wall clock: enough of them to average out host scheduling noise
    if sb is None:
        bank = S.DownChannelizerBank(fs)
        bank.set_chunk(nb_)
        for fc in fcs:
            cid, rate, ofs, path = bank.add_channel(48000, fc)
            bank.set_frontend(cid, -ofs, cutoff, 48000)
        hx = torch.empty((2 * n,), dtype=torch.int16, pin_memory=True)
        hx.copy_(x)
        stride = int(n * 48000.0 / fs) + 64
        hout = torch.empty((len(fcs), stride, 2), dtype=torch.float32, pin_memory=True)
        torch.cuda.synchronize()
        cnt = bank.process(hx.data_ptr(), n, capi.STAGE_FRONTEND, hout.data_ptr(), stride)       # warm-up (allocations)
        t0 = time.perf_counter()
        for _ in range(ksteps):
            cnt = bank.process(hx.data_ptr(), n, capi.STAGE_FRONTEND, hout.data_ptr(), stride)
        dt = time.perf_counter() - t0
        d2h = int(cnt.sum()) * 8
        bank.close()
        return {"value": n * ksteps / dt / 1e6, "unit": "input MS/s", "h2d_bytes_per_step": int(n * 4), "d2h_bytes_per_step": d2h, "steps": ksteps,
                "api": "b200dsp_bank_process: pinned host baseband -> every channel's front-end output in pinned host memory, %d passes per call "
                       "(H2D, kernels and the strided D2H of finished output columns overlapped on three streams inside the library)" % K,
                "samples_per_step": n}
    # N > 1
    bank, info = sb.bank, sb.info
    # the same synthetic block on every rank's host (set-up, untimed): rank 0's buffer through one broadcast
    xd = x if c.rank == 0 else torch.empty((2 * n,), dtype=torch.int16, device=c.dev)
    dist.broadcast(xd.view(torch.int32), src=0)
    hx = torch.empty((2 * n,), dtype=torch.int16, pin_memory=True)
    hx.copy_(xd)
    del xd
    cnt_slice = nb_ // c.world
    nch = max(len(info), 1)
    stride = int(nb_ * 48000.0 / fs) + 64
    pools = [torch.empty((nch, stride, 2), dtype=torch.float32, device=c.dev) for _ in range(2)]
    dcnt = [torch.zeros((nch,), dtype=torch.int64, device=c.dev) for _ in range(2)]
    hout = torch.empty((K, nch, stride, 2), dtype=torch.float32, pin_memory=True)
    hcnt = torch.zeros((K, nch), dtype=torch.int64, pin_memory=True)
    s_out = torch.cuda.Stream(device=c.dev)
    stream = torch.cuda.Stream(device=c.dev)
    sptr = stream.cuda_stream
    torch.cuda.synchronize()

    def slice_ptr(k):
        return hx.data_ptr() + 4 * (k * nb_ + c.rank * cnt_slice)

    def e2e_step():
        ev_out = [None] * K
        sb.ingest_begin(0, slice_ptr(0), nb_)
        for k in range(K):
            if k + 1 < K:
                sb.ingest_begin((k + 1) & 1, slice_ptr(k + 1), nb_)        # the next sub-block moves under this one's kernels
            sb.feed(k & 1, sptr)
            if k >= 2:
                stream.wait_event(ev_out[k - 2])           # the pool being gathered into has left for the host
            bank.gather_dev(capi.STAGE_FRONTEND, pools[k % 2].data_ptr(), stride, dcnt[k % 2].data_ptr(), sptr)
            e_g = torch.cuda.Event()
            e_g.record(stream)
            with torch.cuda.stream(s_out):
                s_out.wait_event(e_g)
                hout[k].copy_(pools[k % 2], non_blocking=True)
                hcnt[k].copy_(dcnt[k % 2], non_blocking=True)
                ev_out[k] = torch.cuda.Event()
                ev_out[k].record(s_out)
        s_out.synchronize()
        stream.synchronize()

    with torch.cuda.stream(stream):
        e2e_step()
        barrier(c)
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks(c, time.perf_counter() - t0)
        tot = torch.tensor([float(hcnt.sum()) * 8], device=c.dev, dtype=torch.float64)
        dist.all_reduce(tot)
    return {"value": n * ksteps / dt / 1e6, "unit": "input MS/s", "h2d_bytes_per_step": int(n * 4), "d2h_bytes_per_step": int(tot.item()), "steps": ksteps,
            "api": "b200dsp_dist_ingest_begin (every rank H2D-copies its 1/%d time slice from pinned host memory, in-place NCCL all-gather) -> "
                   "b200dsp_dist_feed -> b200dsp_bank_gather_dev + D2H of this rank's channels (pinned host), %d sub-blocks per step" % (c.world, K),
            "samples_per_step": n}


def bench_spectrum(c, args, wl_name, wl, steps, warmup, want_e2e=True, want_parity=True):
    import sdrangel_b200 as S
    torch, capi = c.torch, c.capi
    n = args.samples or wl["n"]
    fft = wl["fft"]
    sp = S.SpectrumVis()
    sp.configure(fft, 0, wl["avg_nb"], wl["avg_mode"], 1, False)
    g = torch.Generator(device=c.dev)
    g.manual_seed(2 + c.rank)
    x = torch.randint(-2048, 2048, (2 * n,), dtype=torch.int16, device=c.dev, generator=g)
    cap = n // fft // max(1, wl["avg_nb"]) + 2
    y = torch.empty((cap, fft), dtype=torch.float32, device=c.dev)
    stream = torch.cuda.Stream(device=c.dev)
    sptr = stream.cuda_stream
    parity = None
    if c.rank == 0 and want_parity:       # oracle = checker only: first 40 frames
        _oracle_mod()
        from oracle import portbind
        m = fft * 40
        chk = S.SpectrumVis()
        chk.configure(fft, 0, wl["avg_nb"], wl["avg_mode"], 1, False)
        nf = chk.feed_dev(x.data_ptr(), m, y.data_ptr(), cap, False, sptr)
        stream.synchronize()
        o = portbind.PortSpectrumVis()
        o.configure(fft, 0, wl["avg_nb"], wl["avg_mode"], 1, False)
        want = o.feed(x[: 2 * m].cpu().numpy().reshape(-1, 2))
        got = y[:nf].cpu().numpy()
        ok = np.isfinite(want)
        parity = bool(got.shape == want.shape and np.max(np.abs(got[ok] - want[ok])) <= 1e-2 and
                      float(np.sqrt(np.mean((10 ** (got[ok] / 10) - 10 ** (want[ok] / 10)) ** 2)) / np.sqrt(np.mean((10 ** (want[ok] / 10)) ** 2))) <= 1e-5)
        chk.close()

    def step():
        sp.feed_dev(x.data_ptr(), n, y.data_ptr(), cap, False, sptr)

    total_ms, kern_ms, clocks = timed_steps(c, stream, step, steps, warmup)
    value = c.world * n * steps / (total_ms * 1e-3) / 1e6
    step_ms = float(np.mean(kern_ms))
    bps = 4 + 4.0 / max(1, wl["avg_nb"])
    achieved = n * bps / (step_ms * 1e-3) / 1e9
    instr_per_sample = 68.0
    f_clk = (clocks.get("sm_mhz") or 1965) * 1e6
    issue_roof = c.sm_count * 128 * f_clk / instr_per_sample / 1e6
    res = {"value": value, "ms_per_step": total_ms / steps, "clocks": clocks, "parity": parity, "launches": steps,
           "config": {"workload": wl_name, "desc": wl["desc"], "samples_per_step": n, "fft_size": fft, "frames_per_step": n // fft,
                      "average_nb": wl["avg_nb"], "averaging": "fixed", "window": "BlackmanHarris",
                      "l2": "input %.0f MiB per step > 126 MB L2, streamed from HBM every step" % (n * 4 / 2 ** 20),
                      "parallelism": "replicas x%d (one stream per GPU)" % c.world},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": c.hbm_peak, "unit": "GB/s", "frac": achieved / c.hbm_peak,
                        "traffic": None, "peak_source": c.peak_src, "kernel": "spectrum_kernel", "kernel_ms": step_ms,
                        "algorithmic_bytes_per_sample": bps,
                        "issue": {"instr_per_sample": instr_per_sample, "roof_MSps_at_sampled_clk": issue_roof,
                                  "frac": (n / (step_ms * 1e-3) / 1e6) / issue_roof,
                                  "note": "5 N log2 N flops per frame = 60 flop/sample + window/scale/power/log ~ 8 (SURVEY.md 8d)"}},
           "dtype": "f32", "scaling": "weak"}
    if want_e2e:
        n_e = min(n, 1 << 24)
        hx = torch.empty((2 * n_e,), dtype=torch.int16, pin_memory=True)
        hx.copy_(x[: 2 * n_e])
        hy = torch.empty((n_e // fft // max(1, wl["avg_nb"]) + 2, fft), dtype=torch.float32, pin_memory=True)
        L_ = capi.lib()
        s2 = S.SpectrumVis()
        s2.configure(fft, 0, wl["avg_nb"], wl["avg_mode"], 1, False)
        nn = C.c_int64(0)

        def e2e_step():
            capi.check(L_.b200dsp_spectrum_feed(s2._h, hx.data_ptr(), n_e, 0, hy.data_ptr(), hy.shape[0], C.byref(nn)))

        e2e_step()
        barrier(c)
        t0 = time.perf_counter()
        ksteps = 5
        for _ in range(ksteps):
            e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks(c, time.perf_counter() - t0)
        res["e2e"] = {"value": c.world * n_e * ksteps / dt / 1e6, "unit": "input MS/s", "h2d_bytes_per_step": int(n_e * 4),
                      "d2h_bytes_per_step": int(nn.value * fft * 4), "steps": ksteps,
                      "api": "b200dsp_spectrum_feed (host pointers, pinned)", "samples_per_step": n_e}
        s2.close()
    sp.close()
    return res


def bench_bank_coop(c, args, wl_name, wl, steps, warmup, want_e2e=True, want_parity=True):
    """N > 1: the cooperative bank (sdrangel_b200/coop.py): scatter of time slices -> top tree levels on the slice ->
    all-to-all of the depth-k node streams -> per-rank banks.  Every collective is NCCL over NVLink, in stream order."""
    import sdrangel_b200 as S
    from sdrangel_b200.coop import CoopPlan, CoopRank, HALO
    torch, capi, dist = c.torch, c.capi, c.dist
    fs, fcs = wl["plan"]()
    n = args.samples or wl["n"]
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    stream = torch.cuda.Stream(device=c.dev)
    sptr = stream.cuda_stream

    class Pipe:
        def __init__(self, n_samples, x=None):
            self.plan = p = CoopPlan(fs, fcs, 48000, c.world, n_samples)
            self.rk = CoopRank(p, c.rank, frontend=(cutoff, 48000))
            self.slice = torch.empty(HALO + p.m, dtype=torch.int32, device=c.dev)
            self.scatter_list = None
            if c.rank == 0:
                xi = x.view(torch.int32)
                self.xcat = torch.cat([xi[-HALO:], xi])         # the stream is the buffer repeated: slice 0's halo is its end
                self.scatter_list = [self.xcat[r * p.m: r * p.m + HALO + p.m] for r in range(c.world)]
            self.my_nodes = p.rank_nodes[c.rank]
            self.in_splits = [len(p.rank_nodes[q]) * p.mk for q in range(c.world)]
            self.out_splits = [len(self.my_nodes) * p.mk] * c.world
            self.send = torch.empty(sum(self.in_splits), dtype=torch.int32, device=c.dev)
            self.recv = torch.empty(sum(self.out_splits), dtype=torch.int32, device=c.dev)
            self.streams = torch.empty((len(self.my_nodes), c.world, p.mk), dtype=torch.int32, device=c.dev)
            torch.cuda.synchronize()                            # tensors above were built on the default stream

        def step(self):
            p, rk = self.plan, self.rk
            dist.scatter(self.slice, self.scatter_list, src=0)
            rk.top.reset(sptr)
            rk.top.feed_dev(self.slice.data_ptr(), HALO + p.m, sptr)
            off = 0
            for q in range(c.world):
                for v in p.rank_nodes[q]:
                    rk.top.copy_out_dev(rk.top_ids[v], p.skip, p.mk, self.send.data_ptr() + 4 * off, sptr)
                    off += p.mk
            dist.all_to_all_single(self.recv, self.send, self.out_splits, self.in_splits)
            self.streams.copy_(self.recv.view(c.world, len(self.my_nodes), p.mk).permute(1, 0, 2))
            for j, v in enumerate(self.my_nodes):
                rk.subs[v].feed_dev(self.streams[j].data_ptr(), p.n >> p.k, sptr)

    g = torch.Generator(device=c.dev)
    g.manual_seed(1)
    x = torch.randint(-2048, 2048, (2 * n,), dtype=torch.int16, device=c.dev, generator=g) if c.rank == 0 else None

    parity = None
    if want_parity:       # small instance, second step vs oracle chains fed the buffer twice (the stream is the buffer repeated)
        ns = c.world * 768 * 8 * 16
        xs = torch.randint(-2048, 2048, (2 * ns,), dtype=torch.int16, device=c.dev, generator=g) if c.rank == 0 else None
        pp = Pipe(ns, xs)
        with torch.cuda.stream(stream):
            if c.rank == 0:
                pp.xcat[:HALO].zero_()                          # first step: nothing precedes the stream (the oracle's zero history)
            pp.step()
            if c.rank == 0:
                pp.xcat[:HALO].copy_(pp.xcat[-HALO:])
            pp.step()
        stream.synchronize()
        if c.rank == 0:
            _oracle_mod()
            from oracle import portbind
            hx = xs.cpu().numpy().reshape(-1, 2)
            ok = True
            lo, hi = pp.plan.ranges[0]
            for i in (lo, (lo + hi) // 2, hi - 1):
                rate, ofs, path = pp.plan.chains[i]
                o = portbind.PortDownChannelizer()
                o.configure(fs, 48000, fcs[i])
                fe = portbind.PortFrontEnd(-ofs, rate, 48000, cutoff)
                fe.feed(o.feed(hx))
                ch = o.feed(hx)
                want = fe.feed(ch)
                node, cid = pp.rk.chan[i]
                gc_ = pp.rk.subs[node].fetch(cid)
                ok_c = gc_.shape == ch.shape and np.array_equal(gc_, ch)
                got = pp.rk.subs[node].fetch(cid, capi.STAGE_FRONTEND)
                ok_f = got.shape == want.shape and float(np.sqrt(np.mean((got - want) ** 2)) / np.sqrt(np.mean(want ** 2))) <= 1e-5
                if not (ok_c and ok_f):
                    bad = int(np.argmax(np.any(gc_ != ch, axis=1))) if gc_.shape == ch.shape else -1
                    print("coop parity mismatch: channel %d node %s tree %s (first bad %d of %s, mk %d) front-end %s" %
                          (i, node, ok_c, bad, gc_.shape, pp.plan.mk, ok_f), file=sys.stderr)
                ok = ok and ok_c and ok_f
            parity = bool(ok)
        pp.rk.close()
    barrier(c)

    pipe = Pipe(n, x)
    total_ms, kern_ms, clocks = timed_steps(c, stream, pipe.step, steps, warmup)
    value = n * steps / (total_ms * 1e-3) / 1e6
    step_ms = float(np.mean(kern_ms))
    p = pipe.plan
    lo, hi = p.ranges[c.rank]
    stage_inputs = p.stage_inputs(c.rank)
    out_bytes = (hi - lo) * 48000.0 / fs * 8
    alg_bytes = n * (4.0 * (p.m + HALO) / n + 4.0 * len(pipe.my_nodes) / (1 << p.k) + out_bytes)
    achieved = alg_bytes / (step_ms * 1e-3) / 1e9
    instr_per_sample = stage_inputs * 27 + (hi - lo) * 48000.0 / fs * 160
    f_clk = (clocks.get("sm_mhz") or 1965) * 1e6
    issue_roof = c.sm_count * 128 * f_clk / instr_per_sample / 1e6
    depth = max(len(ch[2]) for ch in p.chains)
    res = {"value": value, "ms_per_step": total_ms / steps, "clocks": clocks, "parity": parity,
           "launches": steps * (p.k + 1 + len(pipe.my_nodes) * (depth - p.k + 3)),
           "config": {"workload": wl_name, "desc": wl["desc"], "samples_per_step": n, "input_rate": fs, "channels": len(fcs),
                      "channels_this_rank": hi - lo, "stage_inputs_per_sample_this_rank": stage_inputs,
                      "l2": "input %.0f MiB per step > 126 MB L2" % (n * 4 / 2 ** 20),
                      "parallelism": "cooperative x%d: NCCL scatter of time slices (+%d-sample halo), top %d tree levels per slice, NCCL all-to-all of the "
                                     "depth-%d node streams, channels below them sharded by frequency" % (c.world, HALO, p.k, p.k)},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": c.hbm_peak, "unit": "GB/s", "frac": achieved / c.hbm_peak,
                        "traffic": None, "peak_source": c.peak_src, "kernel": "hb48_level_kernel (one launch per tree level)", "kernel_ms": step_ms,
                        "algorithmic_bytes_per_sample": alg_bytes / n,
                        "issue": {"instr_per_sample": instr_per_sample, "roof_MSps_at_sampled_clk": issue_roof,
                                  "frac": (n / (step_ms * 1e-3) / 1e6) / issue_roof,
                                  "note": "per rank: 27 instructions per stage-input sample of its share of the tree + ~160 per front-end output"},
                        "nvlink": {"bytes_in_per_step_this_rank": int(4 * (p.m + HALO) * (c.rank != 0) + 4 * len(pipe.my_nodes) * p.mk * (c.world - 1)),
                                   "root_egress_bytes_per_step": int(4 * (p.m + HALO) * (c.world - 1))}},
           "dtype": "s32", "scaling": "strong"}
    if want_e2e:
        # end to end: rank 0's baseband starts in pinned host memory, every rank reads its channels' outputs back
        hx = torch.empty((2 * n,), dtype=torch.int16, pin_memory=True) if c.rank == 0 else None
        if c.rank == 0:
            hx.copy_(x)
        outs = {i: torch.empty((int(n * 48000 / fs) + 64, 2), dtype=torch.float32, pin_memory=True) for i in range(lo, hi)}
        L_ = capi.lib()
        nn = C.c_int64(0)

        def e2e_step():
            if c.rank == 0:
                x.copy_(hx, non_blocking=True)
                pipe.xcat[HALO:].copy_(x.view(torch.int32))
                pipe.xcat[:HALO].copy_(x.view(torch.int32)[-HALO:])
            pipe.step()
            stream.synchronize()
            for i, o in outs.items():
                node, cid = pipe.rk.chan[i]
                capi.check(L_.b200dsp_bank_fetch(pipe.rk.subs[node]._h, cid, capi.STAGE_FRONTEND, o.data_ptr(), o.shape[0], C.byref(nn)))

        with torch.cuda.stream(stream):
            e2e_step()
            barrier(c)
            t0 = time.perf_counter()
            ksteps = 3
            for _ in range(ksteps):
                e2e_step()
            torch.cuda.synchronize()
        dt = max_over_ranks(c, time.perf_counter() - t0)
        res["e2e"] = {"value": n * ksteps / dt / 1e6, "unit": "input MS/s", "h2d_bytes_per_step": int(n * 4) if c.rank == 0 else 0,
                      "d2h_bytes_per_step": int((hi - lo) * nn.value * 8), "steps": ksteps,
                      "api": "pinned host baseband on rank 0 -> cooperative bank -> b200dsp_bank_fetch of every channel's front-end output on its rank",
                      "samples_per_step": n}
    pipe.rk.close()
    return res


def _tree_instr_per_sample(paths):
    """Issued-instruction model of the shared-prefix tree per baseband sample: a child costs 27
    instructions per sample of its parent's stream (12 pre-subtractions/additions + 13 multiply-adds + shift + pack); a
    parent's lower-half and upper-half children share the tap sum: 31 for the two."""
    nodes = set()
    for p in paths:
        for k in range(1, len(p) + 1):
            nodes.add(p[:k])
    total = 0.0
    for parent in {n[:-1] for n in nodes}:
        kids = {n[-1] for n in nodes if n[:-1] == parent}
        units = (27.0 if "C" in kids else 0.0) + (31.0 if ("L" in kids and "U" in kids) else 27.0 if ("L" in kids or "U" in kids) else 0.0)
        total += units * 2.0 ** -len(parent)
    return total


def _node_depths(paths):
    seen = set()
    for p in paths:
        for k in range(1, len(p) + 1):
            seen.add(p[:k])
    return [len(s) for s in seen]


def run_ours(args, wl_name, wl):
    c = setup()
    fns = {"decim": bench_decim, "bank": bench_bank, "spectrum": bench_spectrum, "iqcorr": bench_iqcorr, "tx": bench_tx}
    main_fn = fns[wl["type"]]
    if wl["type"] == "bank" and c.world > 1 and os.environ.get("B200_BENCH_COOP"):
        main_fn = bench_bank_coop          # developer switch: time-sliced top levels + all-to-all; measured slower
    res = main_fn(c, args, wl_name, wl, args.steps, args.warmup, want_e2e=not args.no_e2e)
    also = {}
    if c.world == 1 and not args.no_also and not args.samples:
        for other in WORKLOADS:
            if other == wl_name:
                continue
            try:
                o = WORKLOADS[other]
                full = o["type"] == "decim"                      # the 1-GPU decimation targets: complete figures
                r = fns[o["type"]](c, args, other, o, 10 if full else 5, 3, want_e2e=full and not args.no_e2e)
                also[other] = {"value": r["value"], "unit": "input MS/s", "ms_per_step": r["ms_per_step"],
                               "hbm_frac": r["roofline"].get("hbm", {}).get("frac", r["roofline"]["frac"]),
                               "issue_frac": r["roofline"]["issue"]["frac"], "parity_checked_vs_oracle": r["parity"],
                               "samples_per_step": r["config"]["samples_per_step"]}
                if full:
                    also[other]["roofline"] = r["roofline"]
                    also[other]["e2e"] = r.get("e2e")
                    if other == "decimateii" and not args.no_cpu:
                        also[other]["cpu_baseline"] = cpu_reference(o, 5.0)
            except Exception as e:      # an auxiliary measurement must not take the headline down
                also[other] = {"error": repr(e)}
    if c.rank == 0:
        cpu = None if (args.no_cpu or c.world > 1) else cpu_reference(wl, args.cpu_seconds)       # reported at N = 1 only
        line = {"metric": METRIC, "value": res["value"], "unit": "input MS/s", "n_gpus": c.world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": res["scaling"], "vs_baseline": None,
                "dtype": res["dtype"], "data": "synthetic", "config": res["config"], "roofline": res["roofline"], "cpu_baseline": cpu,
                "e2e": res.get("e2e"), "gpu_launches": res["launches"], "clocks": res["clocks"], "parity_checked_vs_oracle": res["parity"],
                "sm_count": c.sm_count, "also": also}
        print(json.dumps(line))
    if c.world > 1:
        c.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="IQ samples per step (default: workload's, > L2)")
    ap.add_argument("--e2e-samples", type=int, default=1 << 26)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # One flagship workload at every N, so the driver's 1 -> 8 GPU series is one strong-scaling curve: the 1024-channel bank
    # (north_star: "at 1/2/4/8 GPUs for the channel bank"; it fits one GPU).  north_star's 1-GPU decimation figures
    # (sdrbench decimateii / decimatefi) ride along in the N = 1 line under "also", with their own roofline, end-to-end and
    # CPU-reference numbers; `--workload decimateii` makes them the line itself.
    wl_name = args.workload or "bank1024"
    wl = WORKLOADS[wl_name]
    if args.impl == "reference":
        run_reference(args, wl_name, wl)
    else:
        run_ours(args, wl_name, wl)


if __name__ == "__main__":
    main()
