#!/usr/bin/env python
"""bench.py — input MS/s of the SDRangel baseband-to-channel hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

Workloads (BASELINE.json configs):
  decimateii   sdrbench decimateii: int16 IQ, log2 decimation 4, centred, 12 bit  (config 1; N=1 default)
  decimatefi   sdrbench decimatefi: float IQ, log2 decimation 6, centred          (config 2)
A "step" is one decimate call over one batch of synthetic IQ that is larger than L2 (so every step streams from HBM).
`value` is device-resident throughput (inputs already in HBM), `e2e` goes through the host-pointer C-ABI call
(b200dsp_decim_run == Decimators::decimate16_cen on a host buffer) with H2D/D2H inside the timed region.
N > 1: the single-stream decimators do not shard ("replicas only"): every rank runs an independent
replica on its own GPU (weak scaling), timed as max over ranks.
--impl reference times the reference's own CPU code (oracle/_ref, compiled from the reference sources) on the host
cores, on the same metric.
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

WORKLOADS = {
    # name: (kind, log2, mode, bytes per input sample in, bytes out per input sample, default samples per step)
    "decimateii": dict(kind="ii", log2=4, mode=2, in_dtype="int16", in_bytes=4, out_bytes=4 / 16, n=1 << 28,
                       desc="sdrbench decimateii: Decimators<qint32,qint16,16,12>::decimate16_cen, synthetic int16 IQ"),
    "decimatefi": dict(kind="fi", log2=6, mode=2, in_dtype="float32", in_bytes=8, out_bytes=4 / 64, n=1 << 27,
                       desc="sdrbench decimatefi: DecimatorsFI::decimate64_cen, synthetic float IQ"),
}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self._stop, self._t = [], set(), threading.Event(), None
        self.max_mhz = None
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
            self._stop.wait(0.02)

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
def cpu_reference_throughput(wl, seconds=12.0, threads=None):
    """Times the reference's own C++ (oracle/_ref/libsdrref.so) if present, else the oracle port, the way sdrbench
    does (mainbench.cpp:83-104): a 2^20-sample buffer fed repeatedly with the filter state carried; one
    independent decimator object per host thread."""
    from oracle import refbind, portbind
    kind = "reference" if refbind.available() else "port"
    n = 1 << 20
    if kind == "reference":
        buf = refbind.sdrbench_s16(n) if wl["kind"][0] == "i" else refbind.sdrbench_f32(n)
        mk = lambda: refbind.RefDecimators(wl["kind"], 12)
    else:
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
        buf = portbind.sdrbench_s16(n) if wl["kind"][0] == "i" else portbind.sdrbench_f32(n)
        mk = lambda: portbind.PortDecimators(wl["kind"], 12)
    threads = threads or (os.cpu_count() or 1)
    objs = [mk() for _ in range(threads)]
    objs[0].run(wl["log2"], wl["mode"], buf)            # warm
    counts = [0] * threads
    t_end = time.perf_counter() + seconds
    per_thread_time = [0.0] * threads

    def work(i):
        t0 = time.perf_counter()
        while time.perf_counter() < t_end:
            objs[i].run(wl["log2"], wl["mode"], buf)
            counts[i] += 1
        per_thread_time[i] = time.perf_counter() - t0

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    wall = time.perf_counter() - t0
    total = sum(counts) * n
    return {"value": total / wall / 1e6, "unit": "input MS/s", "cores": threads, "kind": kind,
            "sample": "%d x 2^20-sample sdrbench buffer per thread, state carried (sdrbench -r), %.1f s wall" % (max(counts), wall),
            "per_core": total / wall / 1e6 / threads}


def run_reference(args, wl_name, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    vals = []
    per = max(2.0, min(10.0, 60.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        r = cpu_reference_throughput(wl, seconds=per)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    line = {"metric": METRIC, "value": v, "unit": "input MS/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s32" if wl["kind"] == "ii" else "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": wl_name, "desc": wl["desc"], "log2_decim": wl["log2"], "fc_pos": "cen"},
            "cpu_baseline": {"value": v, "unit": "input MS/s", "cores": vals[-1]["cores"], "kind": vals[-1]["kind"], "sample": vals[-1]["sample"]},
            "e2e": {"value": v, "unit": "input MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args, wl_name, wl):
    import torch
    import torch.distributed as dist
    import sdrangel_b200 as S
    from sdrangel_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    capi.init(local)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    n = args.samples or wl["n"]
    cls = {"ii": S.Decimators, "fi": S.DecimatorsFI, "ff": S.DecimatorsFF, "if": S.DecimatorsIF}[wl["kind"]]
    dec = cls(12)
    g = torch.Generator(device=dev)
    g.manual_seed(5489 + rank)
    if wl["in_dtype"] == "int16":
        x = torch.randint(-2048, 2048, (2 * n,), dtype=torch.int16, device=dev, generator=g)
    else:
        x = torch.rand((2 * n,), dtype=torch.float32, device=dev, generator=g) * 2 - 1
    n_out = dec.out_count(wl["log2"], wl["mode"], 2 * n)
    out_dt = torch.int16 if wl["kind"][1] == "i" else torch.float32
    y = torch.empty((n_out, 2), dtype=out_dt, device=dev)
    stream = torch.cuda.Stream(device=dev)
    sptr = stream.cuda_stream

    # parity spot check against the oracle on the first 2^20 samples (oracle = checker only)
    parity = None
    if rank == 0:
        from oracle import portbind
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
        m = 1 << 20
        chk = cls(12)
        if wl["kind"] != "ii":
            chk.set_exact_float(True)
        yy = torch.empty((chk.out_count(wl["log2"], wl["mode"], 2 * m), 2), dtype=out_dt, device=dev)
        chk.run_dev(wl["log2"], wl["mode"], x.data_ptr(), 2 * m, yy.data_ptr(), sptr)
        stream.synchronize()
        want = portbind.PortDecimators(wl["kind"], 12).run(wl["log2"], wl["mode"], x[: 2 * m].cpu().numpy())
        parity = bool(np.array_equal(yy.cpu().numpy(), want))
        chk.close()

    def step():
        dec.run_dev(wl["log2"], wl["mode"], x.data_ptr(), 2 * n, y.data_ptr(), sptr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for a, b in evs:
            a.record(stream)
            step()
            b.record(stream)
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
    total_ms = e0.elapsed_time(e1)
    kern_ms = [a.elapsed_time(b) for a, b in evs]
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * n * args.steps / (total_ms * 1e-3) / 1e6
    kern_avg_ms = float(np.mean(kern_ms))
    alg_bytes = n * (wl["in_bytes"] + wl["out_bytes"])
    achieved = alg_bytes / (kern_avg_ms * 1e-3) / 1e9

    # issue-rate view of the same kernel: 2 comps * 34 instr per stage output, sum over stages
    L = wl["log2"]
    instr_per_sample = 2 * 34 * (1 - 2.0 ** -L)
    sm_count = capi.lib().b200dsp_sm_count()
    f_clk = (clocks.get("sm_mhz") or 1965) * 1e6
    issue_roof = sm_count * 128 * f_clk / instr_per_sample / 1e6

    # end-to-end through the host-pointer C-ABI call (pinned host buffers, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
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
        barrier()
        t0 = time.perf_counter()
        ksteps = max(3, min(args.steps, 10))
        for _ in range(ksteps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * n_e * ksteps / dt / 1e6, "unit": "input MS/s", "h2d_bytes_per_step": int(n_e * wl["in_bytes"]),
               "d2h_bytes_per_step": int(n_out_e * (4 if out_dt == torch.int16 else 8)), "steps": ksteps,
               "api": "b200dsp_decim_run (host pointers, pinned; chunked H2D/compute overlap)", "samples_per_step": n_e}
        launches_e2e = ksteps * ((2 * n_e + (8 << 20) - 1) // (8 << 20))
        d2.close()

    if rank == 0:
        cpu = None
        if not args.no_cpu:
            cpu = cpu_reference_throughput(wl, seconds=args.cpu_seconds)
        line = {"metric": METRIC, "value": value, "unit": "input MS/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "s32" if wl["kind"] == "ii" else "f32", "data": "synthetic",
                "config": {"workload": wl_name, "desc": wl["desc"], "samples_per_step": n, "log2_decim": wl["log2"], "fc_pos": "cen",
                           "input_bits": 12, "l2": "input %.0f MiB per step > 126 MB L2, streamed from HBM every step" % (n * wl["in_bytes"] / 2 ** 20),
                           "parallelism": "replicas x%d (single-stream decimator does not shard)" % world},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                             "traffic": None, "peak_source": peak_src, "kernel": "hb64_cascade_kernel", "kernel_ms": kern_avg_ms,
                             "algorithmic_bytes_per_sample": wl["in_bytes"] + wl["out_bytes"],
                             "issue": {"instr_per_sample": instr_per_sample, "roof_MSps_at_sampled_clk": issue_roof,
                                       "frac": (n / (kern_avg_ms * 1e-3) / 1e6) / issue_roof,
                                       "note": "binding roof: 16 IMAD (FMA-heavy pipe) + 16 IADD (ALU pipe) per real output"}},
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": args.steps, "clocks": clocks, "parity_checked_vs_oracle": parity,
                "sm_count": sm_count}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
    args = ap.parse_args()
    wl_name = args.workload or "decimateii"
    wl = WORKLOADS[wl_name]
    if args.impl == "reference":
        run_reference(args, wl_name, wl)
    else:
        run_ours(args, wl_name, wl)


if __name__ == "__main__":
    main()
