#!/usr/bin/env python
"""Developer aid (GPU box): a few steps of one Tx / demod / bank64 workload for ncu (launch lists and --set full captures).
usage: tx_probe.py interps|upchan|demod|bank64 [steps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdrangel_b200 as S
from bench import WORKLOADS, plan64
S.capi.init(0)
which = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
st = torch.cuda.Stream()
sp = st.cuda_stream
if which == "bank64":
    fs, fcs = plan64()
    n = 3 << 22
    x = torch.randint(-2048, 2048, (2 * n,), dtype=torch.int16, device="cuda")
    cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
    b = S.DownChannelizerBank(fs)
    b.set_chunk(n)
    for fc in fcs:
        cid, rate, ofs, path = b.add_channel(48000, fc)
        b.set_frontend(cid, -ofs, cutoff, 48000)
    step = lambda: b.feed_dev(x.data_ptr(), n, sp)                             # noqa: E731
else:
    wl = WORKLOADS[which]
    n = wl["n"]
    if which == "interps":
        L2 = wl["log2"]
        x = torch.randint(-32768, 32768, (n, 2), dtype=torch.int16, device="cuda")
        y = torch.empty((n << L2, 2), dtype=torch.int16, device="cuda")
        o = S.Interpolators(wl["bits"])
        step = lambda: o.run_dev(L2, x.data_ptr(), y.data_ptr(), n * (2 << L2), sp)      # noqa: E731
    elif which == "upchan":
        o = S.UpChannelizer()
        o.configure(*wl["plan"])
        x = torch.randint(-32768, 32768, (o.source_count(n) + 64, 2), dtype=torch.int16, device="cuda")
        y = torch.empty((n, 2), dtype=torch.int16, device="cuda")
        step = lambda: o.pull_dev(x.data_ptr(), o.source_count(n), y.data_ptr(), n, sp)   # noqa: E731
    else:
        nc = wl["channels"]
        per = n // nc
        x = torch.randn((nc, per, 2), dtype=torch.float32, device="cuda") * 8000
        cnt = torch.full((nc,), per, dtype=torch.int64, device="cuda")
        y = torch.empty((3, nc, per), dtype=torch.float32, device="cuda")
        o = S.Demod(1, 0.25, n_channels=nc)
        step = lambda: o.run_pool_dev(x.data_ptr(), per, cnt.data_ptr(), y[0].data_ptr(), per, y[1].data_ptr(), y[2].data_ptr(), sp)   # noqa: E731
with torch.cuda.stream(st):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step()
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(steps):
        step()
    e1.record(st)
    torch.cuda.synchronize()
print(which, "ms/step", e0.elapsed_time(e1) / steps)
