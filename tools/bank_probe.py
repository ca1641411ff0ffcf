#!/usr/bin/env python
"""Developer aid (GPU box): step time of the 1024-channel bank with and without the front-ends."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdrangel_b200 as S
from bench import plan1024
S.capi.init(0)
fs, fcs = plan1024()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3 << 24
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, len(fcs))
fcs = fcs[lo:hi]
x = torch.randint(-2048, 2048, (2 * n,), dtype=torch.int16, device="cuda")
cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
for fe in (False, True):
    b = S.DownChannelizerBank(fs)
    b.set_chunk(n)
    for fc in fcs:
        cid, rate, ofs, path = b.add_channel(48000, fc)
        if fe:
            b.set_frontend(cid, -ofs, cutoff, 48000)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            b.feed_dev(x.data_ptr(), n, st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(st)
        for _ in range(5):
            b.feed_dev(x.data_ptr(), n, st.cuda_stream)
        e1.record(st)
        host = time.perf_counter() - t0
        torch.cuda.synchronize()
    print("frontends" if fe else "tree only", "ms/step", e0.elapsed_time(e1) / 5, "host ms/step", host / 5 * 1e3)
    b.close()
