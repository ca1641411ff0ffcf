#!/usr/bin/env python
"""Developer aid (GPU box): ONE launch of each kernel that still lacks an `ncu --set full` summary, in one process, so that one
ncu run (`-k regex:...`) captures them all:
  hb48_chain_kernel        channels 0-127 of the 1024-channel plan (one rank's block of an 8-way sharded bank), one feed
  frontend_schedule_kernel the 64-channel plan (non-lattice ratios: the exact scan), one feed
  dc_correct_kernel        2^26 int16 IQ samples
  fftfilt_kernel           runSSB, 1024-point, 2^24 complex64 samples
usage: ncu_probe.py [chain] [bank64] [iqcorr] [ssbfilt]   (default: all four)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdrangel_b200 as S
from bench import plan64, plan1024
S.capi.init(0)
which = sys.argv[1:] or ["chain", "bank64", "iqcorr", "ssbfilt"]
st = torch.cuda.Stream()
sp = st.cuda_stream
cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))


def bank(fs, fcs, n):
    x = torch.randint(-2048, 2048, (2 * n,), dtype=torch.int16, device="cuda")
    b = S.DownChannelizerBank(fs)
    b.set_chunk(n)
    for fc in fcs:
        cid, rate, ofs, path = b.add_channel(48000, fc)
        b.set_frontend(cid, -ofs, cutoff, 48000)
    b.feed_dev(x.data_ptr(), n, sp)
    torch.cuda.synchronize()
    b.close()


with torch.cuda.stream(st):
    if "chain" in which:
        fs, fcs = plan1024()
        bank(fs, fcs[:128], 3 << 22)
    if "bank64" in which:
        fs, fcs = plan64()
        bank(fs, fcs, 3 << 22)
    if "iqcorr" in which:
        n = 1 << 26
        x = torch.randint(-2048, 2048, (n, 2), dtype=torch.int16, device="cuda")
        y = torch.empty_like(x)
        q = S.IQCorrections()
        q.run_dev(x.data_ptr(), y.data_ptr(), n, sp)
        torch.cuda.synchronize()
        q.close()
    if "ssbfilt" in which:
        n = 1 << 24
        x = torch.randn((n, 2), dtype=torch.float32, device="cuda") * 8000
        y = torch.empty((n, 2), dtype=torch.float32, device="cuda")
        o = S.FftFilt(0, 300 / 48000.0, 3000 / 48000.0, 1024)
        o.run_dev(1, x.data_ptr(), n, y.data_ptr(), n, usb=True, get_dc=False, stream=sp)
        torch.cuda.synchronize()
        o.close()
print("probe done:", which)
