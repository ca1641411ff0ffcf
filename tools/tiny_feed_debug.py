"""Scratch triage: the tiny-feed sequence of tests/test_edge_cases_gpu.py, reporting the first feed that fails."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdrangel_b200 import DownChannelizerBank, capi
capi.init(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
fs = 10_000_000
x = np.random.RandomState(77).randint(-30000, 30000, size=(60_000, 2)).astype(np.int16)
cutoff = float(np.float32(np.float32(12500) / np.float32(2.2)))
specs = [(1_000_000, 600_000), (48_000, 1_234_567), (48_000, -3_000_000), (20_000_000, 0)]
if which == "nofe":
    fe_on = False
else:
    fe_on = True
if which == "nos0":
    specs = specs[:3]
b = DownChannelizerBank(fs)
ids = []
for req, fc in specs:
    cid, rate, ofs, path = b.add_channel(req, fc)
    if req == 48_000 and fe_on:
        b.set_frontend(cid, -ofs, cutoff, 48000)
    ids.append(cid)
pos = 0
for n in (0, 1, 1, 2, 3, 0, 5, 127, 1, 128, 1000, 0, 7, 20_001, 2, 30_000):
    try:
        b.feed(x[pos:pos + n])
        for cid in ids:
            b.fetch(cid)
        print(which, "feed", n, "ok", flush=True)
    except Exception as e:
        print(which, "feed", n, "at pos", pos, "FAILED:", e, flush=True)
        break
    pos += n
