#!/usr/bin/env python
"""Developer aid (GPU box): run small decimator cases against the oracle port and print where they first differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import subprocess
subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
from oracle import portbind as P
import sdrangel_b200 as S
S.capi.init(0)
MODES = {"inf": 0, "sup": 1, "cen": 2}
rs = np.random.RandomState(1)
bad = 0
for kind, cls in (("ii", S.Decimators), ("ff", S.DecimatorsFF), ("fi", S.DecimatorsFI), ("if", S.DecimatorsIF)):
    for n in (6000, 200_000):
        if kind[0] == "i":
            x = rs.randint(-32768, 32768, size=2 * n).astype(np.int16)
        else:
            x = (rs.rand(2 * n) * 2 - 1).astype(np.float32)
        for log2 in range(0, 7):
            for mname, mode in MODES.items():
                d, o = cls(12), P.PortDecimators(kind, 12)
                if kind != "ii":
                    d.set_exact_float(True)
                cuts = [0, 1000, 1000 + 2 * 333 + 1, x.size // 2 + 2, x.size]
                for a, b in zip(cuts[:-1], cuts[1:]):
                    got, want = d.run(log2, mode, x[a:b]), o.run(log2, mode, x[a:b])
                    if got.shape != want.shape or not np.array_equal(got, want):
                        bad += 1
                        if got.shape == want.shape:
                            w = np.nonzero(np.any(got != want, axis=1))[0]
                            print(f"MISMATCH {kind} n={n} log2={log2} {mname} call[{a}:{b}] n_out={got.shape[0]} "
                                  f"first={w[0]} count={w.size} last={w[-1]} got={got[w[0]].tolist()} want={want[w[0]].tolist()}")
                        else:
                            print(f"SHAPE {kind} n={n} log2={log2} {mname} call[{a}:{b}] {got.shape} vs {want.shape}")
                        break
print("bad cases:", bad)
sys.exit(1 if bad else 0)
