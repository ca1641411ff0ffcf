#!/usr/bin/env python
"""Golden vectors for Interpolator::interpolate / resample (interpolator.h:39-76) and the stand-alone NCO (nco.h:40-53),
produced by the UNMODIFIED reference compiled in place (oracle/_ref/libsdrref*.so; `make -C oracle ref`).  Run in the
authoring container (needs /root/reference); the fixture travels as tests/golden/golden_interp.npz + .json.

Cases: (rate_in, rate_out) pairs on both sides of 1 -- the Tx plugins' interpolation (48 kS/s audio -> channel rate), the AM
demodulator's Rx interpolation (channel rate < audio rate, amdemod.cpp:113-124) and decimation through `resample`."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402

CASES = [(48000, 156250), (48000, 60000), (32000, 48000), (48000, 48000), (156250, 48000), (60000, 48000)]
CUTS = [0, 1, 2, 700, 701, 5000, 12000]
SEED = 4242


def main():
    rs = np.random.RandomState(SEED)
    n = CUTS[-1]
    x = (rs.randint(-20000, 20000, size=n) + 1j * rs.randint(-20000, 20000, size=n)).astype(np.complex64)
    arrays, meta = {}, {"seed": SEED, "n": n, "cuts": CUTS, "cases": CASES, "modes": {"interpolate": 1, "resample": 2}}
    for strict in (False, True):
        tag = "strict" if strict else "fast"
        for rin, rout in CASES:
            cutoff = float(np.float32(min(rin, rout) / 2.2))
            for name, mode in (("interpolate", 1), ("resample", 2)):
                if mode == 1 and rin > rout:
                    continue                      # the Tx loop only interpolates (distance <= 1)
                fe = refbind.RefFrontEnd(0, max(rin, rout), rout, cutoff, strict=strict)
                # Interpolator::create(16, sampleRate = the higher rate, cutoff); distance = rin / rout
                fe.lib.ref_frontend_destroy(fe.h)
                dist = float(np.float32(np.float32(rin) / np.float32(rout)))
                fe.h = fe.lib.ref_frontend_create(0.0, float(max(rin, rout)), 16, float(max(rin, rout)), cutoff, 4.5, dist)
                outs = [fe.run_c64(mode, x[a:b]) for a, b in zip(CUTS[:-1], CUTS[1:])]
                key = "%s/%s/%d_%d" % (name, tag, rin, rout)
                arrays[key + "/out"] = np.concatenate(outs)
                arrays[key + "/counts"] = np.array([o.shape[0] for o in outs], dtype=np.int64)
                arrays[key + "/remain"] = np.array([fe.remain()], dtype=np.float32)
    # Decimators<qint32,qint16,16,12>: the split-I/Q overloads and decimate2_u (decimators.h:359-461,2638-3888), state carried over 3 calls
    rs2 = np.random.RandomState(SEED + 1)
    sx = rs2.randint(-2048, 2048, size=2 * 6000).astype(np.int16)
    meta["split"] = {"seed": SEED + 1, "n_scalars": 2 * 6000, "cuts_samples": [0, 1000, 1003, 6000]}
    for log2 in range(7):
        d = refbind.RefDecimators("ii", 12)
        outs = []
        for a, b in ((0, 1000), (1000, 1003), (1003, 6000)):
            outs.append(d.run_split(log2, sx[2 * a:2 * b:2], sx[2 * a + 1:2 * b:2]))
        arrays["split/cen/%d" % log2] = np.concatenate(outs)
        arrays["split/cen/%d/counts" % log2] = np.array([o.shape[0] for o in outs], dtype=np.int64)
    d = refbind.RefDecimators("ii", 12)
    arrays["split/2u"] = d.run_split(1, sx[0::2], sx[1::2], u=True)
    for bits in (8, 12, 16):
        arrays["dec2u/%d" % bits] = refbind.RefDecimators("ii", bits).run_2u(sx)
    for freq, rate in ((15433.0, 156250.0), (-4321.0, 60000.0), (0.0, 48000.0), (1e6, 10e6)):
        arrays["nco/%g_%g" % (freq, rate)] = refbind.nco_block(freq, rate, 5000)
    meta["nco_cases"] = [[15433.0, 156250.0], [-4321.0, 60000.0], [0.0, 48000.0], [1e6, 10e6]]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_interp.npz"), **arrays)
    with open(os.path.join(ROOT, "tests", "golden", "golden_interp.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", len(arrays), "arrays")


if __name__ == "__main__":
    main()
