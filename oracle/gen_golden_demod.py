#!/usr/bin/env python
"""Golden vectors for SURVEY.md 8f-4: PhaseDiscriminators (sdrbase/dsp/phasediscri.h:26-198), the AM magnitude lines of
AMDemod::processOneSample (amdemod.cpp:154-156,241) and a FileRecord .sdriq file (sdrbase/dsp/filerecord.cpp:72-148), produced
by the UNMODIFIED reference compiled in place (oracle/_ref/libsdrref*.so via oracle/ref_capi_tx.cpp; `make -C oracle ref`).
Run in the authoring container; the fixture travels as tests/golden/golden_demod.npz + .json."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402

SEED = 8844
CUTS = [0, 1, 2, 3, 1000, 6000]
SCALING = 0.37
FF_CASES = [(0, 300 / 48000.0, 3000 / 48000.0, 1024), (1, 0.0, 6000 / 48000.0, 2048), (0, 0.2, 0.05, 64)]
FF_OPS = [(1, True, False), (1, False, True), (0, True, True), (2, True, False), (2, True, True)]


def demod_input(seed, n):
    rs = np.random.RandomState(seed)
    x = ((rs.randn(n) + 1j * rs.randn(n)) * 8000).astype(np.complex64)
    # the corners of atan2_approximation2: x == 0 with y >, ==, < 0; |z| == 1; zeros
    x[100] = 0; x[200] = 5 + 0j; x[201] = 0 + 7j; x[202] = -3 + 0j; x[203] = 0 - 2j; x[204] = 4 + 4j; x[205] = -4 + 4j; x[300:310] = 0
    return x


def main():
    x = demod_input(SEED, CUTS[-1])
    arrays, meta = {}, {"seed": SEED, "cuts": CUTS, "scaling": SCALING}
    for strict in (False, True):
        tag = "strict" if strict else "fast"
        for kind in range(4):
            d = refbind.RefDemod(kind, SCALING, strict=strict)
            outs = [d.run(x[a:b]) for a, b in zip(CUTS[:-1], CUTS[1:])]
            for j, name in enumerate(("out", "aux0", "aux1")):
                arrays["demod/%s/%d/%s" % (tag, kind, name)] = np.concatenate([o[j] for o in outs])
    # fftfilt (SSB / DSB channel filter): the SSB demodulator's filter (ssbFftLen 1024, 300..3000 Hz at 48 kS/s), the DSB one
    # (2048), a short band-reject one; per-sample calls, two call splits, both sidebands, DC kept / rejected
    meta["fftfilt"] = {"cases": FF_CASES, "ops": FF_OPS, "seed": SEED + 2, "n": 6000, "split": 777}
    xf = demod_input(SEED + 2, 6000)
    for strict in (False, True):
        tag = "strict" if strict else "fast"
        for ci, (kind, f1, f2, flen) in enumerate(FF_CASES):
            arrays["fftfilt/%s/%d/filter" % (tag, ci)] = refbind.RefFftFilt(kind, f1, f2, flen, strict=strict).filter()
            for oi, (op, usb, dc) in enumerate(FF_OPS):
                f = refbind.RefFftFilt(kind, f1, f2, flen, strict=strict)
                arrays["fftfilt/%s/%d/%d" % (tag, ci, oi)] = np.concatenate([f.run(op, xf[:777], usb, dc), f.run(op, xf[777:], usb, dc)])
    rs = np.random.RandomState(SEED + 1)
    iq = rs.randint(-32768, 32768, size=(1000, 2)).astype(np.int16)
    path = os.path.join(tempfile.mkdtemp(), "golden.sdriq")
    count = refbind.filerecord_write(path, 2400000, 434000000, iq, 300)
    raw = np.frombuffer(open(path, "rb").read(), dtype=np.uint8).copy()
    hdr = refbind.filerecord_read_header(path)
    raw[12:20] = 0                                      # time(0) at recording start: masked
    arrays["sdriq/file"] = raw
    meta["sdriq"] = {"sample_rate": 2400000, "center_frequency": 434000000, "n_samples": 1000, "first_feed": 300, "seed": SEED + 1,
                     "byte_count_reported": count, "header_read": {k: hdr[k] for k in ("sample_rate", "center_frequency", "sample_size", "data_offset")}}
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_demod.npz"), **arrays)
    with open(os.path.join(ROOT, "tests", "golden", "golden_demod.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", len(arrays), "arrays")


if __name__ == "__main__":
    main()
