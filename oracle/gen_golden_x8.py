#!/usr/bin/env python
"""ORACLE (test infrastructure): golden vectors for the 8-bit device decimators, generated from the UNMODIFIED reference
compiled in place (oracle/_ref/libsdrref.so): Decimators<qint32,qint8,16,8> (HackRF) and DecimatorsU<qint32,quint8,16,8,127>
(RTL-SDR).  Writes tests/golden/golden_x8.{npz,json}.

Run in the build container (needs /root/reference):   make -C oracle ref && python oracle/gen_golden_x8.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refbind as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
MODES = {"inf": 0, "sup": 1, "cen": 2}
DT = {"i8": np.int8, "u8": np.uint8}


def main():
    meta = {"generator": "oracle/gen_golden_x8.py", "build": R.load().ref_build_info().decode(),
            "hash": "FNV-1a-64 over uint16 words (re then im)"}
    arrays = {}
    # streaming: full-range random bytes, awkward call splits (odd lengths, partial blocks), one object per (kind, log2, mode)
    rs = np.random.RandomState(20181019)
    raw = rs.randint(0, 256, size=2 * 3000).astype(np.uint8)
    raw[200:264] = 255                       # runs of the extreme codes
    raw[900:964] = 0
    raw[1500:1564] = 128
    cuts = [0, 1000, 1000 + 2 * 333 + 1, 3500, 3502, raw.size]
    meta["stream"] = {"seed": 20181019, "n_scalars": int(raw.size), "cuts": cuts, "input": "stream/raw (uint8; the int8 case views the same bytes)",
                      "outputs": "stream/<kind>/<log2>/<mode>", "counts": "stream_counts/<kind>/<log2>/<mode>"}
    arrays["stream/raw"] = raw
    for kind, dt in DT.items():
        x = raw.view(dt)
        for log2 in range(0, 7):
            for mname, mode in MODES.items():
                d = R.RefDecimators(kind)
                outs = [d.run(log2, mode, x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
                arrays[f"stream/{kind}/{log2}/{mname}"] = np.concatenate(outs)
                arrays[f"stream_counts/{kind}/{log2}/{mname}"] = np.array([o.shape[0] for o in outs], dtype=np.int32)
    # one long call per row: the sdrbench int16 buffer's low bytes (2^20 IQ samples), hash only
    buf = R.sdrbench_s16(1 << 20)
    lo = (buf.astype(np.int32) & 0xff).astype(np.uint8)
    meta["long"] = {"input": "low byte of the sdrbench int16 buffer (mt19937 default seed), 2^20 IQ samples", "input_fnv_u8pairs": R.fnv1a64_u16(lo.view(np.uint16)),
                    "rows": {}}
    for kind, dt in DT.items():
        for log2 in range(0, 7):
            for mname, mode in MODES.items():
                out = R.RefDecimators(kind).run(log2, mode, lo.view(dt))
                meta["long"]["rows"][f"{kind}/{log2}/{mname}"] = {"n_out": int(out.shape[0]), "fnv": R.fnv1a64_u16(out), "at100": out[100].tolist()}
    # DSPDeviceSourceEngine::iqCorrections, DC branch: ragged calls across the 1024-sample fill-up, offsets, extreme runs
    rs = np.random.RandomState(20181020)
    xq = (rs.randint(-3000, 3000, size=(6000, 2)) + np.array([700, -1234])).astype(np.int16)
    xq[2500:2600] = 32767
    xq[4000:4100] = -32768
    qcuts = [0, 7, 1000, 1030, 1031, 3333, 6000]
    q = R.RefIQCorrections()
    outs = [q.run(xq[a:b]) for a, b in zip(qcuts[:-1], qcuts[1:])]
    arrays["iqcorr/in"] = xq
    arrays["iqcorr/out"] = np.concatenate(outs)
    meta["iqcorr"] = {"seed": 20181020, "cuts": qcuts, "input": "iqcorr/in", "output": "iqcorr/out", "mode": "DC only (imbalanceCorrection false)",
                      "fnv": R.fnv1a64_u16(arrays["iqcorr/out"])}
    # ... and its I/Q imbalance branch (floating-point flavour): a tone with gain and phase imbalance, DC and noise
    rs = np.random.RandomState(20181021)
    t = np.arange(8000)
    xi_ = 3000 * np.cos(2 * np.pi * 0.0137 * t) + 500 + rs.randn(t.size) * 50
    xq_ = 0.8 * 3000 * np.sin(2 * np.pi * 0.0137 * t + 0.2) - 300 + rs.randn(t.size) * 50
    xm = np.stack([xi_, xq_], 1).astype(np.int16)
    mcuts = [0, 7, 1000, 1030, 1031, 4444, 8000]
    q = R.RefIQCorrections()
    outs = [q.run(xm[a:b], True) for a, b in zip(mcuts[:-1], mcuts[1:])]
    qs = R.RefIQCorrections(strict=True)
    outs_s = [qs.run(xm[a:b], True) for a, b in zip(mcuts[:-1], mcuts[1:])]
    arrays["iqcorr_imb/in"] = xm
    arrays["iqcorr_imb/out"] = np.concatenate(outs)
    arrays["iqcorr_imb/out_strict"] = np.concatenate(outs_s)
    meta["iqcorr_imb"] = {"seed": 20181021, "cuts": mcuts, "input": "iqcorr_imb/in", "output": "iqcorr_imb/out (reference flags), iqcorr_imb/out_strict (-O2, no -ffast-math)",
                          "mode": "imbalanceCorrection true, floating-point flavour (IMBALANCE_INT undefined)", "fnv": R.fnv1a64_u16(arrays["iqcorr_imb/out"])}
    np.savez_compressed(os.path.join(OUT, "golden_x8.npz"), **arrays)
    with open(os.path.join(OUT, "golden_x8.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote %d arrays, golden_x8.npz %.0f KiB" % (len(arrays), os.path.getsize(os.path.join(OUT, "golden_x8.npz")) / 1024))


if __name__ == "__main__":
    main()
