#!/usr/bin/env python
"""ORACLE (test infrastructure): generate tests/golden/* from the UNMODIFIED reference compiled in place
(oracle/_ref/libsdrref.so = reference flags -O3 -ffast-math; libsdrref_strict.so = same sources, -O2 strict IEEE).

Run in the build container (needs /root/reference):   make -C oracle ref && python oracle/gen_golden.py
The fixtures are what travels: the GPU box has no /root/reference.  Inputs are either the sdrbench generator
(mt19937 default seed) or numpy's frozen legacy RandomState with the seeds written below.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refbind as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
MODES = {"inf": 0, "sup": 1, "cen": 2}


def chan_plan_64():
    return [((c - 32) * 150000 + 6250) for c in range(64)]


def chan_plan_1024():
    return [((c - 512) * 120000 + 60000 + 1250 * ((c % 9) - 4)) for c in range(1024)]


def main():
    meta = {"generator": "oracle/gen_golden.py", "build": R.load().ref_build_info().decode(),
            "hash": "FNV-1a-64 over uint16 words (re then im), SURVEY.md Appendix D convention"}
    arrays = {}

    # ---- 1. sdrbench decimateii: Appendix D table, N = 2^20, fresh object per row, one call
    buf = R.sdrbench_s16(1 << 20)
    meta["sdrbench_s16"] = {"n_samples": 1 << 20, "first4": buf[:4].tolist(), "fnv": R.fnv1a64_u16(buf)}
    table = {}
    for bits in (8, 12, 16):
        for log2 in range(0, 7):
            for mname, mode in MODES.items():
                out = R.RefDecimators("ii", bits).run(log2, mode, buf)
                table[f"{bits}/{log2}/{mname}"] = {"n_out": int(out.shape[0]), "fnv": R.fnv1a64_u16(out),
                                                   "head": out[:4].ravel().tolist(), "at100": out[100].tolist()}
    meta["decim_ii_sdrbench"] = table

    # ---- 2. decimateii streaming: full-scale random int16, awkward call splits, one object per (bits, log2, mode)
    rs = np.random.RandomState(20181018)
    x = rs.randint(-32768, 32768, size=2 * 6000).astype(np.int16)
    cuts = [0, 1000, 1000 + 2 * 333 + 1, 7000, 7002, x.size]
    meta["decim_ii_stream"] = {"seed": 20181018, "n_scalars": int(x.size), "cuts": cuts, "outputs": "decim_ii_stream/<bits>/<log2>/<mode>"}
    for bits in (8, 12, 16):
        for log2 in range(0, 7):
            for mname, mode in MODES.items():
                d = R.RefDecimators("ii", bits)
                outs = [d.run(log2, mode, x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
                arrays[f"decim_ii_stream/{bits}/{log2}/{mname}"] = np.concatenate(outs)
                arrays[f"decim_ii_stream_counts/{bits}/{log2}/{mname}"] = np.array([o.shape[0] for o in outs], dtype=np.int32)

    # ---- 3. mode / factor switching on one object (state shared across modes, decimators.h:334-339)
    rs = np.random.RandomState(7)
    sched = [(int(rs.randint(0, 7)), int(rs.randint(0, 3)), int(rs.randint(1, 700))) for _ in range(120)]
    d = R.RefDecimators("ii", 12)
    pos, outs = 0, []
    for log2, mode, n in sched:
        outs.append(d.run(log2, mode, x[pos:pos + n]))
        pos = (pos + n) % (x.size - 800)
    meta["decim_ii_switch"] = {"input": "same x as decim_ii_stream", "seed_sched": 7, "schedule": sched}
    arrays["decim_ii_switch/out"] = np.concatenate(outs)
    arrays["decim_ii_switch/counts"] = np.array([o.shape[0] for o in outs], dtype=np.int32)

    # ---- 4. float cascades (FI/FF/IF): strict build = bit-exact target, fast build = tolerance target
    fb = R.sdrbench_f32(1 << 14)
    ib = R.sdrbench_s16(1 << 14)
    meta["sdrbench_f32"] = {"n_samples": 1 << 14, "first3": [float(v) for v in fb[:3]], "fnv": R.fnv1a64_u16(fb)}
    fcuts = [0, 10001, 10002, fb.size]
    meta["decim_f"] = {"n_samples": 1 << 14, "cuts": fcuts, "strict": "decim_f/strict/<kind>/<log2>/<mode>",
                       "fast": "decim_f/fast/<kind>/<log2>/<mode>"}
    for kind in ("fi", "ff", "if"):
        src = ib if kind[0] == "i" else fb
        for log2 in range(0, 7):
            for mname, mode in MODES.items():
                for tag, strict in (("strict", True), ("fast", False)):
                    d = R.RefDecimators(kind, 12, strict=strict)
                    outs = [d.run(log2, mode, src[a:b]) for a, b in zip(fcuts[:-1], fcuts[1:])]
                    o = np.concatenate(outs)
                    if log2 >= 2:
                        arrays[f"decim_f/{tag}/{kind}/{log2}/{mname}"] = o
                    else:  # large outputs of the trivial factors: hash only (float hashed by its uint16 halves)
                        meta.setdefault("decim_f_hash", {})[f"{tag}/{kind}/{log2}/{mname}"] = {
                            "n_out": int(o.shape[0]), "fnv": R.fnv1a64_u16(o)}
    # config 2 self-check values: 2^20 samples, log2=6 cen, fast build
    o = R.RefDecimators("fi").run(6, 2, R.sdrbench_f32(1 << 20))
    meta["decimatefi_config2"] = {"n_out": int(o.shape[0]), "out100_102": o[100:103].tolist(),
                                  "fnv_fast": R.fnv1a64_u16(o),
                                  "fnv_strict": R.fnv1a64_u16(R.RefDecimators("fi", strict=True).run(6, 2, R.sdrbench_f32(1 << 20)))}

    # ---- 5. DownChannelizer: filter-chain plans + streaming feed
    plans = {}
    for name, fs, fcs in (("bank64", 10_000_000, chan_plan_64()), ("bank1024", 122_880_000, chan_plan_1024())):
        rows = []
        for fc in fcs:
            rate, ofs, modes = R.RefDownChannelizer().configure(fs, 48000, fc)
            rows.append([fc, rate, ofs, "".join("CLU"[m] for m in modes)])
        plans[name] = {"input_rate": fs, "requested_rate": 48000, "channels": rows}
    rs = np.random.RandomState(11)
    rnd = []
    for _ in range(200):
        fs = int(rs.choice([10_000_000, 122_880_000, 2_400_000, 61_440_000, 336_000]))
        req = int(rs.choice([48000, 12500, 200000, 64000, 8000]))
        fc = int(rs.randint(-fs // 2, fs // 2))
        rate, ofs, modes = R.RefDownChannelizer().configure(fs, req, fc)
        rnd.append([fs, req, fc, rate, ofs, "".join("CLU"[m] for m in modes)])
    plans["random"] = rnd
    meta["chan_plans"] = plans

    rs = np.random.RandomState(12)
    cx = rs.randint(-32768, 32768, size=(60000, 2)).astype(np.int16)
    cx[100:130] = -32768   # exercises the -(-32768) int16 wrap of the rotated store (inthalfbandfiltereo.h:164)
    ccuts = [0, 1000, 32769, 32770, 60000]
    meta["chan_feed"] = {"seed": 12, "n": 60000, "min_run": [100, 130], "cuts": ccuts, "input_rate": 10_000_000,
                         "requested_rate": 48000, "offsets": [1234567, -4000000, 0, 17, 2499999, 3300000]}
    for fc in meta["chan_feed"]["offsets"]:
        c = R.RefDownChannelizer()
        c.configure(10_000_000, 48000, fc)
        arrays[f"chan_feed/{fc}"] = np.concatenate([c.feed(cx[a:b]) for a, b in zip(ccuts[:-1], ccuts[1:])])

    # ---- 6. plugin front-end: NCO + Interpolator::decimate (fast build = what a deployment runs; strict too)
    rs = np.random.RandomState(13)
    fx = rs.randint(-20000, 20000, size=(20000, 2)).astype(np.int16)
    fe_cases = [(15433, 156250, 48000), (-2500, 60000, 48000), (0, 78125, 48000), (-15625, 156250, 48000)]
    meta["frontend"] = {"seed": 13, "n": 20000, "cuts": [0, 7, 9000, 20000], "cutoff": "float32(12500/2.2f)",
                        "cases": fe_cases}
    cutoff = np.float32(np.float32(12500) / np.float32(2.2))
    for (freq, rate, outr) in fe_cases:
        for tag, strict in (("fast", False), ("strict", True)):
            fe = R.RefFrontEnd(freq, rate, outr, cutoff, strict=strict)
            key = f"frontend/{tag}/{freq}_{rate}"
            arrays[key + "/taps"] = fe.taps()
            o, i, p = zip(*[fe.feed(fx[a:b], True) for a, b in ((0, 7), (7, 9000), (9000, 20000))])
            arrays[key + "/out"] = np.concatenate(o)
            offs = np.cumsum([0, 7, 8993])
            arrays[key + "/idx"] = np.concatenate([ii + off for ii, off in zip(i, offs)]).astype(np.int32)
            arrays[key + "/phase"] = np.concatenate(p).astype(np.int32)
            meta.setdefault("frontend_inc", {})[f"{freq}_{rate}"] = fe.nco_increment()
    arrays["nco_table"] = R.nco_table()

    # ---- 7. SpectrumVis
    rs = np.random.RandomState(14)
    n = 4096 * 21 + 100
    sx = rs.randint(-2048, 2048, size=(n, 2)).astype(np.int16)
    t = np.arange(n)
    tone = 1500 * np.exp(2j * np.pi * 0.1234 * t)
    sx[:, 0] += tone.real.astype(np.int16)
    sx[:, 1] += tone.imag.astype(np.int16)
    scuts = [0, 5000, 5001, n]
    cases = [(4096, 2, 10, False, False), (4096, 0, 0, False, False), (4096, 1, 10, False, False),
             (1024, 2, 3, False, True), (256, 1, 4, True, False), (64, 0, 0, True, True), (128, 2, 10, True, False)]
    meta["spectrum"] = {"seed": 14, "n": n, "tone": [1500, 0.1234], "cuts": scuts, "window": 1,
                        "cases(fft,avg_mode,avg_nb,linear,positive_only)": cases}
    for (fft, mode, nb, linear, posonly) in cases:
        for tag, strict in (("fast", False), ("strict", True)):
            s = R.RefSpectrumVis(strict=strict)
            s.configure(fft, 0, nb, mode, 1, linear)
            fr = np.concatenate([s.feed(sx[a:b], posonly) for a, b in zip(scuts[:-1], scuts[1:])])
            if fr.shape[0] > 6:   # keep fixtures small: first 3 and last 3 frames + count
                keep = np.concatenate([fr[:3], fr[-3:]])
            else:
                keep = fr
            key = f"spectrum/{tag}/{fft}_{mode}_{nb}_{int(linear)}_{int(posonly)}"
            arrays[key] = keep
            arrays[key + "/nframes"] = np.array([fr.shape[0]], dtype=np.int32)
    for fn in range(6):
        s = R.RefSpectrumVis()
        s.configure(256, 0, 0, 0, fn, False)
        arrays[f"window/{fn}_256"] = s.window()
    s = R.RefSpectrumVis(strict=True)
    s.configure(4096)
    arrays["window/1_4096"] = s.window()
    rs = np.random.RandomState(15)
    for nfft in (64, 128, 4096):
        xx = (rs.randn(nfft) + 1j * rs.randn(nfft)).astype(np.complex64)
        for tag, strict in (("fast", False), ("strict", True)):
            s = R.RefSpectrumVis(strict=strict)
            s.configure(nfft)
            arrays[f"fft/{tag}/{nfft}/in"] = xx
            arrays[f"fft/{tag}/{nfft}/out"] = s.fft(xx)

    np.savez_compressed(os.path.join(OUT, "golden.npz"), **arrays)
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1)
    sz = os.path.getsize(os.path.join(OUT, "golden.npz"))
    print(f"wrote {len(arrays)} arrays, golden.npz {sz/1024:.0f} KiB")


if __name__ == "__main__":
    main()
