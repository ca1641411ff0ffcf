#!/usr/bin/env python
"""Developer aid: print the key metrics + region breakdown of an .ncu-rep (run where ncu is installed).
usage: ncu_summary.py report.ncu-rep [k]    k = index of the captured launch within the report (default 0)"""
import csv, subprocess, sys, io
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2 + which]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
print("kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print("%-70s %-12s %s" % (k, units[i], r[i]))
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and float(r[i] or 0) > 0.05:
        print("%-70s %s" % (h, r[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
ks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]      # several captured launches: the k-th
secs = []
for a, b in zip(ks[:-1], ks[1:]):          # (ncu prints a launch's section twice when the report carries imported source)
    if not secs or rows[secs[-1][0]:secs[-1][1]] != rows[a:b]:
        secs.append((a, b))
rows = rows[secs[which][0]:secs[which][1]]
hdr, data = rows[1], [r for r in rows[2:] if len(r) > 10]
iS, iI, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(x[iI]) for x in data)
print("total warp instr", tot)
prev, start, regions = None, 0, []
for k, x in enumerate(data):
    c = int(x[iI])
    if prev is None or abs(c - prev) > 0.02 * max(c, prev, 1):
        if prev is not None:
            regions.append((start, k - 1, prev))
        start, prev = k, c
regions.append((start, len(data) - 1, prev))
for a, b, c in regions:
    n = b - a + 1
    if n * c / max(tot, 1) > 0.004:
        ops = {}
        for x in data[a:b + 1]:
            t = x[iS].split()
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda z: -z[1])[:9]
        smp = sum(int(x[iN]) for x in data[a:b + 1])
        print("lines %d-%d n=%d exec=%d share=%.3f samples=%.3f %s" % (a, b, n, c, n * c / tot, smp / max(sum(int(x[iN]) for x in data), 1), top))
