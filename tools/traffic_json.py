#!/usr/bin/env python
"""Developer aid: profiles/rNN_traffic.json from an ncu launch list (gpu__time_duration.sum + dram__bytes_{read,write}.sum per
launch, `--csv`) of tools/bank_probe.py: the launches of ONE step (from one frontend_schedule_kernel to the next) are summed.
usage: traffic_json.py launches.csv samples_per_step workload out.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "ID")
recs = {}
for r in rows:
    if len(r) != len(hdr) or r[0] == "ID":
        continue
    d = dict(zip(hdr, r))
    k = int(d["ID"])
    recs.setdefault(k, {"name": d["Kernel Name"].split("(")[0].split("::")[-1]})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
ids = sorted(recs)
starts = [i for i in ids if recs[i]["name"].startswith("frontend_schedule_kernel")]
step = [i for i in ids if starts[0] <= i < starts[1]]
per = {}
for i in step:
    r = recs[i]
    e = per.setdefault(r["name"], {"launches": 0, "time_ns": 0.0, "dram_read": 0.0, "dram_write": 0.0})
    e["launches"] += 1
    e["time_ns"] += r.get("gpu__time_duration.sum", 0.0)
    e["dram_read"] += r.get("dram__bytes_read.sum", 0.0)
    e["dram_write"] += r.get("dram__bytes_write.sum", 0.0)
tot_r = sum(e["dram_read"] for e in per.values())
tot_w = sum(e["dram_write"] for e in per.values())
out = {}
try:
    out = json.load(open(sys.argv[4]))
except Exception:
    pass
out[sys.argv[3]] = {"samples": int(sys.argv[2]), "dram_read": tot_r, "dram_write": tot_w, "kernels": per,
                    "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (one step of tools/bank_probe.py; serialised, cold-cache: compare shares)"}
json.dump(out, open(sys.argv[4], "w"), indent=1)
print(json.dumps(out[sys.argv[3]], indent=1))
