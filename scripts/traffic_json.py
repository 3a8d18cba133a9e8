"""ncu per-launch DRAM traffic CSVs (scripts/profile_r2.sh, --cache-control none) -> profiles/r2_traffic.json, the file bench.py
reads for roofline.traffic.  Usage: python scripts/traffic_json.py <tag> (reads gpurun_out/<tag>_traffic_*.csv)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2f"
SRC = {"LL_p0": (f"{tag}_traffic_LL_chain.csv", 11), "VS_p1": (f"{tag}_traffic_VS_tc.csv", 29)}   # (file, launches per step)
out = {}
for key, (fn, per_step) in SRC.items():
    path = os.path.join(ROOT, "gpurun_out", fn)
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", fn)
    if not os.path.exists(path):
        continue
    rows = [l for l in open(path) if l.startswith('"')]
    rd, wr, us = {}, {}, {}
    name = {}
    for r in csv.DictReader(rows):
        i = int(r["ID"])
        name[i] = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r["Metric Unit"]]
        if r["Metric Name"] == "dram__bytes_read.sum":
            rd[i] = v
        elif r["Metric Name"] == "dram__bytes_write.sum":
            wr[i] = v
        elif r["Metric Name"] == "gpu__time_duration.sum":
            us[i] = v / (1e3 if r["Metric Unit"] in ("ns", "nsecond") else 1.0)
    n = len(name)
    steps = n / per_step
    by = {}
    for i, k in name.items():
        by[k] = by.get(k, 0.0) + rd.get(i, 0.0) + wr.get(i, 0.0)
    out[key] = {"launches": n, "steps": steps,
                "dram_read_bytes_per_step": sum(rd.values()) / steps, "dram_write_bytes_per_step": sum(wr.values()) / steps,
                "by_kernel_bytes_per_step": {k: v / steps for k, v in by.items()},
                "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none over "
                          f"{steps:g} consecutive steps of bench.py's timed region (scripts/profile_r2.sh, {fn}); caches NOT flushed "
                          "between kernels, so parameters / Adam state / activations stay in L2 as in the running step"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print(json.dumps({k: {kk: v[kk] for kk in ("launches", "steps", "dram_read_bytes_per_step", "dram_write_bytes_per_step")} for k, v in out.items()}, indent=1))
