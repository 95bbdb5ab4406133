#!/usr/bin/env python
"""Turn an `ncu --csv --metrics ...` log (one row per kernel x metric) into a per-kernel markdown table:
launches, total / average device time, DRAM bytes per launch, DRAM GB/s, % of the measured HBM peak, SM and DRAM
throughput %, registers.  Usage: ncu_table.py <log.csv> [peak_GBps] > profiles/rNN_kernel_table.md"""
import csv
import io
import json
import os
import re
import sys
from collections import OrderedDict, defaultdict


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def main():
    path = sys.argv[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    per_launch = OrderedDict()
    for r in rows:
        key = r["ID"]
        d = per_launch.setdefault(key, {"name": short(r["Kernel Name"]), "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)  # -> us
        if m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[m] = v
    agg = defaultdict(lambda: defaultdict(float))
    for d in per_launch.values():
        a = agg[d["name"]]
        a["n"] += 1
        for k, v in d.items():
            if isinstance(v, float):
                a[k] += v
        a["grid"] = d["grid"]
        a["block"] = d["block"]
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    print("| kernel | launches | total us | share | avg us | DRAM R+W per launch | DRAM GB/s | %% of %.0f GB/s | SM thr %% | DRAM thr %% | regs |" % peak)
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = a["n"]
        t = a["gpu__time_duration.sum"]
        by = a.get("dram__bytes_read.sum", 0.0) + a.get("dram__bytes_write.sum", 0.0)
        gbs = by / (t * 1e-6) / 1e9 if t > 0 else 0.0
        print("| %s | %d | %.1f | %.3f | %.2f | %.3f MB | %.0f | %.1f | %.1f | %.1f | %d |" % (
            name, n, t, t / tot if tot else 0, t / n, by / n / 1e6, gbs, 100 * gbs / peak,
            a.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0) / n,
            a.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0) / n,
            int(a.get("launch__registers_per_thread", 0) / n)))


if __name__ == "__main__":
    main()
