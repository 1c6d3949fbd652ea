#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    return name.replace("void ", "").strip()[-70:]


def main(path):
    rows = []
    with open(path, newline="") as fp:
        lines = [ln for ln in fp if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((short(r["Kernel Name"]), v))
    agg = OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    print(f"launches: {len(rows)}; summed gpu__time_duration: {tot:.2f} ms (serialised, cold-cache, under ncu)\n")
    print("| kernel | launches | total ms | share | max ms |")
    print("|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f}% | {a[2]:.3f} |")


if __name__ == "__main__":
    main(sys.argv[1])
