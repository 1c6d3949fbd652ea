#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page details --csv` -> compact markdown (one table per captured launch).
usage: ncu -i rep --page details --csv | python profiles/ncu_details_to_md.py > profiles/x.md"""
import csv
import sys
from collections import OrderedDict

KEEP = ("Duration", "SM Frequency", "DRAM Frequency", "Memory Throughput", "DRAM Throughput", "L2 Cache Throughput",
        "L1/TEX Cache Throughput", "Compute (SM) Throughput", "Executed Ipc Active", "Issue Slots Busy",
        "SM Busy", "Mem Busy", "Max Bandwidth", "L1/TEX Hit Rate", "L2 Hit Rate", "Mem Pipes Busy", "Registers Per Thread",
        "Dynamic Shared Memory Per Block", "Grid Size", "Block Size", "Cluster Size", "Theoretical Occupancy",
        "Achieved Occupancy", "Waves Per SM", "One or More Eligible", "No Eligible", "Avg. Active Threads Per Warp",
        "Local Memory Spilling Requests")


def main():
    launches = OrderedDict()
    for r in csv.DictReader(line for line in sys.stdin if not line.startswith("==")):
        key = (r["ID"], r["Kernel Name"], r["Grid Size"], r["Block Size"])
        launches.setdefault(key, []).append(r)
    for (i, name, grid, block), rows in launches.items():
        print(f"### launch {i}: `{name}` grid {grid} block {block}\n")
        print("| section | metric | value | unit |")
        print("|---|---|---:|---|")
        seen = set()
        for r in rows:
            m = r["Metric Name"]
            if m in KEEP and (r["Section Name"], m) not in seen:
                seen.add((r["Section Name"], m))
                print(f"| {r['Section Name']} | {m} | {r['Metric Value']} | {r['Metric Unit']} |")
        rules = [r for r in rows if r.get("Rule Description") and r.get("Rule Type") in ("OPT", "WRN")]
        if rules:
            print("\nncu rule hits:")
            for r in rules[:8]:
                d = " ".join(r["Rule Description"].split())
                print(f"* **{r['Rule Name']}**: {d[:400]}")
        print()


if __name__ == "__main__":
    main()
