#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, mean ns, share of GPU time.

usage: python tools/summarize_launches.py gpurun_out/launches_X.csv "command that produced it" > profiles/X_summary.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for name, ns, grid, blk in rows:
        short = re.sub(r"\(.*$", "", name).replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(short, [0, 0.0, grid, blk])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f"# {note}")
    print("# per-launch times are cold-cache and serialised under ncu: compare shares, not absolutes")
    for k, (n, t, grid, blk) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} launches={n:5d} avg_ns={t / n:12.1f} total_ns={t:14.1f} share={t / total:6.3f} grid={grid} block={blk}")
    print(f"# total launches {len(rows)}, total ns {total:.0f}")


if __name__ == "__main__":
    main()
