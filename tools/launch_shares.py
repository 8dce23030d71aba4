#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.

    python tools/launch_shares.py profiles/r01_launches_ncu_head.csv > profiles/r01_launch_shares_head.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        name = re.sub(r"^void ", "", r["Kernel Name"]).split("(")[0]
        rows.append((name + " grid" + r["Grid Size"].replace(" ", ""), ns / 1e3))
    tot = sum(t for _, t in rows)
    agg = OrderedDict()
    for k, t in rows:
        n, s = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, s + t)
    print(f"# {len(rows)} launches, total {tot:.1f} us (ncu per-launch times: cold-cache, serialised -- compare SHARES)")
    for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:58s} n={n:3d} avg {s / n:9.1f} us  sum {s:10.1f} us  {100 * s / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
