#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (launches, total time, share).
  python tools/ncu_launch_summary.py gpurun_out/launches.csv [header comment ...] > profiles/rN_ncu_launch_summary.txt"""
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows:
        if r is hdr or len(r) <= vi or r[ki] == "Kernel Name":
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void |emd::|\(anonymous namespace\)::|<unnamed>::|unnamed>::", "", name).strip()
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + v)
    tot = sum(t for _, t in agg.values())
    for c in sys.argv[2:]:
        print("#", c)
    print(f"{'kernel':64s} {'launches':>8s} {'total_us':>10s} {'share':>6s}")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:64]:64s} {n:8d} {t:10.1f} {100 * t / tot:5.1f}%")
    print(f"{'TOTAL':64s} {sum(n for n, _ in agg.values()):8d} {tot:10.1f}")


if __name__ == "__main__":
    main()
