#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (name, grid) time and share."""
import collections
import csv
import re
import sys


def main(path, steps):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else (v * 1000 if unit in ("ms", "msecond") else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void ", "", name)[:70]
        key = (name, row["Grid Size"], row["Block Size"])
        agg.setdefault(key, [0, 0.0])
        agg[key][0] += 1
        agg[key][1] += v
        tot += v
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {tot / steps:.1f} us of kernel time per step ({steps} steps)")
    print(f"# {'us/step':>9} {'n/step':>6} {'share':>6}  kernel  grid  block")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / steps:10.1f} {n / steps:6.1f} {100 * t / tot:5.1f}%  {k[0]}  {k[1]}  {k[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
