#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.

    python tools/launch_summary.py gpurun_out/launches_r01.csv > profiles/r01_launches.txt
"""
import csv
import sys
from collections import OrderedDict


def main(path, skip=0):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1 + skip:]:
        name = r[k].split("(")[0][:90]
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += float(r[v]) / 1e3
    tot = sum(d[1] for d in agg.values())
    print(f"# {path}: {sum(d[0] for d in agg.values())} launches, {tot:.1f} us total (per-launch times are cold-cache and "
          f"serialised: compare SHARES)")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us:10.1f} us {100 * us / tot:5.1f}% {n:5d} x  {name}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
