#!/usr/bin/env python
"""Where a kernel's warp-stall samples fall, from `ncu -i X.ncu-rep --page source --csv --print-source sass` (one result):
per block of consecutive SASS instructions the share of samples, executed warp instructions and the opcodes that hold the
samples; then the stall reasons overall.  Measurement aid.

    ncu -i rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > /tmp/k.csv; python tools/ncu_hot.py /tmp/k.csv [block]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    h = next(i for i, r in enumerate(rows) if "# Samples" in r)
    hdr = rows[h]
    data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
    iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    tot = sum(int(r[iS]) for r in data)
    print(rows[0][1][:100] if rows and len(rows[0]) > 1 else "")
    print("total samples", tot, "| SASS instructions", len(data), "| warp instructions executed", sum(int(r[iI]) for r in data))

    def opname(s):
        t = s.split()
        return (t[1] if t[0].startswith("@") else t[0]) if t else "?"

    for b in range(0, len(data), B):
        seg = data[b:b + B]
        s = sum(int(r[iS]) for r in seg)
        ie = sum(int(r[iI]) for r in seg)
        ops = {}
        for r in seg:
            ops[opname(r[iSrc])] = ops.get(opname(r[iSrc]), 0) + int(r[iS])
        top = sorted(ops.items(), key=lambda x: -x[1])[:5]
        print(f"{b:5d}-{b + B:5d} samples {s:6d} ({100 * s / max(tot, 1):4.1f}%) exec {ie:8d}  {top}")
    stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    agg = {c: sum(int(r[hdr.index(c)]) for r in data) for c in stalls}
    print(sorted(agg.items(), key=lambda x: -x[1])[:8])


if __name__ == "__main__":
    main()
