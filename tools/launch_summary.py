#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised times:
compare SHARES). usage: launch_summary.py LAUNCHES.csv [LAUNCHES2.csv ...] -> markdown on stdout"""
import csv
import sys
from collections import OrderedDict


def load(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        rows.append((r[ik], us))
    return rows


def short(name):
    n = name.replace("void ", "").replace("dnaldpc::", "")
    return n.split("(")[0][:70]


def main():
    for path in sys.argv[1:]:
        rows = load(path)
        agg = OrderedDict()
        for k, us in rows:
            a = agg.setdefault(short(k), [0, 0.0, 0.0])
            a[0] += 1; a[1] += us; a[2] = max(a[2], us)
        tot = sum(a[1] for a in agg.values())
        print("## %s (%d launches, %.1f ms)\n" % (path.split("/")[-1], len(rows), tot / 1e3))
        print("| kernel | launches | total us | avg us | max us | share |\n|---|---|---|---|---|---|")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("| %s | %d | %.0f | %.1f | %.1f | %.3f |" % (k, a[0], a[1], a[1] / a[0], a[2], a[1] / tot))
        print()


if __name__ == "__main__":
    main()
