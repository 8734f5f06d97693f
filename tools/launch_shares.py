#!/usr/bin/env python
"""Per-kernel totals and shares from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <command>`).
Usage: python tools/launch_shares.py X.csv "<command line that was profiled>" > X_shares.txt"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    cmd = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = re.sub(r"\(.*\)$", "", r[ki])
        name = re.sub(r"\b(jsdr|void)\b:*\s*", "", name).replace("(int)", "")
        t = tot.setdefault(name, [0, 0.0, 0.0])
        t[0] += 1
        t[1] += v
        t[2] = max(t[2], v)
    all_ms = sum(t[1] for t in tot.values())
    if cmd:
        print(cmd)
    print("(cold-cache, serialised launches: compare SHARES, not absolutes; the scout normally runs beside the data kernels)")
    print("(max = the largest launch: the whole-batch launch of the timed step where smaller launches of the e2e / variant legs share the name)")
    for name, (n, ms, mx) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:70]:70s} launches {n:4d}  total {ms:9.3f} ms  avg {ms / n:8.3f} ms  max {mx:8.3f} ms  share {100 * ms / all_ms:5.1f}%")


if __name__ == "__main__":
    main()
