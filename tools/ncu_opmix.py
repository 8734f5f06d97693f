#!/usr/bin/env python3
"""Opcode mix and stall share of one kernel from an ncu source page.

    ncu --set full --import-source on -k regex:<kernel> -c 1 -o rep <command>
    ncu -i rep.ncu-rep --page source --csv > src.csv
    python tools/ncu_opmix.py src.csv [samples_per_launch]

Prints executed warp-instructions per opcode (and, with the number of samples the launch
processed, thread-instructions per sample) beside the share of warp-stall samples."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nsamp = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[1]
i_src, i_ex = hdr.index("Source"), hdr.index("Instructions Executed")
i_thr = hdr.index("Thread Instructions Executed")
i_st = hdr.index("Warp Stall Sampling (All Samples)")
ops, thr, st = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= i_thr:
        continue
    tok = r[i_src].split()
    if not tok:
        continue
    op = (tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]).split(".")[0]
    ops[op] += int(r[i_ex] or 0)
    thr[op] += int(r[i_thr] or 0)
    st[op] += int(r[i_st] or 0)
tot, tthr, tst = sum(ops.values()), sum(thr.values()), max(sum(st.values()), 1)
print(f"# {rows[0][1][:90]}")
print(f"# warp-instructions {tot}, thread-instructions {tthr}" + (f" = {tthr / nsamp:.1f} per sample" if nsamp else ""))
print(f"{'opcode':10s} {'warp-instr':>12s} {'share':>6s} " + (f"{'thr/sample':>10s} " if nsamp else "") + f"{'stalls':>6s}")
for op, n in ops.most_common(24):
    per = f"{thr[op] / nsamp:10.2f} " if nsamp else ""
    print(f"{op:10s} {n:12d} {100 * n / tot:5.1f}% {per}{100 * st[op] / tst:5.1f}%")
