#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: opcode mix (executed warp instructions),
stall-reason shares, and the hottest SASS lines.  Usage: ncu_source_summary.py source.csv [top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {k: i for i, k in enumerate(hdr)}
ops = collections.defaultdict(lambda: [0, 0, 0])
stall = collections.Counter()
tot_exec = tot_samp = 0
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
lines = []


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src
    base = op.split(".")[0]
    ex, sm = num(r[ix["Instructions Executed"]]), num(r[ix["# Samples"]])
    ops[base][0] += 1
    ops[base][1] += ex
    ops[base][2] += sm
    tot_exec += ex
    tot_samp += sm
    for c in stall_cols:
        stall[c] += num(r[ix[c]])
    lines.append((sm, ex, src))
print("static instrs", len(lines), "executed warp-instrs", tot_exec, "samples", tot_samp)
for k, v in sorted(ops.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k:12s} static {v[0]:5d} exec {v[1]:10d} {100 * v[1] / max(1, tot_exec):5.1f}%  samples {100 * v[2] / max(1, tot_samp):5.1f}%")
print()
tot_st = sum(stall.values())
for k, v in stall.most_common(12):
    print(f"{k:28s} {v:8d} {100 * v / max(1, tot_st):5.1f}%")
print()
for sm, ex, src in sorted(lines, reverse=True)[:top]:
    print(f"{sm:6d} samples  exec {ex:8d}  {src}")
