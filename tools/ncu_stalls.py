#!/usr/bin/env python
"""Aggregate ncu source-page stall samples by opcode and by stall reason: python tools/ncu_stalls.py rep.ncu-rep"""
import collections
import csv
import re
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_op = collections.Counter(); by_reason = collections.Counter(); ex_by_op = collections.Counter()
by_op_reason = collections.defaultdict(collections.Counter)
conf = collections.Counter()
total = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ci["Source"]].strip()
    t = src.split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
    op = re.sub(r"\..*", "", op)
    s = int(r[ci["# Samples"]] or 0)
    by_op[op] += s; total += s
    ex_by_op[op] += int(r[ci["Instructions Executed"]] or 0)
    for c in stall_cols:
        v = int(r[ci[c]] or 0)
        by_reason[c] += v; by_op_reason[op][c] += v
    if op == "LDS":
        conf["ideal"] += int(r[ci["L1 Wavefronts Shared Ideal"]] or 0); conf["actual"] += int(r[ci["L1 Wavefronts Shared"]] or 0)
print("total samples", total)
print("by reason:", [(k, v) for k, v in by_reason.most_common(10)])
print("%-12s %10s %8s %14s  top reasons" % ("opcode", "samples", "share", "warp-instrs"))
for op, s in by_op.most_common(16):
    top = ", ".join(f"{k[6:]}={v}" for k, v in by_op_reason[op].most_common(3))
    print("%-12s %10d %7.1f%% %14d  %s" % (op, s, 100.0 * s / max(total, 1), ex_by_op[op], top))
print("LDS wavefronts ideal/actual:", conf["ideal"], conf["actual"])
print("warp instructions total:", sum(ex_by_op.values()))
print("mix:", [(k, v) for k, v in ex_by_op.most_common(14)])
