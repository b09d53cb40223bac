#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of the shipped .so, the mnemonics that prove the sm_100a code paths (TMA loads,
mbarrier, packed f32x2 math, native integer shared atomics, vector global reductions, 128-bit stores) + short excerpts.
    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "ml_gmpi_b200", "libgmpi_mpi_render.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
print(f"cuobjdump -sass {os.path.relpath(so, ROOT)}   ({os.path.getsize(so)} bytes)\n")
KEYS = ["UTMALDG", "UTMAPF", "SYNCS", "FFMA2", "FMUL2", "FADD2", "LDS", "ATOMS.ADD", "ATOMS.CAST", "REDG.E.ADD.F32x4", "REDG.E.ADD.F32.",
        "STG.E.128", "DFMA", "NANOSLEEP", "BAR.SYNC", "STL", "LDL"]
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0]
    lines = [l for l in f.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
    c = collections.Counter()
    for l in lines:
        for k in KEYS:
            if k in l:
                c[k] += 1
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:110]
    print(f"{short}\n    {len(lines)} instructions; " + ", ".join(f"{k} {c[k]}" for k in KEYS if c[k]))
    if "fwd_staged_kernel<true, false, false>" in short or "bwd_box_kernel<true, false>" in short:
        seen = set()
        for l in lines:
            for k in ("UTMALDG", "ATOMS.ADD", "REDG.E.ADD.F32x4", "STG.E.128", "SYNCS.PHASECHK", "SYNCS.ARRIVE.TRANS64 RZ"):
                if k in l and k not in seen:
                    seen.add(k)
                    print("        " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l.strip()))
    print()
