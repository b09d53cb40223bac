#!/usr/bin/env python
"""A/B of forward kernel builds through the classic C entry point only (works with any ABI version of the library):
    python tools/fwd_ab.py <suffix> ...     # ml_gmpi_b200/libgmpi_mpi_render_<suffix>.so ("" = the in-tree build)
Headline workload (4 MPIs x 1 view, 96 planes, 1024^2); prints the mean kernel time of 30 launches after 10 warm-ups, twice."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ml_gmpi_b200 import synth

dev = torch.device("cuda:0")
case = synth.make_case(n_planes=96, tex=1024, img=1024, n_mpi=4, seed=1234, device=dev)
color = torch.empty((4, 3, 1024, 1024), device=dev)
depth = torch.empty((4, 1, 1024, 1024), device=dev)
flags = torch.zeros(1, dtype=torch.int32, device=dev)
vp, i, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32
import statistics
import time
libs = {}
for suf in sys.argv[1:]:
    name = "libgmpi_mpi_render" + ("_" + suf if suf and suf != "-" else "") + ".so"
    lib = ctypes.CDLL(os.path.join(ROOT, "ml_gmpi_b200", name))
    lib.gmpi_mpi_render_fwd.restype = i
    lib.gmpi_mpi_render_fwd.argtypes = [vp] * 9 + [i] * 7 + [u32, vp]
    libs[suf] = lib
st = torch.cuda.current_stream().cuda_stream


def run(lib):
    rc = lib.gmpi_mpi_render_fwd(case.rgba.data_ptr(), case.view2mpi.data_ptr(), case.dhw.data_ptr(), case.ray_dir.data_ptr(),
                                 case.eye.data_ptr(), case.z_dir.data_ptr(), color.data_ptr(), depth.data_ptr(), flags.data_ptr(),
                                 4, 4, 96, 1024, 1024, 1024, 1024, 1 | 2 | 4, st)
    assert rc == 0


# The boxes throttle within seconds of sustained load, so a build measured later looks slower: rotate the order every round,
# take the median over rounds, and report each build relative to the first one measured in the same round.
names = list(libs)
res = {n: [] for n in names}
rel = {n: [] for n in names}
for rnd in range(9):
    order = names[rnd % len(names):] + names[: rnd % len(names)]
    this = {}
    for n in order:
        for _ in range(5):
            run(libs[n])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(12):
            run(libs[n])
        e1.record()
        torch.cuda.synchronize()
        this[n] = e0.elapsed_time(e1) / 12
        res[n].append(this[n])
    for n in names:
        rel[n].append(this[n] / this[names[0]])
    time.sleep(0.3)
for n in names:
    print(f"{n or 'in-tree':12s} median {statistics.median(res[n]):.4f} ms  min {min(res[n]):.4f}  vs {names[0]}: median ratio {statistics.median(rel[n]):.4f}"
          f"  checksum {float(color.sum()):.3f}", flush=True)
