#!/usr/bin/env python
"""Train-step time (4 MPIs x 1 view, 96 planes, 1024^2) with the gradient zeroed by memsets vs inside the backward kernel.
A/B helper: GMPI_LIB_PATH picks the build, GMPI_ZERO_FRAC the pacing.   python tools/zero_ab.py"""
import os
import sys

import torch

sys.path.insert(0, ".")
import ml_gmpi_b200 as g                     # noqa: E402
from ml_gmpi_b200 import _lib, synth         # noqa: E402

dev = torch.device("cuda:0")
case = synth.make_case(device=dev, n_mpi=4, views_per_mpi=1, n_planes=96, tex=1024, img=1024, seed=0)
rg = case.rgba.requires_grad_(True)
gcol = torch.randn((4, 3, 1024, 1024), device=dev)
lib = _lib.load()


def step():
    rg.grad = None
    c = g.render_views(rg, case.dhw, case.view2mpi, case.ray_dir, case.eye, case.z_dir, color_minus1_1=True)[0]
    (c * gcol).sum().backward()


def timed(n=6, warm=2):
    ts = []
    for i in range(warm + n):
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        torch.cuda.synchronize(dev)
        if i >= warm:
            ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


out = []
for mode in (0, 1, 0, 1):
    lib.gmpi_debug_set_bwd_zero(mode)
    out.append(timed())
print(os.path.basename(os.environ.get("GMPI_LIB_PATH", "default")), "frac", os.environ.get("GMPI_ZERO_FRAC", "-"),
      "memsets %.3f %.3f ms   in-kernel %.3f %.3f ms" % (out[0], out[2], out[1], out[3]))
