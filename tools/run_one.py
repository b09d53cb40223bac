#!/usr/bin/env python
"""One launch of a kernel for ncu / compute-sanitizer: python tools/run_one.py fwd|bwd|small [views]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ml_gmpi_b200 as g
from ml_gmpi_b200 import synth

what = sys.argv[1]
views = int(sys.argv[2]) if len(sys.argv) > 2 else (4 if what in ("fwd", "fwdfact") else 1)
dev = torch.device("cuda:0")
if what == "small":     # sanitizer-sized: staged kernels (>= 120 tiles), few planes
    case = synth.make_case(n_planes=6, tex=256, img=256, n_mpi=2, views_per_mpi=2, seed=7, device=dev, last_alpha_one=True)
else:
    case = synth.make_case(n_planes=96, tex=1024, img=1024, n_mpi=views, seed=1234, device=dev)
if what == "fwdfact":   # the headline shape from the FACTORED MPI (shared colour + per-plane alpha)
    gen = torch.Generator(device=dev).manual_seed(1)
    rgb = torch.rand((views, 3, 1024, 1024), generator=gen, device=dev)
    alpha = torch.rand((views, 96, 1, 1024, 1024), generator=gen, device=dev)
    for _ in range(2):
        c, d = g.render_views_factored(rgb, alpha, case.dhw, case.view2mpi, case.ray_dir, case.eye, case.z_dir, check_last_plane=True,
                                       color_minus1_1=True)
    torch.cuda.synchronize()
    print("ok", what, views)
    sys.exit(0)
if what in ("c4", "c4flat"):   # video config: 30 views of ONE 96-plane 512^2 MPI, with / without the view-grouped tile order
    import numpy as np
    nv = 30
    vc = synth.make_case(n_planes=96, tex=512, img=512, n_mpi=1, views_per_mpi=nv, seed=1234, device=dev,
                         yaws=np.linspace(0.5, -0.5, 120).astype(np.float32)[:nv], pitches=np.zeros(nv, np.float32))
    for _ in range(2):
        c, d = g.render_views(vc.rgba, vc.dhw, vc.view2mpi, vc.ray_dir, vc.eye, vc.z_dir, check_last_plane=True, color_minus1_1=True,
                              view_group=nv if what == "c4" else 1)
    torch.cuda.synchronize()
    print("ok", what, nv)
    sys.exit(0)
if what == "fwd":
    for _ in range(2):
        c, d = g.render_views(case.rgba, case.dhw, case.view2mpi, case.ray_dir, case.eye, case.z_dir, check_last_plane=True,
                              color_minus1_1=True)
else:
    rgba = case.rgba.requires_grad_(True)
    for _ in range(2):
        rgba.grad = None
        c, d = g.render_views(rgba, case.dhw, case.view2mpi, case.ray_dir, case.eye, case.z_dir, color_minus1_1=True)
        (c.sum() + d.sum()).backward()
torch.cuda.synchronize()
print("ok", what, views)
