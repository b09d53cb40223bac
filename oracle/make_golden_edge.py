"""Edge-case fixtures for the ORACLE only (tests/test_oracle_golden.py), produced by the UNMODIFIED reference on CPU:

    python oracle/make_golden_edge.py       # needs /root/reference; writes tests/golden/edge_*.npz

  edge_single_plane        N = 1 (the composite degenerates to alpha_0 * rgb_0)
  edge_ragged_zero_views   three MPIs rendered from 2, 0 and 1 views: an MPI without views contributes nothing and gets a zero gradient
  edge_acfalse_nonsquare   align_corners=False (0.95 shrink + the other unnormalisation, mpi.py:23,95-99) on a non-square texture
  edge_odd_sizes           37x37 image from a 17x19 texture (nothing a multiple of anything)

TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from make_golden import cams, make_renderer, pack_views, rand_rgba, run_mpi, save  # noqa: E402


def main():
    torch.manual_seed(0)
    ref_mpi, ref_r = ref_shim.import_reference()
    r8 = make_renderer(ref_r, 8)
    dhw8 = r8.static_mpi_plane_dhws.numpy()

    ci = cams(r8, 14, [0.2, -0.35], [0.05, -0.12])
    g_ray, g_eye, g_z = [torch.cat(ci["batch_ray_dir"])], [torch.cat(ci["batch_eye_pos"])], [torch.cat(ci["batch_z_dir"])]
    rgba = rand_rgba(31, (1, 1, 4, 12, 12))
    dhw = dhw8[None, -1:].copy()                                  # the far (largest) plane alone
    out = run_mpi(ref_mpi, rgba, dhw, g_ray, g_eye, g_z, True, True, 41)
    save("edge_single_plane", rgba=rgba, dhw=dhw, align_corners=np.int32(1), **pack_views(g_ray, g_eye, g_z), **out)

    ci = cams(r8, 12, [0.3, -0.2, 0.11], [0.1, 0.05, -0.07])
    empty = lambda t: t[:0]
    g_ray = [torch.cat(ci["batch_ray_dir"][:2]), empty(ci["batch_ray_dir"][0]), ci["batch_ray_dir"][2]]
    g_eye = [torch.cat(ci["batch_eye_pos"][:2]), empty(ci["batch_eye_pos"][0]), ci["batch_eye_pos"][2]]
    g_z = [torch.cat(ci["batch_z_dir"][:2]), empty(ci["batch_z_dir"][0]), ci["batch_z_dir"][2]]
    rgba = rand_rgba(32, (3, 8, 4, 16, 16))
    dhw = np.broadcast_to(dhw8[None], (3, 8, 3)).copy()
    out = run_mpi(ref_mpi, rgba, dhw, g_ray, g_eye, g_z, True, False, 42)
    assert not out["g_rgba"][1].any() and out["color"].shape[0] == 3
    save("edge_ragged_zero_views", rgba=rgba, dhw=dhw, align_corners=np.int32(1), **pack_views(g_ray, g_eye, g_z), **out)

    ci = cams(r8, 16, [0.15], [-0.05])
    ray = ci["batch_ray_dir"][0][:, :, 3:13, :].contiguous()      # [1,3,10,16]
    rgba = rand_rgba(33, (1, 8, 4, 18, 26))
    out = run_mpi(ref_mpi, rgba, dhw[:1], [ray], ci["batch_eye_pos"], ci["batch_z_dir"], False, False, 43)
    save("edge_acfalse_nonsquare", rgba=rgba, dhw=dhw[:1], align_corners=np.int32(0),
         **pack_views([ray], ci["batch_eye_pos"], ci["batch_z_dir"]), **out)

    ci = cams(r8, 37, [-0.25], [0.1])
    rgba = rand_rgba(34, (1, 8, 4, 17, 19))
    out = run_mpi(ref_mpi, rgba, dhw[:1], ci["batch_ray_dir"], ci["batch_eye_pos"], ci["batch_z_dir"], True, True, 44)
    save("edge_odd_sizes", rgba=rgba, dhw=dhw[:1], align_corners=np.int32(1),
         **pack_views(ci["batch_ray_dir"], ci["batch_eye_pos"], ci["batch_z_dir"]), **out)


if __name__ == "__main__":
    main()
