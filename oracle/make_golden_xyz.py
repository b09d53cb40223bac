"""Golden fixture for the façade's generator-side helpers (MPIRenderer.get_xyz*, gmpi/core/mpi_renderer.py:154-318),
produced by the UNMODIFIED reference on CPU:

    python oracle/make_golden_xyz.py        # needs /root/reference; writes tests/golden/ffhq_xyz.npz

TEST INFRASTRUCTURE ONLY.
"""
import logging
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    logging.disable(logging.CRITICAL)
    _, ref_r = ref_shim.import_reference()
    out = {}
    for rng_name in ("-11", "01"):
        r = ref_r.MPIRenderer(n_mpi_planes=8, device=torch.device("cpu"), use_normalized_xyz=True, normalized_xyz_range=rng_name,
                              **ref_shim.FFHQ_KWARGS)
        xyz, nxyz = r.get_xyz(16, 16, ret_single_res=True)
        out[f"xyz16_{rng_name}"], out[f"nxyz16_{rng_name}"] = xyz.numpy().copy(), nxyz.numpy().copy()
        out[f"xyzd16_{rng_name}"] = r.mpi_tex_pix_3d_coords.numpy().copy()
        z, nz = r.get_xyz(16, 16, ret_single_res=True, only_z=True)
        out[f"z_{rng_name}"], out[f"nz_{rng_name}"] = z.numpy().copy(), nz.numpy().copy()
        xd, nd = r.get_xyz(16, 16, ret_single_res=False)
        assert sorted(xd) == [4, 8, 16]
        for res in xd:
            out[f"multi_xyz{res}_{rng_name}"], out[f"multi_nxyz{res}_{rng_name}"] = xd[res].numpy().copy(), nd[res].numpy().copy()
    r = ref_r.MPIRenderer(n_mpi_planes=8, device=torch.device("cpu"), use_xyz_ztype="disparity", **ref_shim.FFHQ_KWARGS)
    xd, nd = r.get_xyz(8, 8, ret_single_res=False)
    assert nd[8] is None
    out["disp_xyz4"], out["disp_xyz8"] = xd[4].numpy().copy(), xd[8].numpy().copy()
    for (s, t) in ((8, 12), (32, 96), (8, 8), (96, 32)):
        out[f"interp_{s}_{t}"] = r.get_xyz_interpolate_ws(s, t).numpy().copy()
    cam_r = ref_r.MPIRenderer(n_mpi_planes=4, device=torch.device("cpu"), **ref_shim.FFHQ_KWARGS)
    cam_r.set_cam(12.6, 12, 12)
    c2w = np.array([[0.96, 0.0, -0.28, 0.28], [0.0, 1.0, 0.0, 0.0], [0.28, 0.0, 0.96, 0.04], [0, 0, 0, 1]], np.float32)
    ray, eye, z_dir, tf = cam_r.view_info_from_c2w_mat(cam_r.cam, c2w)
    out["vi_c2w"], out["vi_ray"], out["vi_eye"], out["vi_z"], out["vi_tf"] = c2w, ray.numpy(), eye.numpy(), z_dir.numpy(), tf.numpy()
    # seeded RANDOM poses: the mirror must consume torch's global RNG exactly as the reference does (cam_utils.py:510-555,
    # torch_utils.py:51-76), so that a seeded training / FID run draws the same cameras
    for method in ("truncated_gaussian", "uniform", "normal"):
        kw = dict(ref_shim.FFHQ_KWARGS, cam_sample_method=method)
        rr = ref_r.MPIRenderer(n_mpi_planes=4, device=torch.device("cpu"), **kw)
        rr.set_cam(12.6, 8, 8)
        torch.manual_seed(3)
        y, p, c2w, *_ = rr.sample_cam_poses(5, 0.0, 0.289, 0.0, 0.127, True)
        y2, p2, _c, *_ = rr.sample_cam_poses(3, 0.1, 0.2, -0.05, 0.1, True)          # second draw from the same stream
        out[f"rand_{method}_yaw"], out[f"rand_{method}_pitch"] = y.numpy().copy(), p.numpy().copy()
        out[f"rand_{method}_c2w"] = np.asarray(c2w.numpy() if isinstance(c2w, torch.Tensor) else c2w, np.float32)
        out[f"rand_{method}_yaw2"], out[f"rand_{method}_pitch2"] = y2.numpy().copy(), p2.numpy().copy()
    y, p, c2w, *_ = rr.sample_cam_poses(5, 0.1, 0.289, 0.05, 0.127, False)           # deterministic horizontal sweep
    out["sweep_yaw"], out["sweep_pitch"] = y.numpy().copy(), p.numpy().copy()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "ffhq_xyz.npz"), **out)
    print("wrote ffhq_xyz.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
