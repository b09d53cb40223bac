"""One-off plane geometry of an MPI: plane distances and metric plane sizes (the `dhw` table).

Mirrors: sample_distance gmpi/utils/mpi_utils.py:21-53,
compute_plane_dhws_given_cam_pose_spatial_range[_confined] :652-917 and
MPIRenderer.compute_mpi_spatial_volume gmpi/core/mpi_renderer.py:105-152.  The reference walks 10 001
poses in a Python loop (7.7 s); here the pose envelope is evaluated in one vectorised batch (~10 ms).
"""
import numpy as np
import torch

from .camera import PinholeCamera, sphere_poses


def sample_distance(dmin: float, dmax: float, n: int, method: str = "inverse") -> np.ndarray:
    assert 0 < dmin <= dmax and 1 <= n < 9999
    if method == "uniform":
        d = np.linspace(dmin, dmax, n)
    elif method == "log-uniform":
        d = np.exp(np.linspace(np.log(dmin), np.log(dmax), n))
    elif method == "sqrt":
        d = np.linspace(dmin ** 0.5, dmax ** 0.5, n) ** 2
    elif method == "squared":
        d = np.sqrt(np.linspace(dmin ** 2, dmax ** 2, n))
    elif method == "inverse":                                   # uniform in disparity, near -> far
        d = (1.0 / np.linspace(1.0 / dmax, 1.0 / dmin, n))[::-1]
    else:
        raise ValueError(method)
    return np.asarray(d, dtype=np.float32)


def plane_dhw_table(*, n_planes, plane_min_d, plane_max_d, enlarge_factor, distance_method, fov_deg, sphere_center,
                    sphere_r, h_mean, h_std, v_mean, v_std, n_truncated_stds, confined=True, grid=100) -> np.ndarray:
    """-> [n_planes, 3] fp32 (distance, metric height, metric width), planes near -> far."""
    ds = np.clip(sample_distance(plane_min_d, plane_max_d, n_planes, distance_method), plane_min_d, plane_max_d).astype(np.float32)
    far = np.float32(ds[-1])
    h0, h1 = h_mean - n_truncated_stds * h_std, h_mean + n_truncated_stds * h_std
    v0, v1 = v_mean - n_truncated_stds * v_std, v_mean + n_truncated_stds * v_std
    yy, pp = np.meshgrid(np.linspace(h0, h1, grid), np.linspace(v0, v1, grid), indexing="ij")
    yaws = np.concatenate([yy.reshape(-1), [(h0 + h1) / 2]])        # the mid pose is appended last (mpi_utils.py:833)
    pitches = np.concatenate([pp.reshape(-1), [(v0 + v1) / 2]])
    c2w = sphere_poses(torch.from_numpy(yaws), torch.from_numpy(pitches), sphere_center, sphere_r).numpy().astype(np.float64)
    cam = PinholeCamera.from_fov(fov_deg, 4, 4)                      # only ray directions matter (mpi_renderer.py:128)
    rays = np.matmul(c2w[:, :3, :3], cam.border_dirs64()).astype(np.float32)          # [P,3,4]
    eye = c2w[:, :3, 3].astype(np.float32)
    t = (far - eye[:, 2:3]) / rays[:, 2]                                               # mpi_utils.py:632-636
    x = eye[:, 0:1] + rays[:, 0] * t
    y = eye[:, 1:2] + rays[:, 1] * t
    bound = max(np.abs(x).max(), np.abs(y).max())
    assert bound <= 5.0, (f"You have MPI's plane whose boundary value is up to {bound}. This usually means the camera "
                          f"poses's range is too big, which will cause problems for MPI representation.")   # mpi_utils.py:888-894
    mid_h = 2 * np.abs(y[-1]).max()
    mid_w = 2 * np.abs(x[-1]).max()
    rows = []
    for d in ds[:-1]:
        s = 1.0 if confined else float(d) / float(far)                                 # mpi_utils.py:775-779 vs :904-907
        rows.append([d, mid_h * s, mid_w * s])
    rows.append([far, 2 * np.abs(y).max() * enlarge_factor, 2 * np.abs(x).max() * enlarge_factor])
    return np.asarray(rows, dtype=np.float32)


FFHQ = dict(  # curriculums.py:109-116, configs/gmpi.yml:74-96
    plane_min_d=0.95, plane_max_d=1.12, enlarge_factor=1.001, distance_method="inverse", fov_deg=12.6,
    sphere_center=(0.0, 0.0, 1.0), sphere_r=1.0, h_mean=0.0, h_std=0.289, v_mean=0.0, v_std=0.127, n_truncated_stds=2,
    confined=True,
)
