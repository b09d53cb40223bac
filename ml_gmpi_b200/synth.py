"""Synthetic MPI workloads with the reference's FFHQ geometry: random RGBA
stacks, in-envelope poses, rays from the pinhole camera.  Used by bench.py, smoke() and the
full-size GPU tests; needs neither the reference nor the oracle."""
from dataclasses import dataclass

import numpy as np
import torch

from .camera import PinholeCamera, sphere_poses
from .geometry import FFHQ, plane_dhw_table

_DHW_CACHE = {}


def ffhq_dhw(n_planes: int) -> torch.Tensor:
    if n_planes not in _DHW_CACHE:
        _DHW_CACHE[n_planes] = torch.from_numpy(plane_dhw_table(n_planes=n_planes, **FFHQ))
    return _DHW_CACHE[n_planes]


@dataclass
class Case:
    rgba: torch.Tensor       # [M,N,4,T,T]
    dhw: torch.Tensor        # [M,N,3]
    view2mpi: torch.Tensor   # [V] int32
    ray_dir: torch.Tensor    # [V,3,H,W]
    eye: torch.Tensor        # [V,3]
    z_dir: torch.Tensor      # [V,3]
    c2w: torch.Tensor        # [V,4,4]
    yaws: torch.Tensor
    pitches: torch.Tensor

    def to(self, device, pin=False):
        f = (lambda t: t.pin_memory()) if pin else (lambda t: t.to(device))
        return Case(*[f(getattr(self, k)) for k in self.__dataclass_fields__])


def make_poses(n_views, img, seed=1234, yaws=None, pitches=None, device="cpu"):
    if yaws is None:   # U(-0.5,0.5) x U(-0.2,0.2): inside the 2-sigma envelope
        rng = np.random.default_rng(seed)
        yaws = rng.uniform(-0.5, 0.5, n_views).astype(np.float32)
        pitches = rng.uniform(-0.2, 0.2, n_views).astype(np.float32)
    yaws, pitches = torch.as_tensor(yaws, dtype=torch.float32), torch.as_tensor(pitches, dtype=torch.float32)
    c2w = sphere_poses(yaws, pitches, FFHQ["sphere_center"], FFHQ["sphere_r"]).to(device)
    cam = PinholeCamera.from_fov(FFHQ["fov_deg"], img, img)
    ray_dir, eye, z_dir = cam.generate_rays(c2w)
    return ray_dir, eye, z_dir, c2w, yaws, pitches


def make_case(*, n_planes, tex, img, n_mpi, views_per_mpi=1, seed=1234, device="cpu", last_alpha_one=False,
              yaws=None, pitches=None, rgba=True) -> Case:
    V = n_mpi * views_per_mpi
    ray_dir, eye, z_dir, c2w, yaws, pitches = make_poses(V, img, seed, yaws, pitches, device)
    gen = torch.Generator(device=device).manual_seed(seed)
    t = None
    if rgba:
        t = torch.rand((n_mpi, n_planes, 4, tex, tex), generator=gen, device=device, dtype=torch.float32)
        if last_alpha_one:
            t[:, -1, 3] = 1.0      # production MPIs: networks_cond_on_pos_enc.py:1307-1310
    dhw = ffhq_dhw(n_planes).to(device).unsqueeze(0).expand(n_mpi, -1, -1).contiguous()
    v2m = torch.arange(n_mpi, dtype=torch.int32, device=device).repeat_interleave(views_per_mpi)
    return Case(t, dhw, v2m, ray_dir, eye, z_dir, c2w, yaws, pitches)
