"""Torch-op restatement of the reference render path (TEST INFRASTRUCTURE ONLY).

The reference's renderer *is* a sequence of PyTorch ops (F.grid_sample + cumprod + elementwise,
gmpi/core/mpi.py:26-153,308-436).  The reference itself cannot travel to the GPU box
(/root/reference does not exist there), so this port re-expresses the same op sequence and is
what bench.py times on the box's host cores as the "reference's own PyTorch (grid_sample +
cumprod) path" (`cpu_baseline.kind = "port"`).  It is pinned to the live reference by
tests/test_oracle_golden.py (bit-identical on the golden fixtures, which were produced by the
unmodified reference in the build container with oracle/make_golden.py).

Never imported by the product package.
"""
from typing import List, Optional

import torch
import torch.nn.functional as F

NARROW = 0.95  # mpi.py:23


def warp_planes(tex, dhw, eye, rays, zdir, align_corners: bool = True):
    """Per (view, plane) sample of a plane texture along camera rays.  Mirrors homography(),
    gmpi/core/mpi.py:26-153.  tex [B,4,Ht,Wt], dhw [B,3], eye [B,3], rays [B,3,H,W], zdir [B,3]
    -> rgb [B,3,H,W], disparity [B,1,H,W], alpha [B,1,H,W]."""
    b, _, h, w = rays.shape
    with torch.no_grad():
        gap = (dhw[:, :1] - eye[:, 2:3]).view(b, 1, 1, 1).expand(b, 1, h, w)     # mpi.py:74-75
        t = gap / rays[:, 2:3]                                                   # mpi.py:76
        hit = eye.view(-1, 3, 1, 1) + rays * t                                   # mpi.py:79
        gv = 2 * hit[:, 1] / dhw[:, 1:2].unsqueeze(-1)                           # mpi.py:89
        gu = 2 * hit[:, 0] / dhw[:, 2:3].unsqueeze(-1)                           # mpi.py:90
        if not align_corners:                                                    # mpi.py:95-99
            mv = (gv >= -1) & (gv <= 1)
            gv[mv] = gv[mv] * NARROW
            mu = (gu >= -1) & (gu <= 1)
            gu[mu] = gu[mu] * NARROW
        grid = torch.stack([gu, gv], dim=-1)                                     # mpi.py:101
    out = F.grid_sample(tex, grid, mode="bilinear", padding_mode="zeros",
                        align_corners=align_corners)                             # mpi.py:136-142
    with torch.no_grad():
        zlen = torch.einsum("nchw,nc->nhw", rays, zdir)                          # mpi.py:149
        disparity = 1 / (t * zlen.view(b, 1, h, w))                              # mpi.py:150-151
    return out[:, :3], disparity, out[:, 3:4]


def render_views(batch_rgba: torch.Tensor, batch_dhw: torch.Tensor,
                 batch_ray_dir: List[torch.Tensor], batch_eye_pos: List[torch.Tensor],
                 batch_z_dir: List[torch.Tensor], align_corners: bool = True):
    """Mirrors MPI.forward, gmpi/core/mpi.py:331-436 (staging + compositing), same op order."""
    m_planes = batch_dhw.shape[1]
    per_view_rgba, per_view_dhw = [], []
    for k, r in enumerate(batch_ray_dir):                                        # mpi.py:334-343
        nv = r.shape[0]
        per_view_rgba.append(batch_rgba[k:k + 1].expand(nv, -1, -1, -1, -1))
        per_view_dhw.append(batch_dhw[k:k + 1].expand(nv, -1, -1))
    rgba = torch.cat(per_view_rgba, 0)
    dhw = torch.cat(per_view_dhw, 0)
    rays = torch.cat(batch_ray_dir, 0)                                           # mpi.py:350-354
    eyes = torch.cat(batch_eye_pos, 0)
    zdirs = torch.cat(batch_z_dir, 0)
    nv, _, ih, iw = rays.shape
    th, tw = rgba.shape[-2:]
    rays_f = rays.unsqueeze(1).expand(-1, m_planes, -1, -1, -1).reshape(nv * m_planes, 3, ih, iw)
    eyes_f = eyes.unsqueeze(1).expand(-1, m_planes, -1).reshape(nv * m_planes, 3)
    zdirs_f = zdirs.unsqueeze(1).expand(-1, m_planes, -1).reshape(nv * m_planes, 3)
    rgb, disparity, alpha = warp_planes(rgba.reshape(nv * m_planes, 4, th, tw),
                                        dhw.reshape(nv * m_planes, 3), eyes_f, rays_f, zdirs_f,
                                        align_corners)                           # mpi.py:399-409
    z = 1 / disparity                                                            # mpi.py:411
    alpha = alpha.reshape(nv, m_planes, 1, ih, iw)
    rgb = rgb.reshape(nv, m_planes, 3, ih, iw)
    z = z.reshape(nv, m_planes, 1, ih, iw)
    shifted = torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha + 1e-10], 1)   # mpi.py:421
    wgt = alpha * torch.cumprod(shifted, dim=1)[:, :-1]                          # mpi.py:423
    return torch.sum(wgt * rgb, dim=1), torch.sum(wgt * z, dim=1)                # mpi.py:430,434


def render(batch_rgba, dhw_table, ray_dirs, eyes, zdirs, align_corners=True,
           g_color: Optional[torch.Tensor] = None):
    """MPIRenderer.render's arithmetic (mpi_renderer.py:444-467) for one view per MPI:
    float(), range assert, MPI.forward, 2c-1.  If g_color is given, also backpropagates
    sum(img*g_color) to batch_rgba (the C3/C5 fwd+bwd workload)."""
    b = batch_rgba.shape[0]
    dhw = dhw_table.reshape(1, -1, 3).expand(b, -1, -1)
    batch_rgba = batch_rgba.float()
    assert torch.min(batch_rgba) >= 0.0 and torch.max(batch_rgba) <= 1.0        # mpi_renderer.py:447
    color, depth = render_views(batch_rgba, dhw, [r for r in ray_dirs], [e for e in eyes],
                                [z for z in zdirs], align_corners)
    img = 2 * color - 1                                                          # mpi_renderer.py:467
    if g_color is not None:
        (img * g_color).sum().backward()
    return img, depth
