"""Golden fixture for LightRenderer (gmpi/core/light_renderer.py), produced by the UNMODIFIED reference on CPU:

    python oracle/make_golden_light.py      # needs /root/reference; writes tests/golden/light_2x6x32.npz

The reference samples the light inside `render` (gen_sphere_path, light_renderer.py:136-149) and does not expose it; a
RECORDING wrapper around the function object in the module namespace captures the angles it drew so that the mirror can be
driven with the same light.  The wrapper only observes: no arithmetic of the reference changes.

TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    ref_shim.import_reference()
    import gmpi.core.light_renderer as ref_light
    rec = {}
    orig = ref_light.gen_sphere_path

    def recording(**kw):
        out = orig(**kw)
        rec["c2w"], rec["yaws"], rec["pitches"] = (np.asarray(o.cpu() if isinstance(o, torch.Tensor) else o) for o in out)
        return out
    ref_light.gen_sphere_path = recording

    B, N, H, W = 2, 6, 32, 32
    rng = np.random.default_rng(77)
    mpi = rng.random((B, N, 4, H, W), dtype=np.float32)
    mpi[:, -1, 3] = 1.0
    d = (1.0 / np.linspace(1.0 / 1.12, 1.0 / 0.95, N))[::-1].astype(np.float32)
    dhw = np.stack([d, np.full(N, 0.2473, np.float32), np.full(N, 0.2473, np.float32)], 1).astype(np.float32)
    # texel positions of every plane, [N,H,W,4] (x, y, z, 1), +X right +Y down (mpi_renderer.get_xyz convention)
    xs = (np.arange(W, dtype=np.float32) + 0.5) / W - 0.5
    ys = (np.arange(H, dtype=np.float32) + 0.5) / H - 0.5
    xyz = np.zeros((N, H, W, 4), np.float32)
    for i in range(N):
        xyz[i, :, :, 0] = xs[None, :] * dhw[i, 2]
        xyz[i, :, :, 1] = ys[:, None] * dhw[i, 1]
        xyz[i, :, :, 2] = dhw[i, 0]
        xyz[i, :, :, 3] = 1.0
    lr = ref_light.LightRenderer(sphere_center_z=1.0, sphere_r=1.0, ka_max=0.7, kd_max=0.6, n_grow_iters=10)
    lr.step = 20                                    # fully grown: ka = 0.7, kd = 0.6
    torch.manual_seed(5)
    np.random.seed(5)
    t_mpi = torch.from_numpy(mpi).requires_grad_(True)
    t_dhw, t_xyz = torch.from_numpy(dhw), torch.from_numpy(xyz)
    depth = lr.compute_depth(t_mpi[:, :, 3:], t_dhw[:, :1])
    lr.step = 20
    out = lr.render(t_mpi, t_dhw, t_xyz)
    g_out = torch.from_numpy(np.random.default_rng(78).standard_normal(out.shape).astype(np.float32))
    (out * g_out).sum().backward()
    g_depth = torch.from_numpy(np.random.default_rng(79).standard_normal(depth.shape).astype(np.float32))
    t_alpha = torch.from_numpy(mpi[:, :, 3:]).clone().requires_grad_(True)
    (lr.compute_depth(t_alpha, t_dhw[:, :1]) * g_depth).sum().backward()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "light_2x6x32.npz"), mpi=mpi, dhw=dhw, xyz=xyz, depth=depth.detach().numpy(),
                        out=out.detach().numpy(), g_out=g_out.numpy(), g_mpi=t_mpi.grad.numpy(), g_depth=g_depth.numpy(),
                        g_alpha_depth=t_alpha.grad.numpy(), light_yaws=rec["yaws"].astype(np.float32),
                        light_pitches=rec["pitches"].astype(np.float32), ka=np.float32(lr.cur_ka), kd=np.float32(lr.cur_kd))
    print("wrote light_2x6x32.npz", depth.shape, out.shape, rec["yaws"].ravel(), rec["pitches"].ravel(), lr.cur_ka, lr.cur_kd)


if __name__ == "__main__":
    main()
