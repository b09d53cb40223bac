"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

The reference ships no tests/golden vectors for the render path, so the
oracle and the CUDA kernels are pinned to outputs of the reference itself, produced here with
torch CPU (2.11.0) and committed as small fixtures.  Each fixture stores the exact input
tensors (or the seed that regenerates them bit-identically with numpy) and the reference's
outputs: MPI.forward colour/depth (gmpi/core/mpi.py:308-436) and, through torch autograd,
d(sum(color*Gc)+sum(depth*Gd))/d rgba.

TEST INFRASTRUCTURE ONLY.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def make_renderer(ref_r, n_planes, **over):
    kw = dict(ref_shim.FFHQ_KWARGS)
    kw.update(over)
    import logging
    logging.disable(logging.CRITICAL)
    with quiet(), contextlib.redirect_stderr(io.StringIO()):
        r = ref_r.MPIRenderer(n_mpi_planes=n_planes, device=torch.device("cpu"), **kw)
    return r


def cams(renderer, res, yaws, pitches):
    with quiet():
        renderer.set_cam(renderer.cam_fov, res, res)
        y = torch.tensor(yaws, dtype=torch.float32).view(-1, 1)
        p = torch.tensor(pitches, dtype=torch.float32).view(-1, 1)
        infos = renderer.sample_cam_poses(len(yaws), 0, 0, 0, 0, random_pose=True, given_yaws=y, given_pitches=p)
    keys = ["batch_yaws", "batch_pitches", "batch_tf_c2w", "batch_ray_dir", "batch_eye_pos", "batch_z_dir"]
    return dict(zip(keys, infos))


def rand_rgba(seed, shape):
    # numpy Generator(PCG64) output is stable across numpy versions for .random(dtype=float32)
    return np.random.default_rng(seed).random(shape, dtype=np.float32)


def run_mpi(ref_mpi, rgba, dhw, groups_ray, groups_eye, groups_z, align_corners, check_last, gseed, with_gd=True):
    """groups_*: list (one entry per MPI) of tensors [n_i,3,H,W] / [n_i,3]."""
    mod = ref_mpi.MPI(align_corners=align_corners)
    t_rgba = torch.from_numpy(rgba).clone().requires_grad_(True)
    color, depth = mod(batch_rgba=t_rgba, batch_dhw=torch.from_numpy(dhw), batch_ray_dir=groups_ray,
                       batch_eye_pos=groups_eye, batch_z_dir=groups_z, separate_background=None,
                       assert_not_out_of_last_plane=check_last)
    rng = np.random.default_rng(gseed)
    gc = rng.standard_normal(color.shape).astype(np.float32)
    gd = rng.standard_normal(depth.shape).astype(np.float32) if with_gd else None
    loss = (color * torch.from_numpy(gc)).sum()
    if with_gd:
        loss = loss + (depth * torch.from_numpy(gd)).sum()
    loss.backward()
    out = dict(color=color.detach().numpy(), depth=depth.detach().numpy(), g_color=gc,
               g_rgba=t_rgba.grad.numpy())
    if with_gd:
        out["g_depth"] = gd
    with torch.no_grad():
        c_over, d_over = mod.old_forward(batch_rgba=torch.from_numpy(rgba), batch_dhw=torch.from_numpy(dhw),
                                         batch_ray_dir=groups_ray, batch_eye_pos=groups_eye,
                                         batch_z_dir=groups_z, separate_background=None,
                                         assert_not_out_of_last_plane=False)
    out["color_over"] = c_over.numpy()
    out["depth_over"] = d_over.numpy()
    return out


def pack_views(groups_ray, groups_eye, groups_z):
    v2m = np.concatenate([np.full(r.shape[0], k, np.int32) for k, r in enumerate(groups_ray)])
    return dict(view2mpi=v2m, ray_dir=torch.cat(groups_ray).numpy(), eye=torch.cat(groups_eye).numpy(),
                z_dir=torch.cat(groups_z).numpy())


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref_mpi, ref_r = ref_shim.import_reference()

    # ---------------------------------------------------------------- geometry tables (A10)
    r8 = make_renderer(ref_r, 8)
    r32 = make_renderer(ref_r, 32)
    r96 = make_renderer(ref_r, 96)
    r8u = make_renderer(ref_r, 8, use_confined_volume=False)
    save("ffhq_dhw", n8=r8.static_mpi_plane_dhws.numpy(), n32=r32.static_mpi_plane_dhws.numpy(),
         n96=r96.static_mpi_plane_dhws.numpy(), n8_unconfined=r8u.static_mpi_plane_dhws.numpy())

    # ---------------------------------------------------------------- camera / pose fixtures (A2-A4)
    yaws = [0.0, 0.3, -0.2, 0.5, -0.5, 0.11]
    pitches = [0.0, 0.1, 0.05, -0.2, 0.2, -0.07]
    ci = cams(r8, 20, yaws, pitches)
    save("ffhq_cams_20", yaws=np.array(yaws, np.float32), pitches=np.array(pitches, np.float32),
         c2w=ci["batch_tf_c2w"].numpy(), ray_dir=torch.cat(ci["batch_ray_dir"]).numpy(),
         eye=torch.cat(ci["batch_eye_pos"]).numpy(), z_dir=torch.cat(ci["batch_z_dir"]).numpy(),
         fov=np.float32(12.6), sphere_center=np.array([0, 0, 1.0]), sphere_r=np.float32(1.0))

    # ---------------------------------------------------------------- tiny: 2 MPIs, 3 views, img != tex
    dhw8 = r8.static_mpi_plane_dhws.numpy()
    ci = cams(r8, 12, [0.3, -0.2, 0.11], [0.1, 0.05, -0.07])
    g_ray = [torch.cat(ci["batch_ray_dir"][:2]), ci["batch_ray_dir"][2]]
    g_eye = [torch.cat(ci["batch_eye_pos"][:2]), ci["batch_eye_pos"][2]]
    g_z = [torch.cat(ci["batch_z_dir"][:2]), ci["batch_z_dir"][2]]
    rgba = rand_rgba(11, (2, 8, 4, 16, 16))
    dhw = np.broadcast_to(dhw8[None], (2, 8, 3)).copy()
    for ac in (True, False):
        out = run_mpi(ref_mpi, rgba, dhw, g_ray, g_eye, g_z, ac, False, 21)
        save("tiny_2mpi_3view" + ("" if ac else "_acfalse"), rgba=rgba, dhw=dhw, align_corners=np.int32(ac),
             **pack_views(g_ray, g_eye, g_z), **out)

    # ---------------------------------------------------------------- alpha == 1 planes (production MPI has
    # alpha==1 on the last plane, networks_cond_on_pos_enc.py:1307-1310; sanity mode sets every alpha to 1)
    ci = cams(r8, 24, [0.25, -0.4], [0.12, -0.1])
    g_ray, g_eye, g_z = ci["batch_ray_dir"], ci["batch_eye_pos"], ci["batch_z_dir"]
    rgba = rand_rgba(12, (2, 8, 4, 24, 24))
    rgba[:, -1, 3] = 1.0
    rgba[0, 3, 3, 6:18, 4:20] = 1.0
    rgba[1, 2, 3, :, :12] = 0.0
    rgba[1, 5, 3] = 1.0
    out = run_mpi(ref_mpi, rgba, dhw, g_ray, g_eye, g_z, True, True, 22)
    save("alpha_one_planes", rgba=rgba, dhw=dhw, align_corners=np.int32(1), **pack_views(g_ray, g_eye, g_z), **out)
    rgba_s = rand_rgba(13, (1, 8, 4, 24, 24))
    rgba_s[:, :, 3] = 1.0   # eval/prepare_fake_data.py:51-56 "stylegan2_sanity_check"
    out = run_mpi(ref_mpi, rgba_s, dhw[:1], g_ray[:1], g_eye[:1], g_z[:1], True, True, 23)
    save("sanity_all_alpha_one", rgba=rgba_s, dhw=dhw[:1], align_corners=np.int32(1),
         **pack_views(g_ray[:1], g_eye[:1], g_z[:1]), **out)

    # ---------------------------------------------------------------- rays leaving the planes (zero padding)
    ci = cams(r8, 20, [0.9, -1.1], [0.5, -0.45])   # far outside the 2-sigma pose envelope
    g_ray, g_eye, g_z = ci["batch_ray_dir"], ci["batch_eye_pos"], ci["batch_z_dir"]
    rgba = rand_rgba(14, (2, 8, 4, 16, 16))
    out = run_mpi(ref_mpi, rgba, dhw, g_ray, g_eye, g_z, True, False, 24)
    save("out_of_plane", rgba=rgba, dhw=dhw, align_corners=np.int32(1), **pack_views(g_ray, g_eye, g_z), **out)

    # ---------------------------------------------------------------- minification / magnification
    ci = cams(r8, 40, [0.2], [-0.1])
    rgba = rand_rgba(15, (1, 8, 4, 16, 16))
    out = run_mpi(ref_mpi, rgba, dhw[:1], ci["batch_ray_dir"], ci["batch_eye_pos"], ci["batch_z_dir"], True, True, 25)
    save("magnify_40_from_16", rgba=rgba, dhw=dhw[:1], align_corners=np.int32(1),
         **pack_views(ci["batch_ray_dir"], ci["batch_eye_pos"], ci["batch_z_dir"]), **out)
    ci = cams(r8, 10, [-0.3], [0.15])
    rgba = rand_rgba(16, (1, 8, 4, 64, 64))
    out = run_mpi(ref_mpi, rgba, dhw[:1], ci["batch_ray_dir"], ci["batch_eye_pos"], ci["batch_z_dir"], True, True, 26)
    save("minify_10_from_64", rgba=rgba, dhw=dhw[:1], align_corners=np.int32(1),
         **pack_views(ci["batch_ray_dir"], ci["batch_eye_pos"], ci["batch_z_dir"]), **out)

    # ---------------------------------------------------------------- non-square texture / image (MPI.forward
    # itself does not require squares; only MPIRenderer does, mpi_renderer.py:84,155)
    ci = cams(r8, 16, [0.1], [0.0])
    ray = ci["batch_ray_dir"][0][:, :, 2:14, :].contiguous()   # [1,3,12,16]
    rgba = rand_rgba(17, (1, 8, 4, 20, 28))
    out = run_mpi(ref_mpi, rgba, dhw[:1], [ray], ci["batch_eye_pos"], ci["batch_z_dir"], True, False, 27)
    save("nonsquare", rgba=rgba, dhw=dhw[:1], align_corners=np.int32(1),
         **pack_views([ray], ci["batch_eye_pos"], ci["batch_z_dir"]), **out)

    # ---------------------------------------------------------------- C1 reduced and C1 full
    # BASELINE.json configs[0]: single 256x256 view, 32 planes, random RGBA, identity pose.
    dhw32 = r32.static_mpi_plane_dhws.numpy()[None].copy()
    for res, tag, store_rgba in ((64, "c1_small_64", True), (256, "c1_full_256", False)):
        ci = cams(r32, res, [0.0], [0.0])
        rgba = rand_rgba(1234, (1, 32, 4, res, res))
        out = run_mpi(ref_mpi, rgba, dhw32, ci["batch_ray_dir"], ci["batch_eye_pos"], ci["batch_z_dir"], True, True,
                      28, with_gd=False)
        # the renderer-level output (img in [-1,1]) through MPIRenderer.render, mpi_renderer.py:387-469
        with quiet(), torch.no_grad():
            img, dep, c2w, ang = r32.render(torch.from_numpy(rgba), res, res, given_cam_infos=ci)
        extra = dict(render_img=img.numpy(), render_depth=dep.numpy(), render_c2w=c2w.numpy(), render_angles=ang.numpy())
        if not store_rgba:
            out.pop("g_rgba")   # 33 MB at 256^2; the small case keeps the reference gradient
        extra["rgba_seed"] = np.int64(1234)   # inputs regenerate bit-identically from the seed
        extra["rgba_shape"] = np.array(rgba.shape)
        save(tag, dhw=dhw32, align_corners=np.int32(1),
             **pack_views(ci["batch_ray_dir"], ci["batch_eye_pos"], ci["batch_z_dir"]), **out, **extra)

    # ---------------------------------------------------------------- C2 reduced: 4 MPIs x 1 view, 32 planes, 64^2
    rng = np.random.default_rng(1234)
    yw = rng.uniform(-0.5, 0.5, 4).astype(np.float32).tolist()
    pt = rng.uniform(-0.2, 0.2, 4).astype(np.float32).tolist()
    ci = cams(r32, 64, yw, pt)
    rgba = rand_rgba(1235, (4, 32, 4, 64, 64))
    rgba[:, -1, 3] = 1.0
    dhw = np.broadcast_to(dhw32, (4, 32, 3)).copy()
    out = run_mpi(ref_mpi, rgba, dhw, ci["batch_ray_dir"], ci["batch_eye_pos"], ci["batch_z_dir"], True, True, 29)
    out.pop("g_rgba")
    g = out  # keep g_color/g_depth so the gradient can be checked against the oracle instead
    save("c2_small_4x32x64", rgba_seed=np.int64(1235), rgba_shape=np.array(rgba.shape), last_alpha_one=np.int32(1),
         dhw=dhw, align_corners=np.int32(1), yaws=np.array(yw, np.float32), pitches=np.array(pt, np.float32),
         **pack_views(ci["batch_ray_dir"], ci["batch_eye_pos"], ci["batch_z_dir"]), **g)


if __name__ == "__main__":
    main()
