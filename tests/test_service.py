"""Render service harness (ml_gmpi_b200/service.py) host logic on CPU: view sharding and ordering of the video sweep across
2 gloo ranks, the reference's strided FID image numbering, the sweep angles -- with the renderer replaced by a stub."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ml_gmpi_b200 import service


def test_sweep_angles_match_render_video():
    a = service.sweep_angles(100, True, 0.1)
    assert len(a) == 100 and abs(a[0] - 0.6) < 1e-6 and abs(a[-1] + 0.4) < 1e-6          # render_video.py:236-237
    b = service.sweep_angles(100, False)
    assert abs(b[0] - 0.3) < 1e-6 and abs(b[-1] + 0.3) < 1e-6                            # render_video.py:239-240


def test_fid_indices_are_the_reference_strided_numbering():
    for world in (1, 2, 3, 8):
        parts = [service.fid_image_indices(50, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(50))
        assert all(p == list(range(r, 50, world)) for r, p in enumerate(parts))           # fid_evaluation.py:86,129-133


def _fake_video(rgba, dhw, c2w, img_size, fov, near, far, fast, factored):
    # a frame whose every pixel encodes the camera's x position (i.e. the yaw): order and sharding become checkable
    V = c2w.shape[0]
    code = ((c2w[:, 0, 3] + 1.0) * 100).round().to(torch.uint8)
    img = code.view(V, 1, 1, 1).expand(V, img_size, img_size, 3).contiguous()
    return img, img[..., :1].contiguous() + 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        angles = service.sweep_angles(7, True)
        kw = dict(img_size=4, fov_deg=12.6, ray_start=0.95, ray_end=1.12, sphere_center=np.array([0, 0, 1.0]), sphere_r=1.0)
        img, depth = service.render_video_frames(None, torch.zeros(1, 2, 3), angles, rank=rank, world=world, render_fn=_fake_video, **kw)
        mine, _ = service.render_video_frames(None, torch.zeros(1, 2, 3), angles, rank=rank, world=world, gather=False,
                                              render_fn=_fake_video, **kw)
        q.put((rank, img[:, 0, 0, 0].tolist(), depth[:, 0, 0, 0].tolist(), mine.shape[0]))
    finally:
        dist.destroy_process_group()


def test_video_frames_are_sharded_and_gathered_in_view_order():
    kw = dict(img_size=4, fov_deg=12.6, ray_start=0.95, ray_end=1.12, sphere_center=np.array([0, 0, 1.0]), sphere_r=1.0)
    angles = service.sweep_angles(7, True)
    single, sd = service.render_video_frames(None, torch.zeros(1, 2, 3), angles, render_fn=_fake_video, **kw)
    codes = single[:, 0, 0, 0].tolist()
    assert codes == sorted(codes, reverse=True) and len(set(codes)) == 7                  # yaw 0.5 -> -0.5: x = sin(yaw) decreasing
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, c, dd, n_mine in res:
        assert c == codes and dd == [x + 1 for x in codes]                                 # every rank holds all 7 frames, in order
        assert n_mine == (4 if rank == 0 else 3)                                           # shard_range(7, r, 2)


def test_dump_fid_images_numbers_and_counts():
    calls = []

    class R:
        sphere_center, sphere_r, cam_fov = np.array([0, 0, 1.0]), 1.0, 12.6

    def fake_render(renderer, batch, img_size, yaws, pitches):
        assert yaws.shape == (batch.shape[0], 1) and float(yaws.abs().max()) <= 2 * 0.289 + 1e-6      # truncated at 2 sigma
        return torch.full((batch.shape[0], img_size, img_size, 3), len(calls), dtype=torch.uint8)

    written = []
    done = service.dump_fid_images(R(), lambda k: (calls.append(k), torch.zeros(3, 2, 4, 8, 8))[1], num_imgs=11, rank=1, world=4,
                                   img_size=8, writer=lambda i, im: written.append((i, im.shape)), render_fn=fake_render)
    assert done == [1, 5, 9] and [w[0] for w in written] == [1, 5, 9] and written[0][1] == (8, 8, 3)
    assert calls == [0]                                                                    # one batch of 3 covered the 3 images


def test_uint8_truncating_conversion_equals_the_reference_numpy_lines():
    """prepare_fake_data.py:72-74 / render_video.py:119-121: img = np.clip((img + 1) / 2.0, 0.0, 1.0); (img * 255).astype(np.uint8)."""
    x = torch.rand(3, 5, 7, 3, generator=torch.Generator().manual_seed(2)) * 2.4 - 1.2       # also outside [-1, 1]
    x[0, 0, 0] = torch.tensor([-1.0, 1.0, 0.0])
    want = (np.clip((x.numpy() + 1) / 2.0, 0.0, 1.0) * 255).astype(np.uint8)
    assert np.array_equal(service.to_uint8_truncating(x).numpy(), want)
    assert want[0, 0, 0].tolist() == [0, 255, 127]


def test_render_eval_views_orders_views_mpi_major_and_draws_the_renderers_poses():
    from ml_gmpi_b200.renderer import MPIRenderer
    r = MPIRenderer(n_mpi_planes=4, plane_min_d=0.95, plane_max_d=1.12, plan_spatial_enlarge_factor=1.001,
                    plane_distances_sample_method="inverse", cam_fov=12.6, sphere_center_z=1.0, sphere_r=1.0, horizontal_mean=0.0,
                    horizontal_std=0.289, vertical_mean=0.0, vertical_std=0.127, cam_pose_n_truncated_stds=2,
                    cam_sample_method="truncated_gaussian", use_confined_volume=True)
    seen = {}

    def fake(renderer, batch, n_imgs, img_size, yaws, pitches):
        seen["yaws"], seen["pitches"] = yaws.clone(), pitches.clone()
        V = batch.shape[0] * n_imgs
        v = torch.arange(V, dtype=torch.float32).view(V, 1, 1, 1)
        return (v / V * 2 - 1).expand(V, 3, img_size, img_size), (v + 0.5).expand(V, 1, img_size, img_size)

    torch.manual_seed(11)
    img, depth, angles = service.render_eval_views(r, torch.zeros(2, 4, 4, 8, 8), n_imgs=3, img_size=6, render_fn=fake)
    assert img.shape == (6, 6, 6, 3) and img.dtype == np.uint8 and depth.shape == (6, 6, 6, 1) and depth.dtype == np.float32
    assert img[:, 0, 0, 0].tolist() == [int(np.float32(v / 6) * 255) for v in range(6)] and depth[:, 0, 0, 0].tolist() == [v + 0.5 for v in range(6)]
    # the poses are what MPIRenderer.render would draw for a batch of 6 from the same generator state, angles = (pitch, yaw)
    torch.manual_seed(11)
    r.set_cam(12.6, 6, 6)
    y, p, *_ = r.sample_cam_poses(6, r.horizontal_mean, r.horizontal_std, r.vertical_mean, r.vertical_std, True)
    assert torch.equal(seen["yaws"], y) and torch.equal(seen["pitches"], p)
    assert np.array_equal(angles, torch.cat([p, y], -1).numpy())
    # the default renderer is the CUDA path: no CPU fallback
    import pytest
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        service.render_eval_views(r, torch.zeros(1, 4, 4, 8, 8), n_imgs=2, img_size=6)


def test_default_eval_render_calls_render_frames_with_a_valid_argument_list(monkeypatch):
    """The default (CUDA) path cannot run here; bind its call against render_frames' real signature and check the tensors."""
    import inspect
    from ml_gmpi_b200 import mpi as mpi_mod
    from ml_gmpi_b200.renderer import MPIRenderer
    real = mpi_mod.render_frames
    got = {}

    def checker(**kw):
        inspect.signature(real).bind(**kw)
        got.update(kw)
        V, H, W = kw["ray_dir"].shape[0], kw["ray_dir"].shape[2], kw["ray_dir"].shape[3]
        return torch.zeros(V, 3, H, W), torch.ones(V, 1, H, W)

    monkeypatch.setattr(mpi_mod, "render_frames", checker)
    r = MPIRenderer(n_mpi_planes=4, plane_min_d=0.95, plane_max_d=1.12, plan_spatial_enlarge_factor=1.001,
                    plane_distances_sample_method="inverse", cam_fov=12.6, sphere_center_z=1.0, sphere_r=1.0, horizontal_mean=0.0,
                    horizontal_std=0.289, vertical_mean=0.0, vertical_std=0.127, cam_pose_n_truncated_stds=2,
                    cam_sample_method="truncated_gaussian", use_confined_volume=True)
    img, depth, angles = service.render_eval_views(r, torch.zeros(2, 4, 4, 8, 8), n_imgs=3, img_size=6)
    assert got["view2mpi"].tolist() == [0, 0, 0, 1, 1, 1] and got["view2mpi"].dtype == torch.int32 and got["view_group"] == 3
    assert got["dhw"].shape == (2, 4, 3) and got["ray_dir"].shape == (6, 3, 6, 6) and got["eye"].shape == (6, 3) and got["z_dir"].shape == (6, 3)
    assert got["check_last_plane"] is True and got.get("video") is None
    assert img.shape == (6, 6, 6, 3) and int(img[0, 0, 0, 0]) == 127 and float(depth.min()) == 1.0 and angles.shape == (6, 2)
