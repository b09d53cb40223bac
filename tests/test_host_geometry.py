"""Host-side camera / pose / plane-geometry code against fixtures produced by
the unmodified reference (tests/golden/ffhq_dhw.npz, ffhq_cams_20.npz)."""
import numpy as np
import pytest
import torch

from ml_gmpi_b200 import camera, geometry, synth
from conftest import load_golden


def test_plane_table_matches_reference():
    ref = load_golden("ffhq_dhw")
    for n, key in ((8, "n8"), (32, "n32"), (96, "n96")):
        ours = geometry.plane_dhw_table(n_planes=n, **geometry.FFHQ)
        assert ours.shape == ref[key].shape and ours.dtype == np.float32
        np.testing.assert_allclose(ours, ref[key], rtol=2e-6, atol=0)
    kw = dict(geometry.FFHQ); kw["confined"] = False
    np.testing.assert_allclose(geometry.plane_dhw_table(n_planes=8, **kw), ref["n8_unconfined"], rtol=2e-6)


def test_sample_distance_methods():
    d = geometry.sample_distance(0.95, 1.12, 32, "inverse")
    assert d[0] == np.float32(0.95) and abs(d[-1] - 1.12) < 1e-6 and np.all(np.diff(d) > 0)
    assert np.allclose(np.diff(1.0 / d.astype(np.float64)), np.diff(1.0 / d.astype(np.float64))[0], rtol=1e-4)
    for m in ("uniform", "log-uniform", "sqrt", "squared"):
        x = geometry.sample_distance(0.5, 2.0, 5, m)
        assert abs(x[0] - 0.5) < 1e-6 and abs(x[-1] - 2.0) < 1e-6


def test_poses_and_rays_match_reference():
    ref = load_golden("ffhq_cams_20")
    c2w = camera.sphere_poses(torch.from_numpy(ref["yaws"]), torch.from_numpy(ref["pitches"]), ref["sphere_center"],
                              float(ref["sphere_r"]))
    np.testing.assert_allclose(c2w.numpy(), ref["c2w"], atol=2e-7)
    cam = camera.PinholeCamera.from_fov(float(ref["fov"]), 20, 20)
    ray, eye, z = cam.generate_rays(torch.from_numpy(ref["c2w"]))
    np.testing.assert_allclose(ray.numpy(), ref["ray_dir"], atol=2e-7)
    assert np.array_equal(eye.numpy(), ref["eye"]) and np.array_equal(z.numpy(), ref["z_dir"])


def test_identity_pose_is_identity_matrix():
    c2w = camera.sphere_poses(torch.zeros(1), torch.zeros(1), (0, 0, 1.0), 1.0)
    np.testing.assert_allclose(c2w[0].numpy(), np.eye(4), atol=1e-7)


def test_truncated_normal_stays_in_range():
    g = torch.Generator().manual_seed(0)
    x = camera.truncated_normal(10000, 0.0, 0.289, 2, g)
    assert x.shape == (10000, 1) and float(x.abs().max()) <= 2 * 0.289 + 1e-6


def test_synth_case_shapes():
    c = synth.make_case(n_planes=8, tex=16, img=12, n_mpi=2, views_per_mpi=3, seed=1)
    assert c.rgba.shape == (2, 8, 4, 16, 16) and c.ray_dir.shape == (6, 3, 12, 12)
    assert c.view2mpi.tolist() == [0, 0, 0, 1, 1, 1] and c.dhw.shape == (2, 8, 3)
    assert float(c.rgba.min()) >= 0 and float(c.rgba.max()) <= 1


def test_renderer_facade_constructor_matches_reference_table():
    from ml_gmpi_b200.renderer import MPIRenderer
    r = MPIRenderer(n_mpi_planes=8, plane_min_d=0.95, plane_max_d=1.12, plan_spatial_enlarge_factor=1.001,
                    plane_distances_sample_method="inverse", cam_fov=12.6, sphere_center_z=1.0, sphere_r=1.0,
                    horizontal_mean=0.0, horizontal_std=0.289, vertical_mean=0.0, vertical_std=0.127,
                    cam_pose_n_truncated_stds=2, cam_sample_method="truncated_gaussian", use_confined_volume=True)
    np.testing.assert_allclose(r.static_mpi_plane_dhws.numpy(), load_golden("ffhq_dhw")["n8"], rtol=2e-6)
    r.set_cam(12.6, 20, 20)
    ref = load_golden("ffhq_cams_20")
    infos = r.sample_cam_poses(6, 0, 0, 0, 0, True, torch.from_numpy(ref["yaws"]).view(-1, 1), torch.from_numpy(ref["pitches"]).view(-1, 1))
    np.testing.assert_allclose(torch.cat(infos[3]).numpy(), ref["ray_dir"], atol=3e-7)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        r.render(torch.rand(1, 8, 4, 16, 16), 20, 20)


def test_general_camera_matches_reference_fixture_and_pinhole():
    """`Camera` (the reference's general-K class, camera.py:13-211) against the reference fixture and against the batched
    PinholeCamera the render path uses: same fp32 rays bit for bit."""
    ref = load_golden("ffhq_cams_20")
    f = camera.focal_from_fov(float(ref["fov"]), 20)
    cam = camera.gen_cam(h=20, w=20, f=f, ray_from_pix_center=True)
    pin = camera.PinholeCamera(20, 20, f)
    assert cam.height == 20 and cam.width == 20 and cam.intrinsic_matrix[0, 2] == 10 and "Camera: height=20" in repr(cam)
    assert torch.equal(cam.ray_dir_torch, pin.cam_dirs("cpu"))
    np.testing.assert_allclose(cam.ray_dir_border_np, pin.border_dirs64(), atol=1e-15)
    c2w = torch.from_numpy(ref["c2w"])
    batched = pin.generate_rays(c2w)
    for v in range(c2w.shape[0]):
        ray, eye, z = cam.generate_rays(c2w[v])
        assert ray.shape == (3, 20, 20) and torch.equal(eye, batched[1][v]) and torch.equal(z, batched[2][v])
        np.testing.assert_allclose(ray.numpy(), ref["ray_dir"][v], atol=2e-7)
        ray64, eye64, _ = cam.generate_rays(ref["c2w"][v].astype(np.float64))
        assert ray64.dtype == np.float64 and np.abs(ray64 - ray.numpy()).max() < 2e-7
    assert cam.generate_rays(ref["c2w"][0], border_only=True)[0].shape == (3, 2, 2)
    with pytest.raises(ValueError):
        cam.generate_rays([[1.0]])
    with pytest.raises(AssertionError, match="Expecting a 3x3 intrinsics"):
        camera.Camera(4, 4, np.eye(4))
    # skew and an off-centre principal point: K [x y 1]^T lands back on the pixel
    K = np.array([[310.5, 0.7, 63.2], [0, 295.1, 40.9], [0, 0, 1.0]])
    g = camera.Camera(80, 120, K, ray_from_pix_center=True)
    uv = K @ g.homogeneous_coordinates.reshape(3, -1)
    np.testing.assert_allclose(uv[0].reshape(80, 120)[5, 7], 7.5, atol=1e-9)
    np.testing.assert_allclose(uv[1].reshape(80, 120)[5, 7], 5.5, atol=1e-9)
