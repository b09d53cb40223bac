"""Drop-in boundary: the host-side mirrors must accept exactly what the reference's callers pass.
Runs only where /root/reference exists (the build container); the GPU box skips it."""
import inspect

import pytest

import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference not mounted")


def _params(fn):
    return [(n, p.kind, p.default) for n, p in inspect.signature(fn).parameters.items() if n != "self"]


def test_mpi_signatures():
    ref_mpi, _ = ref_shim.import_reference()
    import ml_gmpi_b200 as g
    ours, theirs = _params(g.MPI.forward), _params(ref_mpi.MPI.forward)
    assert ours == theirs                                            # keyword-only, same names, same defaults
    assert _params(g.MPI.check_shapes) == _params(ref_mpi.MPI.check_shapes)
    init_ref = _params(ref_mpi.MPI.__init__)
    init_ours = _params(g.MPI.__init__)
    assert init_ours[: len(init_ref)] == init_ref                    # ours adds the optional `validate` after align_corners


def test_renderer_signatures():
    _, ref_r = ref_shim.import_reference()
    from ml_gmpi_b200.renderer import MPIRenderer
    for name in ("render", "sample_cam_poses", "set_cam", "compute_mpi_spatial_volume"):
        assert _params(getattr(MPIRenderer, name)) == _params(getattr(ref_r.MPIRenderer, name)), name
    ref_init = _params(ref_r.MPIRenderer.__init__)
    ours_init = _params(MPIRenderer.__init__)
    assert ours_init[: len(ref_init)] == ref_init                    # ours adds the optional `validate`


def test_reference_call_site_binds():
    """mpi_renderer.py:451-461 calls self.mpi(batch_rgba=..., ..., c2w_mat=..., sphere_c=...): must bind to our forward."""
    import ml_gmpi_b200 as g
    sig = inspect.signature(g.MPI.forward)
    sig.bind(None, batch_rgba=1, batch_dhw=2, batch_ray_dir=3, batch_eye_pos=4, batch_z_dir=5, separate_background=None,
             assert_not_out_of_last_plane=True, c2w_mat=6, sphere_c=7)


def test_unmodified_reference_renderer_drives_the_drop_in():
    """INTEGRATION.md's patch, executed: `gmpi.core.mpi_renderer.MPI = ml_gmpi_b200.MPI`, then the UNMODIFIED reference
    MPIRenderer is constructed (mpi_renderer.py:47 instantiates our class) and `render` is called with real tensors.  The call
    must get through the reference's own pose sampling / ray generation (mpi_renderer.py:418-449) and our check_shapes with the
    reference's real argument list (mpi_renderer.py:451-461), and stop exactly at the "CUDA devices only" check -- there is no
    GPU in the build container and no CPU fallback by design."""
    import torch
    import ml_gmpi_b200 as g
    _, ref_r = ref_shim.import_reference()
    old = ref_r.MPI
    ref_r.MPI = g.MPI
    try:
        r = ref_r.MPIRenderer(n_mpi_planes=4, plane_min_d=0.95, plane_max_d=1.12, plan_spatial_enlarge_factor=1.001,
                              plane_distances_sample_method="inverse", cam_fov=12.6, sphere_center_z=1.0, sphere_r=1.0,
                              horizontal_mean=0.0, horizontal_std=0.289, vertical_mean=0.0, vertical_std=0.127,
                              cam_pose_n_truncated_stds=2, cam_sample_method="truncated_gaussian", mpi_align_corners=True,
                              use_confined_volume=True, device=torch.device("cpu"))
        assert isinstance(r.mpi, g.MPI) and r.mpi._align_corners is True
        rgba = torch.rand(2, 4, 4, 16, 16)
        with pytest.raises(RuntimeError, match="CUDA devices only"):
            r.render(rgba, 16, 16, given_yaws=torch.zeros(2, 1), given_pitches=torch.zeros(2, 1))
        # a malformed MPI is rejected by OUR check_shapes with the reference's message before any device work
        with pytest.raises(AssertionError, match="Expected rgba to be of shape"):
            r.mpi(batch_rgba=torch.rand(2, 4, 3, 16, 16), batch_dhw=torch.rand(2, 4, 3), batch_ray_dir=[torch.rand(1, 3, 8, 8)] * 2,
                  batch_eye_pos=[torch.rand(1, 3)] * 2, batch_z_dir=[torch.rand(1, 3)] * 2, separate_background=None)
    finally:
        ref_r.MPI = old
