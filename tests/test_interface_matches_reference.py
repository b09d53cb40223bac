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
