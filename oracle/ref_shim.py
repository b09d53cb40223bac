"""Import the UNMODIFIED reference renderer from /root/reference (build container only).

TEST INFRASTRUCTURE.  Only `oracle/make_golden.py` and tests that are skipped when
/root/reference is absent may use this.  Nothing on the product path imports it and nothing
on the GPU box can (the reference is not there).

The reference needs two pure-Python packages that are not installed here and cannot be
fetched (no network): `lazy` (gmpi/core/camera.py:10) and `yacs` (gmpi/utils/config.py:6,
pulled in by gmpi/utils/__init__.py:2).  Both are import-time only for the render path:
`lazy.lazy` is a memoising property (functools.cached_property has the same semantics) and
`yacs.config.CfgNode` is never instantiated by gmpi.core.*.  Neither touches arithmetic.
"""
import functools
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GMPI_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "gmpi", "core", "mpi.py"))


def _install_shims() -> None:
    if "lazy" not in sys.modules:
        m = types.ModuleType("lazy")
        m.lazy = functools.cached_property
        sys.modules["lazy"] = m
    if "yacs" not in sys.modules:
        class CfgNode(dict):
            def __init__(self, init_dict=None, key_list=None, new_allowed=False):
                super().__init__(init_dict or {})

            def clone(self):
                return CfgNode(dict(self))

            def freeze(self):
                pass

            def defrost(self):
                pass

        y = types.ModuleType("yacs")
        yc = types.ModuleType("yacs.config")
        yc.CfgNode = CfgNode
        y.config = yc
        sys.modules["yacs"] = y
        sys.modules["yacs.config"] = yc


def import_reference():
    """Returns (gmpi.core.mpi, gmpi.core.mpi_renderer) modules of the untouched reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import gmpi.core.mpi as ref_mpi  # noqa: E402
    import gmpi.core.mpi_renderer as ref_renderer  # noqa: E402

    return ref_mpi, ref_renderer


FFHQ_KWARGS = dict(  # train.py:277-298 with curriculums.py:109-116 and configs/gmpi.yml:74-96
    plane_min_d=0.95,
    plane_max_d=1.12,
    plan_spatial_enlarge_factor=1.001,
    plane_distances_sample_method="inverse",
    cam_fov=12.6,
    sphere_center_z=1.0,
    sphere_r=1.0,
    horizontal_mean=0.0,
    horizontal_std=0.289,
    vertical_mean=0.0,
    vertical_std=0.127,
    cam_pose_n_truncated_stds=2,
    cam_sample_method="truncated_gaussian",
    mpi_align_corners=True,
    use_confined_volume=True,
)
