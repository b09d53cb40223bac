"""Drop-in for `gmpi.core.mpi.MPI` (reference gmpi/core/mpi.py:156-436), backed by the sm_100a
kernels through the C ABI.  Same constructor, same keyword-only `forward`, same return values,
same assertion messages; differentiable w.r.t. `batch_rgba` (first order), which is all the
reference's callers need (the sampling grid and the depth are built under no_grad,
mpi.py:65,148).

No CPU path: tensors must live on a CUDA device, otherwise this raises.
"""
import ctypes
import warnings
from typing import List, Optional, Union

import numpy as np
import torch
from torch import nn

from . import _lib


class MPIOutOfPlaneError(AssertionError):
    """Rays leave the last plane (reference: prints the poses and sys.exit(1), mpi.py:103-128)."""


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _as_f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class _RenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgba, dhw, view2mpi, ray_dir, eye, z_dir, options, flags):
        lib = _lib.load()
        M, N, _, Ht, Wt = rgba.shape
        V, _, H, W = ray_dir.shape
        color = torch.empty((V, 3, H, W), device=rgba.device, dtype=torch.float32)
        depth = torch.empty((V, 1, H, W), device=rgba.device, dtype=torch.float32)
        # training: the forward also saves the transmittance in front of every plane (4 B per pixel-plane) so that the
        # backward is ONE staged back-to-front sweep (torch autograd keeps ~30 such tensors alive for the reference)
        trans = None
        if ctx.needs_input_grad[0]:
            trans = torch.empty((V, N, H, W), device=rgba.device, dtype=torch.float32)
        with torch.cuda.device(rgba.device):
            if trans is not None:
                _lib.check(lib.gmpi_mpi_render_fwd_train(
                    rgba.data_ptr(), view2mpi.data_ptr(), dhw.data_ptr(), ray_dir.data_ptr(), eye.data_ptr(), z_dir.data_ptr(),
                    color.data_ptr(), depth.data_ptr(), trans.data_ptr(), flags.data_ptr(),
                    M, V, N, Ht, Wt, H, W, options, _stream_ptr(rgba.device)))
            else:
                _lib.check(lib.gmpi_mpi_render_fwd(
                    rgba.data_ptr(), view2mpi.data_ptr(), dhw.data_ptr(), ray_dir.data_ptr(), eye.data_ptr(),
                    z_dir.data_ptr(), color.data_ptr(), depth.data_ptr(), flags.data_ptr(),
                    M, V, N, Ht, Wt, H, W, options, _stream_ptr(rgba.device)))
        ctx.save_for_backward(rgba, dhw, view2mpi, ray_dir, eye, z_dir, trans)
        ctx.options = options
        ctx.set_materialize_grads(False)
        return color, depth

    @staticmethod
    @torch.autograd.function.once_differentiable     # raw kernels: a double backward (create_graph=True) must raise, not
    def backward(ctx, g_color, g_depth):             # silently treat g_rgba as constant (the reference's R1 only differentiates D)
        rgba, dhw, view2mpi, ray_dir, eye, z_dir, trans = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return (None,) * 8
        lib = _lib.load()
        M, N, _, Ht, Wt = rgba.shape
        V, _, H, W = ray_dir.shape
        if g_color is None:
            g_color = torch.zeros((V, 3, H, W), device=rgba.device, dtype=torch.float32)
        g_color = _as_f32c(g_color)
        gd_ptr = None
        if g_depth is not None:
            g_depth = _as_f32c(g_depth)
            gd_ptr = g_depth.data_ptr()
        g_rgba = torch.empty_like(rgba)
        with torch.cuda.device(rgba.device):   # autograd worker threads do not inherit the device
            if trans is not None:
                _lib.check(lib.gmpi_mpi_render_bwd_saved(
                    rgba.data_ptr(), view2mpi.data_ptr(), dhw.data_ptr(), ray_dir.data_ptr(), eye.data_ptr(), z_dir.data_ptr(),
                    trans.data_ptr(), g_color.data_ptr(), gd_ptr, g_rgba.data_ptr(),
                    M, V, N, Ht, Wt, H, W, ctx.options | _lib.OPT_ZERO_GRAD, _stream_ptr(rgba.device)))
            else:
                _lib.check(lib.gmpi_mpi_render_bwd(
                    rgba.data_ptr(), view2mpi.data_ptr(), dhw.data_ptr(), ray_dir.data_ptr(), eye.data_ptr(),
                    z_dir.data_ptr(), g_color.data_ptr(), gd_ptr, g_rgba.data_ptr(),
                    M, V, N, Ht, Wt, H, W, ctx.options | _lib.OPT_ZERO_GRAD, _stream_ptr(rgba.device)))
        return g_rgba, None, None, None, None, None, None, None


_warned_direct = set()


def _warn_if_direct(rgba, V, H, W):
    """Surface the direct-kernel performance cliff (several times slower than the TMA-staged kernels) once per shape."""
    M, N, _, Ht, Wt = rgba.shape
    key = (V, N, Ht, Wt, H, W, rgba.data_ptr() & 15)
    if key in _warned_direct:
        return
    _warned_direct.add(key)
    why = ctypes.c_uint32(0)
    if _lib.load().gmpi_mpi_render_fwd_plan(V, N, Ht, Wt, H, W, rgba.data_ptr(), ctypes.byref(why)) == _lib.PLAN_DIRECT \
            and (why.value & ~2 or V * N * H * W >= 1 << 26):      # "few tiles" only matters when the problem is not tiny
        reasons = "; ".join(t for b, t in _lib.WHY.items() if why.value & b)
        warnings.warn(f"ml_gmpi_b200: rendering V={V} N={N} tex={Ht}x{Wt} img={H}x{W} with the direct (one thread per pixel) "
                      f"kernels, several times slower than the TMA-staged path: {reasons}", RuntimeWarning, stacklevel=3)


def render_views(rgba, dhw, view2mpi, ray_dir, eye, z_dir, *, align_corners=True, check_last_plane=False,
                 color_minus1_1=False, flags: Optional[torch.Tensor] = None):
    """Functional form on packed tensors (no list handling, no host sync).
    rgba [M,N,4,Ht,Wt], dhw [M,N,3], view2mpi [V] int32, ray_dir [V,3,H,W], eye/z_dir [V,3].
    Returns (color [V,3,H,W], depth [V,1,H,W]); `flags` (uint32 tensor of 1, int32 storage) is OR-ed into."""
    if not rgba.is_cuda:
        raise RuntimeError("ml_gmpi_b200 renders on CUDA devices only (no CPU fallback); got a CPU tensor")
    if flags is None:
        flags = torch.zeros(1, dtype=torch.int32, device=rgba.device)
    _warn_if_direct(rgba, ray_dir.shape[0], ray_dir.shape[2], ray_dir.shape[3])
    options = (_lib.OPT_ALIGN_CORNERS if align_corners else 0) | (_lib.OPT_CHECK_LAST_PLANE if check_last_plane else 0) \
        | (_lib.OPT_COLOR_MINUS1_1 if color_minus1_1 else 0)
    return _RenderFn.apply(_as_f32c(rgba), _as_f32c(dhw), view2mpi, _as_f32c(ray_dir), _as_f32c(eye), _as_f32c(z_dir),
                           options, flags)


def check_range(rgba: torch.Tensor, flags: torch.Tensor) -> None:
    """One streaming pass: RGBA/alpha in [0,1] (mpi_renderer.py:447-449, mpi.py:185-187) -> flag bits."""
    lib = _lib.load()
    M, N, _, Ht, Wt = rgba.shape
    with torch.cuda.device(rgba.device):
        _lib.check(lib.gmpi_mpi_check_range(rgba.data_ptr(), M, N, Ht, Wt, flags.data_ptr(), _stream_ptr(rgba.device)))


class MPI(nn.Module):
    """`validate`:
         "full"  (default) every data-dependent assert of the reference is evaluated on the device
                 (alpha range scan, plane-behind-camera, rays leaving the last plane) and raised
                 after ONE host sync per call (the reference syncs six or more times);
         "defer" the geometric flags are still computed inside the render kernel (free) but nothing
                 is scanned or synced; read them later with `.raise_if_flagged()`;
         "off"   like "defer" without the last-plane check.
    """

    def __init__(self, align_corners=True, validate: str = "full"):
        super().__init__()
        assert validate in ("full", "defer", "off"), validate
        self._align_corners = align_corners
        self.validate = validate
        self._flags = None
        self._flag_ctx = None

    # -- reference: MPI.check_shapes, mpi.py:161-216 (shape part; the alpha range is checked on the device)
    def check_shapes(self, *, batch_rgba, batch_dhw, batch_ray_dir, batch_eye_pos, batch_z_dir, separate_background):
        assert (batch_rgba.ndim == 5) and (batch_rgba.shape[2] == 4), (
            f"Expected rgba to be of shape (#mpi, #planes, 4, texture_height, texture_width), "
            f"but instead got {batch_rgba.shape}")
        assert ((batch_dhw.ndim == 3) and (batch_dhw.shape[0] == batch_rgba.shape[0])
                and (batch_dhw.shape[1] == batch_rgba.shape[1]) and (batch_dhw.shape[2] == 3)), (
            f"Expected dhw to be of shape (#mpi, #planes, 3), but instead got {batch_dhw.shape} (rgba: {batch_rgba.shape})")
        assert len(batch_ray_dir) == batch_rgba.shape[0], f"{len(batch_ray_dir)}, {batch_rgba.shape[0]}"
        assert len(batch_eye_pos) == batch_rgba.shape[0], f"{len(batch_eye_pos)}, {batch_rgba.shape[0]}"
        assert len(batch_z_dir) == batch_rgba.shape[0], f"{len(batch_z_dir)}, {batch_rgba.shape[0]}"
        for i in range(len(batch_ray_dir)):
            assert (batch_ray_dir[i].ndim == 4) and (batch_ray_dir[i].shape[1] == 3), (
                f"Expected ray_dir to be of shape (minibatch, 3, image_height, image_width), "
                f"but instead got {batch_ray_dir[i].shape} for {i} th elem.")
            assert (batch_eye_pos[i].ndim == 2) and (batch_eye_pos[i].shape[1] == 3), (
                f"Expected eye_pos to be of shape (minibatch, 3), but instead got {batch_eye_pos[i].shape} for {i} th elem.")
            assert (batch_z_dir[i].ndim == 2) and (batch_z_dir[i].shape[1] == 3), (
                f"Expected z_dir to be of shape (minibatch, 3), but instead got {batch_z_dir[i].shape} for {i} th elem.")
        if separate_background is not None:
            assert separate_background.ndim == 4 and separate_background.shape[1] == 3, (
                f"Expect background to be of shape (#mpi, 3, h, w), but instead get {separate_background.shape}.")

    @staticmethod
    def pack_views(batch_ray_dir, batch_eye_pos, batch_z_dir, device):
        """mpi.py:334-354 without the copies of the MPI: views stay MPI-major and a [V] int32 index
        replaces expand+cat of rgba/dhw."""
        counts = [int(r.shape[0]) for r in batch_ray_dir]
        if all(c == 1 for c in counts):
            view2mpi = torch.arange(len(counts), dtype=torch.int32, device=device)
        else:
            view2mpi = torch.repeat_interleave(torch.arange(len(counts), dtype=torch.int32),
                                               torch.tensor(counts)).to(device=device, dtype=torch.int32)
        cat = (lambda xs: xs[0] if len(xs) == 1 else torch.cat(xs, dim=0))
        return view2mpi, cat(list(batch_ray_dir)), cat(list(batch_eye_pos)), cat(list(batch_z_dir))

    def forward(self, *, batch_rgba: torch.Tensor, batch_dhw: torch.Tensor, batch_ray_dir: List[torch.Tensor],
                batch_eye_pos: List[torch.Tensor], batch_z_dir: List[torch.Tensor],
                separate_background: Union[None, torch.Tensor], assert_not_out_of_last_plane: bool = False,
                c2w_mat: torch.Tensor = None, sphere_c: np.ndarray = None):
        self.check_shapes(batch_rgba=batch_rgba, batch_dhw=batch_dhw, batch_ray_dir=batch_ray_dir,
                          batch_eye_pos=batch_eye_pos, batch_z_dir=batch_z_dir, separate_background=separate_background)
        if not batch_rgba.is_cuda:
            raise RuntimeError("ml_gmpi_b200.MPI renders on CUDA devices only (no CPU fallback); got a CPU tensor")
        dev = batch_rgba.device
        view2mpi, ray_dir, eye, z_dir = self.pack_views(batch_ray_dir, batch_eye_pos, batch_z_dir, dev)
        rgba = _as_f32c(batch_rgba)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        if self.validate == "full":
            check_range(rgba.detach(), flags)
        color, depth = render_views(rgba, batch_dhw.to(dev), view2mpi, ray_dir.to(dev), eye.to(dev), z_dir.to(dev),
                                    align_corners=self._align_corners,
                                    check_last_plane=bool(assert_not_out_of_last_plane) and self.validate != "off",
                                    flags=flags)
        self._flags = flags
        self._flag_ctx = (batch_dhw, eye, c2w_mat, sphere_c)
        if self.validate == "full":
            self.raise_if_flagged(ignore=_lib.FLAG_RGBA_RANGE)   # rgba range is MPIRenderer.render's assert
        return color, depth

    def last_flags(self) -> int:
        """Flag word of the most recent forward (one host sync)."""
        return 0 if self._flags is None else int(self._flags.item()) & 0xFFFFFFFF

    def raise_if_flagged(self, ignore: int = 0):
        f = self.last_flags() & ~ignore
        if f == 0:
            return
        dhw, eye, c2w, sphere_c = self._flag_ctx
        if f & _lib.FLAG_ALPHA_RANGE:
            raise AssertionError("Expected alpha to be within the the range [0, 1]")             # mpi.py:185-187
        if f & _lib.FLAG_RGBA_RANGE:
            raise AssertionError("MPI rgba outside [0, 1]")                                       # mpi_renderer.py:447-449
        if f & _lib.FLAG_PLANE_BEHIND_EYE:
            raise AssertionError(f"Camera must be placed closer to origin than MPI. {dhw[..., 0]}, {eye[0, ...]}")  # mpi.py:70-72
        if f & _lib.FLAG_LAST_PLANE_OOB:
            msg = f"Ray's U/V direction goes out of plane at {dhw[:, -1, 0]}"                    # mpi.py:106-109
            if c2w is not None and sphere_c is not None:
                msg += f"; c2w: {c2w.detach().cpu().numpy().tolist()}"
            raise MPIOutOfPlaneError(msg)
