"""Multi-GPU plumbing for the render path: one process per GPU, views sharded across ranks, ONE all-gather of
the rendered frames.  The render itself needs no collective: every frame depends on
one MPI and one pose (mpi.py:308-436 has no cross-view term).  NCCL on GPUs, gloo in the CPU tests."""
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` (first n_items % world ranks get one more)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_views_mpi_major(view2mpi: List[int], rank: int, world: int) -> Tuple[int, int]:
    """Partition by MPI first, then by view: whole MPIs go to one rank whenever there are at least
    as many MPIs as ranks, so d/d rgba never needs a cross-rank reduction; with fewer MPIs than ranks the views
    of an MPI are split and the (replicated) MPI's gradient must be all-reduced by the caller."""
    n_mpi = (max(view2mpi) + 1) if len(view2mpi) else 0
    if n_mpi >= world:
        m_lo, m_hi = shard_range(n_mpi, rank, world)
        idx = [i for i, m in enumerate(view2mpi) if m_lo <= m < m_hi]
        return (idx[0], idx[-1] + 1) if idx else (0, 0)
    return shard_range(len(view2mpi), rank, world)


def pack_frames(color: torch.Tensor, depth: torch.Tensor) -> torch.Tensor:
    """[V,3,H,W] + [V,1,H,W] -> [V,4,H,W] (RGB + depth = one 'frame', 16*H*W bytes)."""
    return torch.cat([color, depth], dim=1)


def all_gather_frames(frames: torch.Tensor, counts: List[int] = None, group=None) -> torch.Tensor:
    """Gather every rank's [V_r,4,H,W] frames into [sum V_r,4,H,W] on every rank, rank-major (= view order when
    views were sharded with shard_range).  Equal counts use one all_gather_into_tensor (ncclAllGather);
    ragged counts pad to the max and slice."""
    world = dist.get_world_size(group)
    if world == 1:
        return frames
    if counts is None:
        counts = [frames.shape[0]] * world
    vmax = max(counts)
    if all(c == vmax for c in counts):
        out = torch.empty((world * vmax,) + tuple(frames.shape[1:]), dtype=frames.dtype, device=frames.device)
        dist.all_gather_into_tensor(out, frames.contiguous(), group=group)
        return out
    pad = torch.zeros((vmax,) + tuple(frames.shape[1:]), dtype=frames.dtype, device=frames.device)
    pad[: frames.shape[0]] = frames
    out = torch.empty((world * vmax,) + tuple(frames.shape[1:]), dtype=frames.dtype, device=frames.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * vmax: r * vmax + counts[r]] for r in range(world)], dim=0)
