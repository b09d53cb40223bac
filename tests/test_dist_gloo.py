"""N>1 host logic on CPU (gloo, world_size 2): view sharding and the single all-gather of frames
(SURVEY.md section 8e).  The render itself needs no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ml_gmpi_b200 import dist as gdist


def test_shard_range_is_balanced_partition():
    for n in (0, 1, 4, 7, 120, 121):
        for world in (1, 2, 3, 8):
            parts = [gdist.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    assert gdist.shard_range(120, 3, 8) == (45, 60)      # BASELINE configs[3]: 120 views over 8 GPUs -> 15 each


def test_shard_views_keeps_whole_mpis_on_one_rank():
    v2m = [0, 0, 1, 1, 1, 2, 3, 3]
    parts = [gdist.shard_views_mpi_major(v2m, r, 2) for r in range(2)]
    assert parts == [(0, 5), (5, 8)]
    for lo, hi in parts:            # no MPI is split across ranks -> d/d rgba needs no cross-rank reduction
        assert not (set(v2m[lo:hi]) & (set(v2m[:lo]) | set(v2m[hi:])))
    # fewer MPIs than ranks (video render, one MPI, 120 views): split the views
    assert [gdist.shard_views_mpi_major([0] * 120, r, 8) for r in (0, 7)] == [(0, 15), (105, 120)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, counts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = sum(counts)
        lo = sum(counts[:rank])
        allf = torch.arange(total * 4 * 3 * 5, dtype=torch.float32).reshape(total, 4, 3, 5)   # the "true" frames
        mine = allf[lo: lo + counts[rank]].clone()
        color, depth = mine[:, :3], mine[:, 3:]
        got = gdist.all_gather_frames(gdist.pack_frames(color, depth), counts)
        q.put((rank, bool(torch.equal(got, allf)), tuple(got.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[3, 3], [4, 2]])
def test_all_gather_frames_world2(counts):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok and shape == (sum(counts), 4, 3, 5), (rank, ok, shape)
