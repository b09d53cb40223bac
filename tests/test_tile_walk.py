"""Tile order of the persistent kernels (csrc/mpi_fwd_staged.cuh: TileWalk), evaluated on the host through the C ABI.

Every tile of every view must be walked by exactly one CTA, full-height tiles before partial ones, and the partial
(cheap) tiles must go to the CTAs that received fewer full tiles.  No GPU work: runs in the CPU suite.
"""
import ctypes

import numpy as np
import pytest

from ml_gmpi_b200 import _lib

TILE_W, TILE_H = 64, 30


def walk(H, W, V, grid, cta):
    lib = _lib.load()
    n = lib.gmpi_debug_tile_walk(H, W, V, grid, cta, None, 0)
    assert n >= 0
    out = np.zeros((max(n, 1), 3), dtype=np.int32)
    m = lib.gmpi_debug_tile_walk(H, W, V, grid, cta, out.ctypes.data_as(ctypes.c_void_p), n)
    assert m == n
    return [tuple(int(x) for x in row) for row in out[:n]]


@pytest.mark.parametrize("H,W,V,grid", [
    (1024, 1024, 4, 148),     # bench workload: 2176 full + 64 partial tiles
    (512, 512, 15, 148), (256, 256, 8, 148), (224, 224, 3, 148),
    (12, 16, 1, 1), (24, 24, 2, 2), (40, 40, 1, 2), (64, 64, 4, 12),      # fewer tiles than SMs: grid = number of tiles
    (60, 128, 5, 7), (90, 70, 3, 148), (31, 65, 2, 5), (1000, 1000, 2, 132),
])
def test_every_tile_exactly_once_full_before_partial(H, W, V, grid):
    tiles_x, tiles_y = -(-W // TILE_W), -(-H // TILE_H)
    expected = {(v, x * TILE_W, y * TILE_H) for v in range(V) for x in range(tiles_x) for y in range(tiles_y)}
    seen = []
    loads = []
    for cta in range(grid):
        t = walk(H, W, V, grid, cta)
        part = [py0 + TILE_H > H for (_, _, py0) in t]
        assert part == sorted(part), "a CTA must finish its full tiles before its partial ones"
        loads.append((len(part) - sum(part), sum(part)))
        seen += t
    assert len(seen) == len(expected) and set(seen) == expected
    full = [f for f, _ in loads]
    assert max(full) - min(full) <= 1
    if H % TILE_H and max(full) != min(full):      # partial tiles only where a full tile is missing
        assert all(p == 0 for f, p in loads if f == max(full))


def test_rejects_bad_arguments():
    lib = _lib.load()
    assert lib.gmpi_debug_tile_walk(0, 64, 1, 1, 0, None, 0) < 0
    assert lib.gmpi_debug_tile_walk(64, 64, 1, 4, 4, None, 0) < 0
    assert b"tile_walk" in lib.gmpi_last_error()
