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


def walk_ex(H, W, V, tile_h, group, grid, cta):
    lib = _lib.load()
    n = lib.gmpi_debug_tile_walk_ex(H, W, V, tile_h, group, grid, cta, None, 0)
    assert n >= 0
    out = np.zeros((max(n, 1), 3), dtype=np.int32)
    assert lib.gmpi_debug_tile_walk_ex(H, W, V, tile_h, group, grid, cta, out.ctypes.data_as(ctypes.c_void_p), n) == n
    return [tuple(int(x) for x in row) for row in out[:n]]


@pytest.mark.parametrize("H,W,V,tile_h,group,grid", [
    (512, 512, 120, 30, 120, 148),     # video render: 120 views of one MPI
    (512, 512, 15, 30, 15, 148), (256, 256, 8, 30, 4, 148), (224, 224, 12, 30, 3, 37),
    (1024, 1024, 4, 24, 1, 148), (512, 512, 4, 24, 1, 148), (200, 200, 6, 24, 2, 148),      # backward tiles (64 x 24)
    (100, 70, 4, 24, 4, 5), (48, 64, 2, 24, 2, 2),
])
def test_grouped_walk_covers_every_tile_once(H, W, V, tile_h, group, grid):
    tiles_x, tiles_y = -(-W // TILE_W), -(-H // tile_h)
    expected = {(v, x * TILE_W, y * tile_h) for v in range(V) for x in range(tiles_x) for y in range(tiles_y)}
    seen, first = [], []
    for cta in range(grid):
        t = walk_ex(H, W, V, tile_h, group, grid, cta)
        part = [py0 + tile_h > H for (_, _, py0) in t]
        assert part == sorted(part)
        seen += t
        if t:
            first.append(t[0])
    assert len(seen) == len(expected) and set(seen) == expected
    if group > 1 and grid >= 2 * group and tiles_x * (H // tile_h) >= 2:
        # the CTAs that start together work on the same tile position of the views of one group (L2 reuse)
        pos0 = [(px, py) for (v, px, py) in first[:group]]
        assert len(set(pos0)) == 1 and sorted(v for (v, _, _) in first[:group]) == list(range(group))


def test_group_that_does_not_divide_v_falls_back_to_view_major():
    assert walk_ex(256, 256, 7, 30, 4, 148, 0) == walk_ex(256, 256, 7, 30, 1, 148, 0)


@pytest.mark.parametrize("n_rows", list(range(0, 48, 4)))
def test_copy_plan_tiles_the_footprint_with_at_most_three_copies(n_rows):
    """The expanded forward stages a footprint of n_rows rows as the binary digits of n_rows / 4 (TMA copies of 32, 16, 8, 4 rows;
    UTMALDG copies are issued one after another, so few tall ones): they must tile [0, n_rows) exactly, tallest first."""
    lib = _lib.load()
    out = np.zeros((4, 2), dtype=np.int32)
    n = lib.gmpi_debug_copy_plan(n_rows, out.ctypes.data_as(ctypes.c_void_p), 4)
    copies = [tuple(int(x) for x in r) for r in out[:n]]
    assert n == bin(n_rows // 4).count("1") <= 3
    row = 0
    for first, rows in copies:
        assert first == row and rows in (32, 16, 8, 4)
        row += rows
    assert row == n_rows
    assert [r for _, r in copies] == sorted((r for _, r in copies), reverse=True)


def test_copy_plan_rejects_bad_row_counts():
    lib = _lib.load()
    assert lib.gmpi_debug_copy_plan(6, None, 0) < 0 and lib.gmpi_debug_copy_plan(48, None, 0) < 0 and lib.gmpi_debug_copy_plan(-4, None, 0) < 0
