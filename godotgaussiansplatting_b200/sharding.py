"""Tile-row band sharding across the GPUs of one node (no counterpart in the single-device reference).

Every rank holds the full SoA splat set and runs the projection over all N splats, but clamps each splat's tile
rect to its own contiguous band of tile rows, so it sorts, scans and blends only its band.  Keys keep the global
tile id, therefore concatenating the per-rank sorted key arrays in band order reproduces the single-GPU sorted
array (tests/test_gpu_pipeline.py::test_bands_concatenate_to_full_frame).  One collective per frame: the band
framebuffers are gathered on the presenting rank (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

TILE_SIZE = 16


def band_partition(tiles_y: int, world: int) -> list[tuple[int, int]]:
    """Equal-height bands of ceil(tiles_y/world) tile rows; trailing ranks may get a short or empty band."""
    rows = (tiles_y + world - 1) // world
    return [(min(r * rows, tiles_y), min((r + 1) * rows, tiles_y)) for r in range(world)]


def padded_height(height: int, world: int) -> int:
    """Rows of the gather buffer: world equal slabs of band_rows*16 pixel rows (>= height)."""
    tiles_y = (height + TILE_SIZE - 1) // TILE_SIZE
    rows = (tiles_y + world - 1) // world
    return rows * TILE_SIZE * world


def slab_rows(height: int, world: int) -> int:
    return padded_height(height, world) // world


def gather_bands(frame, rank: int, world: int, dst: int = 0):
    """Gathers every rank's slab of `frame` ([padded_height, W, 4] tensor on each rank) into rank `dst`'s `frame`,
    in place: one torch.distributed.gather per frame.  Works on CUDA tensors (NCCL) and CPU tensors (gloo)."""
    import torch.distributed as dist
    rows = frame.shape[0] // world
    mine = frame[rank * rows:(rank + 1) * rows]
    if rank == dst:
        dist.gather(mine, [frame[r * rows:(r + 1) * rows] for r in range(world)], dst=dst)
    else:
        dist.gather(mine, None, dst=dst)
    return frame
