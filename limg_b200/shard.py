"""Host-side sharding of the encode path across the GPUs of one box.

The path partitions by independent units, so there is no data-path collective:
  * batch mode   -- frame i goes to rank i % world; every frame is one encode call with its own dither chain;
  * row-band mode -- one very large image is cut into bands of whole 8x8 block rows, one band per rank. Areas cannot cross
                     bands and every band restarts the dither chain, so the result is, by construction, the reference run per band.
Only the final gather of the (small) per-unit results uses torch.distributed.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

BLOCK = 8


def frames_for_rank(n_frames: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_frames, world))


def row_bands(height: int, world: int) -> List[Tuple[int, int]]:
    """Pixel row ranges [y0, y1) of the bands: ceil(blockY / world) block rows each (the last ones may be empty)."""
    block_rows = (height + BLOCK - 1) // BLOCK
    per = (block_rows + world - 1) // world
    bands = []
    for r in range(world):
        y0 = min(r * per * BLOCK, height)
        y1 = min((r + 1) * per * BLOCK, height)
        bands.append((y0, y1))
    return bands


def encode_frames_sharded(frames: Sequence, encode_fn: Callable, rank: int, world: int) -> List[Tuple[int, object]]:
    """Encodes this rank's frames; returns [(frame index, result)]."""
    return [(i, encode_fn(frames[i])) for i in frames_for_rank(len(frames), rank, world)]


def encode_bands_sharded(image, encode_fn: Callable, rank: int, world: int):
    """Encodes this rank's row band of `image` (2-D array); returns (y0, y1, result) or None for an empty band."""
    y0, y1 = row_bands(image.shape[0], world)[rank]
    if y1 <= y0:
        return None
    return (y0, y1, encode_fn(image[y0:y1]))


def gather_to_rank0(obj, rank: int, world: int):
    """Final gather of per-rank results (python objects) on rank 0; no-op for a single process."""
    if world == 1:
        return [obj]
    import torch.distributed as dist
    out = [None] * world if rank == 0 else None
    dist.gather_object(obj, out, dst=0)
    return out
