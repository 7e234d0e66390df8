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


def band_block_rows(y0: int, y1: int) -> Tuple[int, int]:
    """Block rows [lo, hi) a band of pixel rows [y0, y1) owns in the exact row-band mode; an empty band (more ranks than block rows)
    owns none, whatever the image height is."""
    if y1 <= y0:
        return (0, 0)
    return (y0 // BLOCK, (y1 + BLOCK - 1) // BLOCK)


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


# ---------------------------------------------------------------------------------------------------------------------------------
# whole-image-exact row bands (C ABI: limgcu_pass1 / limgcu_merge / limgcu_encode_areas / limgcu_finalize_rows)
# ---------------------------------------------------------------------------------------------------------------------------------

def padded_block_rows(height: int, world: int) -> int:
    """Block rows of the table / source when every rank's band is padded to the same size (what an all-gather needs): world * ceil(blockY / world)."""
    block_rows = (height + BLOCK - 1) // BLOCK
    return world * ((block_rows + world - 1) // world)


class RowBandExact:
    """One rank of the exact row-band encode: the same stream as ONE encode of the whole image (areas may cross the bands, one dither chain).

    Exchange steps: the source bands and the pass-1 table (64 B per block) are ALL-GATHERED (every rank's band is the same number of block rows,
    the last ones padded, so the chunks are equal and the gather runs in place); the per-area results (80 B per area, zero wherever another rank
    encodes the area) are SUM all-reduced with an integer view, which is exact. The scan runs redundantly on every rank (it is deterministic), the
    per-area refit + shift search and the per-pixel finalize are sharded.
    Phases: pass1() -> [all-gather table] -> merge_and_encode() -> [all-reduce results] -> finalize(). `encode_rowbands_exact` drives them with
    torch.distributed on the codec's own stream (no host synchronisation between the phases); the tests drive several ranks on one GPU by hand."""

    def __init__(self, codec, d_src, width: int, height: int, has_alpha: bool, rank: int, world: int, error_factor: int = 100, fast_bit_crushing: bool = True):
        import torch
        from ._lib import AREA_DTYPE, FLAG_FAST_BIT_CRUSH
        self.codec, self.src, self.w, self.h, self.alpha = codec, d_src, width, height, bool(has_alpha)
        self.rank, self.world, self.ef = rank, world, int(error_factor)
        self.flags = FLAG_FAST_BIT_CRUSH if fast_bit_crushing else 0
        self.bx, self.by = (width + 7) // 8, (height + 7) // 8
        self.y0, self.y1 = row_bands(height, world)[rank]
        self.row_lo, self.row_hi = band_block_rows(self.y0, self.y1)
        self.rows_per_rank = padded_block_rows(height, world) // world
        dev = d_src.device
        blocks = self.bx * self.by
        words = int(codec.lib.limgcu_area_result_words())
        # limgcu_decomp[padded block rows][bx], 64 B each; chunk r = block rows [r * rows_per_rank, (r + 1) * rows_per_rank)
        self.table = torch.zeros(world * self.rows_per_rank * self.bx * 16, dtype=torch.int32, device=dev)
        self.results = torch.zeros(blocks * words, dtype=torch.int32, device=dev)
        self.areas = torch.zeros(blocks * AREA_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        self.count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.block_to_area = torch.zeros(blocks, dtype=torch.int32, device=dev)
        self.codes = [torch.zeros((height, width), dtype=torch.uint8, device=dev) for _ in range(3)]  # only rows [y0, y1) are written
        # the buffers were zeroed on torch's stream; the codec launches on its own stream
        torch.cuda.current_stream(dev).synchronize()

    def _ck(self, rc, what):
        self.codec._ck(rc, what)

    def table_chunk(self):
        """this rank's equal-sized chunk of the table (the input of the in-place all-gather)"""
        n = self.rows_per_rank * self.bx * 16
        return self.table[self.rank * n:(self.rank + 1) * n]

    def pass1(self):
        """three-factor fit of the blocks of this rank's band -> its rows of the table; stream ordered, no synchronisation"""
        if self.y1 > self.y0:
            lib, c = self.codec.lib, self.codec
            self._ck(lib.limgcu_pass1(c.h, self.src.data_ptr() + self.y0 * self.w * 4, self.w, self.y1 - self.y0, int(self.alpha), self.table.data_ptr() + self.row_lo * self.bx * 64), "limgcu_pass1")
        return self.table

    def merge_and_encode(self):
        """the identical scan on the complete table, then refit + shift search of the areas that start in this rank's block rows; stream ordered
        (a scan that timed out or overflowed is reported by finalize(), on every rank alike: they all run the same scan)"""
        lib, c = self.codec.lib, self.codec
        self._ck(lib.limgcu_merge(c.h, self.table.data_ptr(), self.w, self.h, int(self.alpha), self.areas.data_ptr(), self.count.data_ptr(), self.block_to_area.data_ptr()), "limgcu_merge")
        self._ck(lib.limgcu_encode_areas(c.h, self.src.data_ptr(), self.w, self.h, int(self.alpha), self.ef, self.flags, self.table.data_ptr(), self.areas.data_ptr(), self.row_lo, self.row_hi,
                                         self.results.data_ptr()), "limgcu_encode_areas")
        return self.results

    def finalize(self):
        """complete per-area results -> area table, dither chain, codes of this rank's pixel rows. Synchronises (and checks the scan's hard-error flags)."""
        from ._lib import Stream
        import ctypes as C
        lib, c = self.codec.lib, self.codec
        st = Stream()
        st.areas, st.area_count, st.block_to_area = self.areas.data_ptr(), self.count.data_ptr(), self.block_to_area.data_ptr()
        st.codesA, st.codesB, st.codesC = (t.data_ptr() for t in self.codes)
        self._ck(lib.limgcu_finalize_rows(c.h, self.src.data_ptr(), self.w, self.h, int(self.alpha), self.flags, self.areas.data_ptr(), self.results.data_ptr(), self.block_to_area.data_ptr(),
                                          C.byref(st), None, self.y0, self.y1), "limgcu_finalize_rows")
        return self

    def area_table(self):
        import numpy as np
        from ._lib import AREA_DTYPE
        n = int(self.count.item())
        return np.frombuffer(self.areas.cpu().numpy().tobytes(), dtype=AREA_DTYPE, count=n).copy()


def encode_rowbands_exact(codec, d_src, width: int, height: int, has_alpha: bool, rank: int, world: int, error_factor: int = 100, fast_bit_crushing: bool = True,
                          band: "RowBandExact | None" = None) -> RowBandExact:
    """All ranks call this with the whole source in device memory (all-gather it first if every rank holds only its band, see bench.py). Returns the
    rank's RowBandExact: the complete area table and the codes of its own pixel rows. The collectives are issued on the codec's stream, so the only
    host synchronisation is the one at the end of finalize(). `band`: a RowBandExact to reuse (its buffers are allocated and zeroed once)."""
    import torch
    r = band if band is not None else RowBandExact(codec, d_src, width, height, has_alpha, rank, world, error_factor, fast_bit_crushing)
    r.src = d_src
    if world == 1:
        r.pass1()
        r.merge_and_encode()
        return r.finalize()
    import torch.distributed as dist
    with torch.cuda.stream(torch.cuda.ExternalStream(codec.stream, device=d_src.device)):
        if band is not None:
            r.results.zero_()  # areas of other ranks must contribute zeros to the sum
        r.pass1()
        dist.all_gather_into_tensor(r.table, r.table_chunk())
        r.merge_and_encode()
        dist.all_reduce(r.results)
    return r.finalize()
