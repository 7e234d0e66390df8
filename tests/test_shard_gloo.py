"""Multi-process host logic (world size 2, gloo, CPU): frame and row-band sharding + the final gather reproduce the single-process
result. The encoder stand-in is the C oracle (there is no GPU in this test); on the GPU box the same sharding drives Codec."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from limg_b200 import shard, synth


def _encode(img):
    from oracle import oracle as lo
    o = lo.blocked_encode3d(np.ascontiguousarray(img), False, 100, True)
    return (len(o["areas"]), int(o["planes"]["pDecoded"].astype(np.uint64).sum()), o["areas"]["shift"].tobytes())


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = [synth.photo_like(96, 64, 40 + i, 3) for i in range(5)]
    mine = shard.encode_frames_sharded(frames, _encode, rank, world)
    big = synth.photo_like(128, 136, 77, 3)
    band = shard.encode_bands_sharded(big, _encode, rank, world)
    gathered = shard.gather_to_rank0((mine, band), rank, world)
    dist.barrier()
    if rank == 0:
        q.put(gathered)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partitions():
    assert shard.frames_for_rank(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((shard.frames_for_rank(1024, r, 8) for r in range(8)), [])) == list(range(1024))
    bands = shard.row_bands(4320, 8)
    assert bands[0] == (0, 544) and bands[-1][1] == 4320 and all(b[0] % 8 == 0 for b in bands)
    assert sum(y1 - y0 for y0, y1 in bands) == 4320
    assert shard.row_bands(20, 4) == [(0, 8), (8, 16), (16, 20), (20, 20)]
    # exact row-band mode: the block-row ranges [y0 // 8, ceil(y1 / 8)) of the ranks own every block row exactly once
    # (also with more ranks than block rows and a ragged height: the empty bands own nothing)
    for h in (8, 20, 36, 200, 1080, 2160, 4320, 4321):
        for world in (1, 2, 3, 4, 8):
            owners = []
            for y0, y1 in shard.row_bands(h, world):
                lo, hi = shard.band_block_rows(y0, y1)
                owners += list(range(lo, hi))
            assert owners == list(range((h + 7) // 8)), (h, world)
    assert [shard.band_block_rows(*b) for b in shard.row_bands(36, 8)] == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (0, 0), (0, 0), (0, 0)]


@pytest.mark.timeout(300)
def test_world_size_2_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    frames = [synth.photo_like(96, 64, 40 + i, 3) for i in range(5)]
    want = {i: _encode(f) for i, f in enumerate(frames)}
    got = {}
    for mine, _ in gathered:
        for i, r in mine:
            got[i] = r
    assert got == want  # batch mode: identical to the single-process run, frame by frame

    big = synth.photo_like(128, 136, 77, 3)
    for (_, band), (y0, y1) in zip(gathered, shard.row_bands(136, world)):
        assert band[0] == y0 and band[1] == y1
        assert band[2] == _encode(big[y0:y1])  # row-band mode: the reference run per band
