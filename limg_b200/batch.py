"""Batches of independent frames on one GPU (batch mode; BASELINE.json config 5).

Every frame is its own limg_blocked_encode3d_test call with its own dither chain, so frames need no exchange. One frame does not fill a
B200: the area scan is a latency-bound dependency chain (one CTA per SM, IPC 0.14), so `lanes` independent contexts (own streams and
scratch, limgcu_create each) work on different frames at the same time; the throughput kernels of one frame fill the gaps of another's
scan. Measured on B200, 1920x1080 frames, device-resident: 895 Mpixel/s with one lane, 2105 with four (profiles/r1_i_batch_time.txt).
The host-buffer entry points block until their result is in host memory, so each lane is driven by its own thread (ctypes releases the GIL).
"""
from __future__ import annotations

import ctypes as C
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, List, Sequence

import numpy as np

from ._lib import FLAG_FAST_BIT_CRUSH, LimgError
from .api import Codec


class BatchCodec:
    def __init__(self, device: int = 0, lanes: int = 4):
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        self.codecs: List[Codec] = [Codec(device) for _ in range(lanes)]
        self.pool = ThreadPoolExecutor(max_workers=lanes)

    @property
    def lanes(self) -> int:
        return len(self.codecs)

    def close(self):
        self.pool.shutdown(wait=True)
        for c in self.codecs:
            c.close()
        self.codecs = []

    def map(self, fn: Callable, items: Sequence) -> list:
        """fn(codec, item) for every item; item i runs on lane i % lanes, the items of one lane in order. Results in item order."""
        k = self.lanes
        out = [None] * len(items)

        def lane(j):
            for i in range(j, len(items), k):
                out[i] = fn(self.codecs[j], items[i])

        for f in [self.pool.submit(lane, j) for j in range(min(k, len(items)))]:
            f.result()
        return out

    def encode_streams(self, frames: Sequence, has_alpha: bool, error_factor: int = 100, fast_bit_crushing: bool = True, decoded: bool = False) -> list:
        return self.map(lambda c, f: c.encode_stream(f, has_alpha, error_factor, fast_bit_crushing, False, decoded), frames)

    def _ctx_array(self):
        return (C.c_void_p * self.lanes)(*[c.h for c in self.codecs])

    def _ck(self, rc: int, what: str):
        if rc != 0:
            texts = [c.lib.limgcu_last_error(c.h).decode() for c in self.codecs]
            raise LimgError(f"{what} failed with {rc}: {[t for t in texts if t]}")

    def encode_containers(self, frames: Sequence, has_alpha: bool, error_factor: int = 100, fast_bit_crushing: bool = True) -> List[bytes]:
        """frames of one size -> "LIMGB200" containers, through the C batch entry point (one std::thread per lane)."""
        frames = [np.ascontiguousarray(f, dtype=np.uint32) for f in frames]
        if not frames:
            return []
        h, w = frames[0].shape
        if any(f.shape != (h, w) for f in frames):
            raise ValueError("all frames of a batch must have the same size")
        lib = self.codecs[0].lib
        n = len(frames)
        cap = lib.limgcu_container_bound(w, h, int(has_alpha))
        outs = [np.zeros(cap, np.uint8) for _ in range(n)]
        written = (C.c_size_t * n)()
        rc = lib.limgcu_batch_host_encode_containers(self._ctx_array(), self.lanes, (C.c_void_p * n)(*[f.ctypes.data for f in frames]), n, w, h, int(has_alpha), int(error_factor),
                                                     FLAG_FAST_BIT_CRUSH if fast_bit_crushing else 0, (C.c_void_p * n)(*[o.ctypes.data for o in outs]), (C.c_size_t * n)(*([cap] * n)), written)
        self._ck(rc, "limgcu_batch_host_encode_containers")
        return [outs[i][: written[i]].tobytes() for i in range(n)]

    def decode_containers(self, containers: Sequence[bytes]) -> list:
        n = len(containers)
        if n == 0:
            return []
        infos = [self.codecs[0].container_info(d) for d in containers]
        bufs = [np.frombuffer(d, np.uint8) for d in containers]
        outs = [np.zeros((i["height"], i["width"]), np.uint32) for i in infos]
        lib = self.codecs[0].lib
        rc = lib.limgcu_batch_host_decode_containers(self._ctx_array(), self.lanes, (C.c_void_p * n)(*[b.ctypes.data for b in bufs]), (C.c_size_t * n)(*[b.size for b in bufs]), n,
                                                     (C.c_void_p * n)(*[o.ctypes.data for o in outs]), (C.c_size_t * n)(*[o.size for o in outs]))
        self._ck(rc, "limgcu_batch_host_decode_containers")
        return outs
