"""Python mirror of the reference's operator interface (limg.h) on top of the C ABI.

`Codec` owns one limgcu context (one CUDA device, one stream). Host-buffer methods take / return numpy
arrays with the reference's argument meaning; device-buffer methods take raw device pointers (ints), e.g.
`torch.Tensor.data_ptr()`, and are stream ordered on `Codec.stream`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import AREA_DTYPE, DECOMP_DTYPE, FLAG_DITHER_AES, FLAG_FAST_BIT_CRUSH, FLAG_NO_MERGE, PLANE_ORDER, PLANES_U8, LimgError, Planes, Stream

PHASES = ("pass1", "predicate_windows", "merge_scan", "area_encode", "dither_scan", "finalize")


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Codec:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.limgcu_create(int(device), C.byref(h))
        if rc != 0:
            raise LimgError(f"limgcu_create(device={device}) failed with {rc} (200 = no CUDA device; there is no CPU fallback)")
        self.h = h
        self.device = device
        self.aes = False

    def close(self):
        if getattr(self, "h", None):
            self.lib.limgcu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str):
        if rc != 0:
            raise LimgError(f"{what} failed with {rc}: {self.lib.limgcu_last_error(self.h).decode()}")

    @property
    def stream(self) -> int:
        return int(self.lib.limgcu_stream_handle(self.h) or 0)

    def sync(self):
        self._ck(self.lib.limgcu_sync(self.h), "limgcu_sync")

    def status(self):
        """Synchronises and raises if the last merge scan flagged a hard error (watchdog, row-list overflow): the stream-ordered
        device entry points cannot report those themselves."""
        self._ck(self.lib.limgcu_status(self.h), "limgcu_status")

    def launch_count(self) -> int:
        return int(self.lib.limgcu_launch_count(self.h))

    def set_dither_mode(self, aes: bool):
        """Dither generator: False = the reference's LCG (default), True = its AES-round chain (what it uses on hosts with AES-NI)."""
        self.aes = bool(aes)
        self._ck(self.lib.limgcu_set_dither_mode(self.h, int(self.aes)), "limgcu_set_dither_mode")

    def set_decode_variant(self, variant: int):
        self._ck(self.lib.limgcu_debug_set_decode_variant(self.h, int(variant)), "limgcu_debug_set_decode_variant")

    def set_rsqrt_lut(self, lut=None):
        if lut is not None:
            lut = np.ascontiguousarray(lut, dtype=np.uint16)
            assert lut.size == 2048
        self._ck(self.lib.limgcu_set_rsqrt_lut(self.h, _vp(lut)), "limgcu_set_rsqrt_lut")

    def enable_phase_timing(self, on: bool = True):
        self.lib.limgcu_enable_phase_timing(self.h, int(on))

    def phase_ms(self) -> dict:
        return {name: float(self.lib.limgcu_phase_ms(self.h, i)) for i, name in enumerate(PHASES)}

    def debug_counters(self) -> np.ndarray:
        out = np.zeros(32, np.uint32)
        self._ck(self.lib.limgcu_debug_counters(self.h, _vp(out)), "limgcu_debug_counters")
        return out

    def debug_wave(self) -> np.ndarray:
        out = np.zeros(256, np.uint32)
        self._ck(self.lib.limgcu_debug_wave(self.h, _vp(out)), "limgcu_debug_wave")
        return out

    def predicate_check(self, table, w: int, h: int, has_alpha: bool) -> np.ndarray:
        """(scored pairs, decided by the shortcut, disagreements, reserved) -- see limgcu_debug_predicate_check."""
        table = np.ascontiguousarray(table, dtype=DECOMP_DTYPE)
        out = np.zeros(4, np.uint64)
        self._ck(self.lib.limgcu_debug_predicate_check(self.h, _vp(table), w, h, int(has_alpha), _vp(out)), "limgcu_debug_predicate_check")
        return out

    def debug_wave_rows(self, block_y: int) -> np.ndarray:
        out = np.zeros(block_y * 8 + 512, np.uint32)
        self._ck(self.lib.limgcu_debug_wave_rows(self.h, _vp(out), block_y), "limgcu_debug_wave_rows")
        self.wave_events = out[block_y * 8:].reshape(4, 64, 2)
        return out[: block_y * 8].reshape(2, block_y, 4)

    # ---- host-buffer operators (reference argument meaning) -------------------------------------------------

    @staticmethod
    def alloc_planes(h: int, w: int, names=PLANE_ORDER) -> dict:
        return {k: np.zeros((h, w), dtype=np.uint8 if k in PLANES_U8 else np.uint32) for k in names}

    @staticmethod
    def _planes_struct(p: dict) -> Planes:
        s = Planes()
        for k in PLANE_ORDER:
            a = p.get(k)
            setattr(s, k, a.ctypes.data if a is not None else None)
        return s

    def blocked_encode3d_test(self, img, has_alpha: bool, planes: dict | None = None, error_factor: int = 100, fast_bit_crushing: bool = True) -> dict:
        """limg_blocked_encode3d_test (limg.h:46)."""
        img = np.ascontiguousarray(img, dtype=np.uint32)
        h, w = img.shape
        planes = self.alloc_planes(h, w) if planes is None else planes
        s = self._planes_struct(planes)
        self._ck(self.lib.limgcu_host_blocked_encode3d(self.h, _vp(img), w, h, int(has_alpha), C.byref(s), int(error_factor), int(fast_bit_crushing)), "limg_blocked_encode3d_test")
        return planes

    def encode3d_test(self, img, has_alpha: bool, planes: dict | None = None, error_factor: int = 100, fast_bit_crushing: bool = True, pool_threads: int = 0) -> dict:
        """limg_encode3d_test (limg.h:35): every 8x8 block its own area. pool_threads == 0: pThreadPool == nullptr, one dither chain over all blocks
        in raster order; > 0: the reference's run with a pool of that many threads (one chain per y-band, limg.cpp:2108-2137)."""
        img = np.ascontiguousarray(img, dtype=np.uint32)
        h, w = img.shape
        names = [k for k in PLANE_ORDER if k not in ("pBlockError", "pBitsPerPixel", "pBlockIndex")]
        planes = self.alloc_planes(h, w, names) if planes is None else planes
        s = self._planes_struct(planes)
        self.lib.limgcu_set_pool_threads(self.h, int(pool_threads))
        try:
            self._ck(self.lib.limgcu_host_encode3d(self.h, _vp(img), w, h, int(has_alpha), C.byref(s), int(error_factor), int(fast_bit_crushing)), "limg_encode3d_test")
        finally:
            self.lib.limgcu_set_pool_threads(self.h, 0)
        return planes

    def encode_stream(self, img, has_alpha: bool, error_factor: int = 100, fast_bit_crushing: bool = True, no_merge: bool = False, decoded: bool = False) -> dict:
        """Compact stream: area table + right-aligned code planes (image layout)."""
        img = np.ascontiguousarray(img, dtype=np.uint32)
        h, w = img.shape
        areas = np.zeros(((h + 7) // 8) * ((w + 7) // 8), dtype=AREA_DTYPE)
        codes = [np.zeros((h, w), np.uint8) for _ in range(3)]
        dec = np.zeros((h, w), np.uint32) if decoded else None
        n = C.c_uint32(0)
        flags = (FLAG_FAST_BIT_CRUSH if fast_bit_crushing else 0) | (FLAG_NO_MERGE if no_merge else 0) | (FLAG_DITHER_AES if self.aes else 0)
        self._ck(self.lib.limgcu_host_encode_stream(self.h, _vp(img), w, h, int(has_alpha), int(error_factor), flags, _vp(areas), C.byref(n),
                                                    _vp(codes[0]), _vp(codes[1]), _vp(codes[2]), _vp(dec)), "limgcu_host_encode_stream")
        out = {"areas": areas[: n.value].copy(), "codesA": codes[0], "codesB": codes[1], "codesC": codes[2], "width": w, "height": h, "has_alpha": has_alpha}
        if decoded:
            out["decoded"] = dec
        return out

    def encode_container(self, img, has_alpha: bool, error_factor: int = 100, fast_bit_crushing: bool = True, no_merge: bool = False) -> bytes:
        """Image -> "LIMGB200" container bytes (include/limgcu.h): header, area table, bit-packed codes."""
        img = np.ascontiguousarray(img, dtype=np.uint32)
        h, w = img.shape
        buf = np.zeros(self.lib.limgcu_container_bound(w, h, int(has_alpha)), np.uint8)
        n = C.c_size_t(0)
        flags = (FLAG_FAST_BIT_CRUSH if fast_bit_crushing else 0) | (FLAG_NO_MERGE if no_merge else 0) | (FLAG_DITHER_AES if self.aes else 0)
        self._ck(self.lib.limgcu_host_encode_container(self.h, _vp(img), w, h, int(has_alpha), int(error_factor), flags, _vp(buf), buf.size, C.byref(n)), "limgcu_host_encode_container")
        return buf[: n.value].tobytes()

    def container_info(self, data: bytes) -> dict:
        buf = np.frombuffer(data, np.uint8)
        w, h, a, n, pb = C.c_size_t(0), C.c_size_t(0), C.c_int(0), C.c_uint32(0), C.c_uint64(0)
        self._ck(self.lib.limgcu_container_info(_vp(buf), buf.size, C.byref(w), C.byref(h), C.byref(a), C.byref(n), C.byref(pb)), "limgcu_container_info")
        return {"width": w.value, "height": h.value, "has_alpha": bool(a.value), "area_count": n.value, "payload_bytes": pb.value}

    def decode_container(self, data: bytes) -> np.ndarray:
        info = self.container_info(data)
        buf = np.frombuffer(data, np.uint8)
        out = np.zeros((info["height"], info["width"]), np.uint32)
        self._ck(self.lib.limgcu_host_decode_container(self.h, _vp(buf), buf.size, _vp(out), out.size), "limgcu_host_decode_container")
        return out

    def decode(self, areas, codesA, codesB, codesC, has_alpha: bool) -> np.ndarray:
        """Reconstruction from a stream (limg_decode_block_from_factors_3d per area, limg_decode.h:326)."""
        areas = np.ascontiguousarray(areas, dtype=AREA_DTYPE)
        codesA = np.ascontiguousarray(codesA, np.uint8); codesB = np.ascontiguousarray(codesB, np.uint8); codesC = np.ascontiguousarray(codesC, np.uint8)
        h, w = codesA.shape
        out = np.zeros((h, w), np.uint32)
        self._ck(self.lib.limgcu_host_decode(self.h, _vp(areas), areas.size, _vp(codesA), _vp(codesB), _vp(codesC), w, h, int(has_alpha), _vp(out)), "limgcu_host_decode")
        return out

    def pass1(self, img, has_alpha: bool) -> np.ndarray:
        img = np.ascontiguousarray(img, dtype=np.uint32)
        h, w = img.shape
        table = np.zeros(((h + 7) // 8) * ((w + 7) // 8), dtype=DECOMP_DTYPE)
        self._ck(self.lib.limgcu_host_pass1(self.h, _vp(img), w, h, int(has_alpha), _vp(table)), "limgcu_host_pass1")
        return table

    def merge(self, table, w: int, h: int, has_alpha: bool) -> np.ndarray:
        table = np.ascontiguousarray(table, dtype=DECOMP_DTYPE)
        areas = np.zeros(table.size, dtype=AREA_DTYPE)
        n = C.c_uint32(0)
        self._ck(self.lib.limgcu_host_merge(self.h, _vp(table), w, h, int(has_alpha), _vp(areas), C.byref(n)), "limgcu_host_merge")
        return areas[: n.value].copy()

    def compare(self, a, b, has_alpha: bool):
        """limg_compare (limg.h:48): returns (psnr, mse, max_error)."""
        a = np.ascontiguousarray(a, dtype=np.uint32); b = np.ascontiguousarray(b, dtype=np.uint32)
        mse = C.c_double(); mx = C.c_double()
        psnr = self.lib.limgcu_host_compare(self.h, _vp(a), _vp(b), a.shape[1], a.shape[0], int(has_alpha), C.byref(mse), C.byref(mx))
        return float(psnr), mse.value, mx.value

    # ---- device-buffer operators (raw device pointers, stream ordered) ---------------------------------------

    def blocked_encode3d_device(self, d_src: int, w: int, h: int, has_alpha: bool, error_factor: int = 100, fast_bit_crushing: bool = True, no_merge: bool = False,
                                stream: dict | None = None, planes: dict | None = None):
        st = Stream()
        for k in ("areas", "area_count", "block_to_area", "codesA", "codesB", "codesC"):
            setattr(st, k, (stream or {}).get(k))
        pl = Planes()
        for k in PLANE_ORDER:
            setattr(pl, k, (planes or {}).get(k))
        flags = (FLAG_FAST_BIT_CRUSH if fast_bit_crushing else 0) | (FLAG_NO_MERGE if no_merge else 0) | (FLAG_DITHER_AES if self.aes else 0)
        self._ck(self.lib.limgcu_blocked_encode3d(self.h, d_src, w, h, int(has_alpha), int(error_factor), flags, C.byref(st), C.byref(pl)), "limgcu_blocked_encode3d")

    def decode_device(self, d_areas: int, d_block_to_area: int, d_codesA: int, d_codesB: int, d_codesC: int, w: int, h: int, has_alpha: bool, d_dst: int):
        self._ck(self.lib.limgcu_decode(self.h, d_areas, d_block_to_area, d_codesA, d_codesB, d_codesC, w, h, int(has_alpha), d_dst), "limgcu_decode")

    def compare_device(self, d_a: int, d_b: int, w: int, h: int, has_alpha: bool):
        psnr = C.c_double(); mse = C.c_double(); mx = C.c_double()
        self._ck(self.lib.limgcu_compare(self.h, d_a, d_b, w, h, int(has_alpha), C.byref(psnr), C.byref(mse), C.byref(mx)), "limgcu_compare")
        return psnr.value, mse.value, mx.value
