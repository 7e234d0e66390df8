"""ctypes loader of limg_b200/liblimgcu.so (the C ABI of include/limgcu.h).

The library is built in-tree by __graft_entry__.build() / `make -C limg_b200/csrc`. There is no CPU
fallback: if the shared library or a CUDA device is missing the loader raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LIMGCU_LIB") or os.path.join(_HERE, "liblimgcu.so")  # LIMGCU_LIB: another build of the same library (profile counters)

DECOMP_DTYPE = np.dtype([
    ("avg", "<f4", (4,)),
    ("dirA_min", "<i2", (4,)), ("dirA_max", "<i2", (4,)),
    ("dirB_offset", "<i2", (4,)), ("dirB_mag", "<i2", (4,)),
    ("dirC_offset", "<i2", (4,)), ("dirC_mag", "<i2", (4,)),
], align=True)

AREA_DTYPE = np.dtype([
    ("ox", "<u4"), ("oy", "<u4"), ("rx", "<u4"), ("ry", "<u4"), ("stage", "<u4"),
    ("px_x", "<u4"), ("px_y", "<u4"), ("px_w", "<u4"), ("px_h", "<u4"),
    ("shift", "u1", (3,)), ("pad", "u1"),
    ("ditherBefore", "<u8"), ("ditherAfter", "<u8"),
    ("decomp", DECOMP_DTYPE),
], align=True)
assert DECOMP_DTYPE.itemsize == 64 and AREA_DTYPE.itemsize == 120

PLANE_ORDER = ("pDecoded", "pFactorsA", "pFactorsB", "pFactorsC", "pBlockError", "pBitsPerPixel",
               "pShiftABCX", "pColAMin", "pColAMax", "pColBMin", "pColBMax", "pColCMin", "pColCMax", "pBlockIndex")
PLANES_U8 = ("pFactorsA", "pFactorsB", "pFactorsC", "pBlockError", "pBitsPerPixel")

FLAG_FAST_BIT_CRUSH = 1
FLAG_NO_MERGE = 2
FLAG_DITHER_AES = 4

# every symbol include/limgcu.h declares (tests/test_abi.py checks the header against this list and the .so)
SYMBOLS = (
    "limgcu_create", "limgcu_destroy", "limgcu_last_error", "limgcu_device_count", "limgcu_set_rsqrt_lut",
    "limgcu_stream_handle", "limgcu_sync", "limgcu_status", "limgcu_set_dither_mode", "limgcu_set_pool_threads", "limgcu_host_has_aesni", "limgcu_launch_count", "limgcu_enable_phase_timing", "limgcu_phase_ms",
    "limgcu_debug_counters", "limgcu_debug_wave", "limgcu_debug_predicate_check", "limgcu_debug_wave_rows", "limgcu_debug_set_decode_variant", "limgcu_pass1", "limgcu_merge", "limgcu_blocked_encode3d", "limgcu_decode", "limgcu_build_block_map", "limgcu_compare",
    "limgcu_host_blocked_encode3d", "limgcu_host_encode3d", "limgcu_host_encode_stream", "limgcu_host_decode",
    "limgcu_host_pass1", "limgcu_host_merge", "limgcu_host_compare",
    "limgcu_container_bound", "limgcu_container_info", "limgcu_host_encode_container", "limgcu_host_decode_container", "limgcu_pack_payload", "limgcu_unpack_payload",
    "limgcu_batch_host_encode_containers", "limgcu_batch_host_decode_containers",
    "limgcu_area_result_words", "limgcu_encode_areas", "limgcu_finalize_rows",
)


class Planes(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in PLANE_ORDER]


class Stream(C.Structure):
    _fields_ = [("areas", C.c_void_p), ("area_count", C.c_void_p), ("block_to_area", C.c_void_p),
                ("codesA", C.c_void_p), ("codesB", C.c_void_p), ("codesC", C.c_void_p)]


class LimgError(RuntimeError):
    pass


_lib = None


def load():
    """Load liblimgcu.so and declare the signatures. Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LimgError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, sz, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
    lib.limgcu_create.argtypes = [i32, C.POINTER(vp)]
    lib.limgcu_destroy.argtypes = [vp]
    lib.limgcu_destroy.restype = None
    lib.limgcu_last_error.argtypes = [vp]
    lib.limgcu_last_error.restype = C.c_char_p
    lib.limgcu_device_count.restype = i32
    lib.limgcu_set_rsqrt_lut.argtypes = [vp, vp]
    lib.limgcu_stream_handle.argtypes = [vp]
    lib.limgcu_stream_handle.restype = vp
    lib.limgcu_sync.argtypes = [vp]
    lib.limgcu_status.argtypes = [vp]
    lib.limgcu_launch_count.argtypes = [vp]
    lib.limgcu_set_dither_mode.argtypes = [vp, i32]
    lib.limgcu_set_pool_threads.argtypes = [vp, i32]
    lib.limgcu_host_has_aesni.restype = i32
    lib.limgcu_container_bound.argtypes = [C.c_size_t, C.c_size_t, C.c_int]
    lib.limgcu_container_bound.restype = C.c_size_t
    lib.limgcu_container_info.argtypes = [vp, sz, C.POINTER(sz), C.POINTER(sz), C.POINTER(i32), C.POINTER(u32), C.POINTER(C.c_uint64)]
    lib.limgcu_host_encode_container.argtypes = [vp, vp, sz, sz, i32, u32, u32, vp, sz, C.POINTER(sz)]
    lib.limgcu_host_decode_container.argtypes = [vp, vp, sz, vp, sz]
    lib.limgcu_pack_payload.argtypes = [vp, vp, vp, u32, vp, vp, vp, vp, sz, sz, i32, vp, vp]
    lib.limgcu_unpack_payload.argtypes = [vp, vp, u32, vp, vp, vp, sz, sz, i32, vp, vp, vp]
    lib.limgcu_area_result_words.restype = sz
    lib.limgcu_encode_areas.argtypes = [vp, vp, sz, sz, i32, u32, u32, vp, vp, u32, u32, vp]
    lib.limgcu_finalize_rows.argtypes = [vp, vp, sz, sz, i32, u32, vp, vp, vp, C.POINTER(Stream), C.POINTER(Planes), sz, sz]
    lib.limgcu_batch_host_encode_containers.argtypes = [vp, i32, vp, i32, sz, sz, i32, u32, u32, vp, vp, vp]
    lib.limgcu_batch_host_decode_containers.argtypes = [vp, i32, vp, vp, i32, vp, vp]
    lib.limgcu_debug_set_decode_variant.argtypes = [vp, C.c_int]
    lib.limgcu_launch_count.restype = C.c_uint64
    lib.limgcu_enable_phase_timing.argtypes = [vp, i32]
    lib.limgcu_phase_ms.argtypes = [vp, i32]
    lib.limgcu_phase_ms.restype = C.c_float
    lib.limgcu_debug_counters.argtypes = [vp, vp]
    lib.limgcu_debug_wave.argtypes = [vp, vp]
    lib.limgcu_debug_predicate_check.argtypes = [vp, vp, sz, sz, i32, vp]
    lib.limgcu_debug_wave_rows.argtypes = [vp, vp, C.c_size_t]
    lib.limgcu_pass1.argtypes = [vp, vp, sz, sz, i32, vp]
    lib.limgcu_merge.argtypes = [vp, vp, sz, sz, i32, vp, vp, vp]
    lib.limgcu_blocked_encode3d.argtypes = [vp, vp, sz, sz, i32, u32, u32, C.POINTER(Stream), C.POINTER(Planes)]
    lib.limgcu_decode.argtypes = [vp, vp, vp, vp, vp, vp, sz, sz, i32, vp]
    lib.limgcu_build_block_map.argtypes = [vp, vp, u32, sz, sz, vp]
    lib.limgcu_compare.argtypes = [vp, vp, vp, sz, sz, i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.limgcu_host_blocked_encode3d.argtypes = [vp, vp, sz, sz, i32, C.POINTER(Planes), u32, i32]
    lib.limgcu_host_encode3d.argtypes = [vp, vp, sz, sz, i32, C.POINTER(Planes), u32, i32]
    lib.limgcu_host_encode_stream.argtypes = [vp, vp, sz, sz, i32, u32, u32, vp, C.POINTER(u32), vp, vp, vp, vp]
    lib.limgcu_host_decode.argtypes = [vp, vp, u32, vp, vp, vp, sz, sz, i32, vp]
    lib.limgcu_host_pass1.argtypes = [vp, vp, sz, sz, i32, vp]
    lib.limgcu_host_merge.argtypes = [vp, vp, sz, sz, i32, vp, C.POINTER(u32)]
    lib.limgcu_host_compare.argtypes = [vp, vp, vp, sz, sz, i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.limgcu_host_compare.restype = C.c_double
    _lib = lib
    return lib
