"""ctypes binding of oracle/liblimg_oracle.so (the plain-C restatement, limg_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
Never imported by limg_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblimg_oracle.so")

DECOMP_DTYPE = np.dtype([
    ("avg", "<f4", (4,)),
    ("dirA_min", "<i2", (4,)), ("dirA_max", "<i2", (4,)),
    ("dirB_offset", "<i2", (4,)), ("dirB_mag", "<i2", (4,)),
    ("dirC_offset", "<i2", (4,)), ("dirC_mag", "<i2", (4,)),
], align=True)
assert DECOMP_DTYPE.itemsize == 64

AREA_DTYPE = np.dtype([
    ("ox", "<u4"), ("oy", "<u4"), ("rx", "<u4"), ("ry", "<u4"), ("stage", "<u4"),
    ("px_x", "<u4"), ("px_y", "<u4"), ("px_w", "<u4"), ("px_h", "<u4"),
    ("shift", "u1", (3,)), ("pad", "u1"),
    ("ditherBefore", "<u8"), ("ditherAfter", "<u8"),
    ("decomp", DECOMP_DTYPE),
], align=True)

PLANE_ORDER = ("pDecoded", "pFactorsA", "pFactorsB", "pFactorsC", "pBlockError", "pBitsPerPixel",
               "pShiftABCX", "pColAMin", "pColAMax", "pColBMin", "pColBMax", "pColCMin", "pColCMax", "pBlockIndex")
PLANES_U8 = ("pFactorsA", "pFactorsB", "pFactorsC", "pBlockError", "pBitsPerPixel")

DITHER_LCG, DITHER_AES = 0, 1
FIELDS = ("dirA_min", "dirA_max", "dirB_offset", "dirB_mag", "dirC_offset", "dirC_mag")


class Planes(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in PLANE_ORDER]


def build() -> None:
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.lo_rsqrt.restype = C.c_float
        _lib.lo_rsqrt.argtypes = [C.c_float]
        _lib.lo_dither.restype = C.c_uint64
        _lib.lo_merge.restype = C.c_size_t
        _lib.lo_blocked_encode3d.restype = C.c_size_t
        _lib.lo_compare.restype = C.c_double
        _lib.lo_counter.restype = C.c_uint64
    return _lib


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _sz(v):
    return C.c_size_t(int(v))


def channels(has_alpha: bool) -> int:
    return 4 if has_alpha else 3


def decomp_from_ref(rec: np.ndarray, has_alpha: bool) -> np.ndarray:
    """Reference record (48 B RGB / 64 B RGBA, limg_internal.h:343-353) -> unified 64-byte record(s)."""
    rec = np.ascontiguousarray(rec, dtype=np.uint8)
    ch = channels(has_alpha)
    flat = rec.reshape(-1, 16 * ch)
    out = np.zeros(flat.shape[0], dtype=DECOMP_DTYPE)
    out["avg"][:, :ch] = flat[:, : 4 * ch].copy().view("<f4")
    ints = flat[:, 4 * ch:].copy().view("<i2").reshape(-1, 6, ch)
    for i, name in enumerate(FIELDS):
        out[name][:, :ch] = ints[:, i, :]
    return out if rec.ndim > 1 else out[0]


def decomp_from_ref_area(area) -> np.ndarray:
    """oracle.ref AREA_DTYPE entry (avg + dec[6][4]) -> unified record."""
    out = np.zeros((), dtype=DECOMP_DTYPE)
    out["avg"] = area["avg"]
    for i, name in enumerate(FIELDS):
        out[name] = area["dec"][i]
    return out


def rsqrt(x: float) -> float:
    return lib().lo_rsqrt(C.c_float(x))


def fit(pixels, has_alpha: bool) -> np.ndarray:
    pixels = np.ascontiguousarray(pixels, dtype=np.uint32).ravel()
    out = np.zeros((), dtype=DECOMP_DTYPE)
    lib().lo_fit(_vp(pixels), _sz(pixels.size), channels(has_alpha), _vp(out))
    return out


def matches(has_alpha: bool, a, b) -> bool:
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return bool(lib().lo_matches(channels(has_alpha), _vp(a), _vp(b)))


def project(has_alpha: bool, d, pixels):
    pixels = np.ascontiguousarray(pixels, dtype=np.uint32).ravel()
    d = np.ascontiguousarray(d)
    n = pixels.size
    fa = np.zeros(n, np.uint8); fb = np.zeros(n, np.uint8); fc = np.zeros(n, np.uint8)
    lib().lo_project(channels(has_alpha), _vp(d), _vp(pixels), _sz(n), _vp(fa), _vp(fb), _vp(fc))
    return fa, fb, fc


def thresholds(error_factor: int):
    return 0x6 * (error_factor // 2) * 7, 0x4 * (error_factor // 2) * 7


def trial(has_alpha: bool, error_factor: int, d, pixels, fa, fb, fc, shift, block_error_in: int = 0):
    pixels = np.ascontiguousarray(pixels, dtype=np.uint32).ravel()
    d = np.ascontiguousarray(d)
    sh = np.ascontiguousarray(shift, dtype=np.uint8)
    mp, mb = thresholds(error_factor)
    be = C.c_uint64(block_error_in)
    ok = lib().lo_trial(channels(has_alpha), C.c_uint64(mp), C.c_uint64(mb), _vp(d), _vp(pixels), _sz(pixels.size), _vp(fa), _vp(fb), _vp(fc), _vp(sh), C.byref(be))
    return bool(ok), be.value


def search(has_alpha: bool, error_factor: int, fast: bool, d, pixels, fa, fb, fc) -> np.ndarray:
    pixels = np.ascontiguousarray(pixels, dtype=np.uint32).ravel()
    d = np.ascontiguousarray(d)
    sh = np.zeros(3, np.uint8)
    lib().lo_search(channels(has_alpha), C.c_uint32(error_factor), int(fast), _vp(d), _vp(pixels), _sz(pixels.size), _vp(fa), _vp(fb), _vp(fc), _vp(sh))
    return sh


def dither(mode: int, shift: int, state: int, factors):
    f = np.ascontiguousarray(factors, dtype=np.uint8).copy()
    new = lib().lo_dither(int(mode), C.c_uint8(shift), _sz(f.size), C.c_uint64(state), _vp(f))
    return f, int(new)


def decode(has_alpha: bool, d, shift, fa, fb, fc, rx: int, ry: int) -> np.ndarray:
    d = np.ascontiguousarray(d)
    sh = np.ascontiguousarray(shift, dtype=np.uint8)
    out = np.zeros((ry, rx), np.uint32)
    lib().lo_decode(channels(has_alpha), _vp(out), _sz(rx), _sz(rx), _sz(ry), _vp(fa), _vp(fb), _vp(fc), _vp(d), _vp(sh))
    return out


def pass1(img, has_alpha: bool) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint32)
    h, w = img.shape
    table = np.zeros(((h + 7) // 8) * ((w + 7) // 8), dtype=DECOMP_DTYPE)
    lib().lo_pass1(_vp(img), _sz(w), _sz(h), channels(has_alpha), _vp(table))
    return table


def merge(table, bx: int, by: int, has_alpha: bool):
    table = np.ascontiguousarray(table)
    areas = np.zeros(max(bx * by, 1), dtype=AREA_DTYPE)
    stats = np.zeros(8, np.uint64)
    n = lib().lo_merge(_vp(table), _sz(bx), _sz(by), channels(has_alpha), _vp(areas), _vp(stats))
    return areas[:n].copy(), stats


def alloc_planes(h: int, w: int) -> dict:
    return {k: np.zeros((h, w), dtype=np.uint8 if k in PLANES_U8 else np.uint32) for k in PLANE_ORDER}


def _planes_struct(p: dict) -> Planes:
    s = Planes()
    for k in PLANE_ORDER:
        a = p.get(k)
        setattr(s, k, a.ctypes.data if a is not None else None)
    return s


def blocked_encode3d(img, has_alpha: bool, error_factor: int = 100, fast: bool = True, dither_mode: int = DITHER_LCG) -> dict:
    img = np.ascontiguousarray(img, dtype=np.uint32)
    h, w = img.shape
    p = alloc_planes(h, w)
    s = _planes_struct(p)
    areas = np.zeros(max(((h + 7) // 8) * ((w + 7) // 8), 1), dtype=AREA_DTYPE)
    n = lib().lo_blocked_encode3d(_vp(img), _sz(w), _sz(h), int(has_alpha), C.c_uint32(error_factor), int(fast), int(dither_mode), C.byref(s), _vp(areas))
    return {"planes": p, "areas": areas[:n].copy()}


def encode3d(img, has_alpha: bool, error_factor: int = 100, fast: bool = True, dither_mode: int = DITHER_LCG, pool_threads: int = 0) -> dict:
    img = np.ascontiguousarray(img, dtype=np.uint32)
    h, w = img.shape
    p = alloc_planes(h, w)
    s = _planes_struct(p)
    lib().lo_encode3d(_vp(img), _sz(w), _sz(h), int(has_alpha), C.c_uint32(error_factor), int(fast), int(dither_mode), _sz(pool_threads), C.byref(s))
    for k in ("pBlockError", "pBitsPerPixel", "pBlockIndex"):
        p.pop(k)
    return p


def decode_areas(has_alpha: bool, areas, fa, fb, fc, h: int, w: int) -> np.ndarray:
    areas = np.ascontiguousarray(areas)
    out = np.zeros((h, w), np.uint32)
    lib().lo_decode_areas(channels(has_alpha), _vp(areas), _sz(areas.size), _vp(fa), _vp(fb), _vp(fc), _vp(out), _sz(w))
    return out


def compare(a, b, has_alpha: bool):
    a = np.ascontiguousarray(a, dtype=np.uint32); b = np.ascontiguousarray(b, dtype=np.uint32)
    mse = C.c_double(); mx = C.c_double()
    psnr = lib().lo_compare(_vp(a), _vp(b), _sz(a.shape[1]), _sz(a.shape[0]), int(has_alpha), C.byref(mse), C.byref(mx))
    return psnr, mse.value, mx.value


def counter(which: int) -> int:
    return int(lib().lo_counter(int(which)))
