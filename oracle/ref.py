"""ctypes binding of oracle/_ref/libref.so (the unmodified reference compiled by oracle/Makefile).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. Never imported by limg_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")

AREA_DTYPE = np.dtype([
    ("ox", "<u4"), ("oy", "<u4"), ("rx", "<u4"), ("ry", "<u4"), ("stage", "<u4"),
    ("px_x", "<u4"), ("px_y", "<u4"), ("px_w", "<u4"), ("px_h", "<u4"),
    ("shift", "u1", (3,)), ("pad", "u1"),
    ("ditherBefore", "<u8"), ("ditherAfter", "<u8"),
    ("avg", "<f4", (4,)), ("dec", "<i2", (6, 4)),
], align=True)

PLANES_U32 = ("pDecoded", "pShiftABCX", "pColAMin", "pColAMax", "pColBMin", "pColBMax", "pColCMin", "pColCMax", "pBlockIndex")
PLANES_U8 = ("pFactorsA", "pFactorsB", "pFactorsC", "pBitsPerPixel")


def available() -> bool:
    return os.path.exists(LIB_PATH)


def build() -> bool:
    """(Re)build libref.so when the reference sources are present (this container only)."""
    subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=False, capture_output=True)
    return available()


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref.so is missing; run `make -C oracle ref` where /root/reference exists")
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_compare.restype = C.c_double
        _lib.ref_time_blocked.restype = C.c_double
        _lib.ref_rsqrtss.restype = C.c_float
        _lib.ref_rsqrtss.argtypes = [C.c_float]
        _lib.ref_dither.restype = C.c_uint64
        _lib.ref_blocked_trace.restype = C.c_int64
        assert _lib.ref_sizeof_area() == AREA_DTYPE.itemsize, (_lib.ref_sizeof_area(), AREA_DTYPE.itemsize)
    return _lib


def set_modes(sse41: bool = True, aesni: bool = False) -> None:
    lib().ref_set_modes(int(sse41), int(aesni))


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _img(img):
    img = np.ascontiguousarray(img, dtype=np.uint32)
    assert img.ndim == 2
    return img


def alloc_planes(h: int, w: int) -> dict:
    out = {k: np.zeros((h, w), dtype=np.uint32) for k in PLANES_U32}
    out.update({k: np.zeros((h, w), dtype=np.uint8) for k in PLANES_U8})
    out["pBlockError"] = np.zeros((h, w), dtype=np.uint8)
    return out


def blocked_encode3d(img, has_alpha: bool, error_factor: int = 100, fast: bool = True, threads: int = 0) -> dict:
    """limg_blocked_encode3d_test (limg.h:46) through the real public entry point."""
    img = _img(img)
    h, w = img.shape
    p = alloc_planes(h, w)
    rc = lib().ref_blocked_encode3d(_vp(img), C.c_size_t(w), C.c_size_t(h), int(has_alpha), C.c_uint32(error_factor), int(fast), int(threads),
                                    _vp(p["pDecoded"]), _vp(p["pFactorsA"]), _vp(p["pFactorsB"]), _vp(p["pFactorsC"]), _vp(p["pBlockError"]), _vp(p["pBitsPerPixel"]),
                                    _vp(p["pShiftABCX"]), _vp(p["pColAMin"]), _vp(p["pColAMax"]), _vp(p["pColBMin"]), _vp(p["pColBMax"]), _vp(p["pColCMin"]), _vp(p["pColCMax"]), _vp(p["pBlockIndex"]))
    assert rc == 0, rc
    return p


def encode3d(img, has_alpha: bool, error_factor: int = 100, fast: bool = True, threads: int = 0) -> dict:
    """limg_encode3d_test (limg.h:35): every 8x8 block is its own area."""
    img = _img(img)
    h, w = img.shape
    p = alloc_planes(h, w)
    rc = lib().ref_encode3d(_vp(img), C.c_size_t(w), C.c_size_t(h), int(has_alpha), C.c_uint32(error_factor), int(fast), int(threads),
                            _vp(p["pDecoded"]), _vp(p["pFactorsA"]), _vp(p["pFactorsB"]), _vp(p["pFactorsC"]),
                            _vp(p["pShiftABCX"]), _vp(p["pColAMin"]), _vp(p["pColAMax"]), _vp(p["pColBMin"]), _vp(p["pColBMax"]), _vp(p["pColCMin"]), _vp(p["pColCMax"]))
    assert rc == 0, rc
    for k in ("pBlockError", "pBitsPerPixel", "pBlockIndex"):
        p.pop(k)
    return p


def compare(a, b, has_alpha: bool):
    a = _img(a); b = _img(b)
    mse = C.c_double(); mx = C.c_double()
    psnr = lib().ref_compare(_vp(a), _vp(b), C.c_size_t(a.shape[1]), C.c_size_t(a.shape[0]), int(has_alpha), C.byref(mse), C.byref(mx))
    return psnr, mse.value, mx.value


def time_blocked(img, has_alpha: bool, error_factor: int = 100, fast: bool = True, threads: int = 0, reps: int = 1, perf_path: bool = False) -> float:
    img = _img(img)
    return lib().ref_time_blocked(_vp(img), C.c_size_t(img.shape[1]), C.c_size_t(img.shape[0]), int(has_alpha), C.c_uint32(error_factor), int(fast), int(threads), int(reps), int(perf_path))


def rec_size(has_alpha: bool) -> int:
    return 64 if has_alpha else 48


def pass1(img, has_alpha: bool) -> np.ndarray:
    img = _img(img)
    h, w = img.shape
    bx, by = (w + 7) // 8, (h + 7) // 8
    table = np.zeros((by * bx, rec_size(has_alpha) + 16), dtype=np.uint8)  # slack: the RGB SSE store writes 16 B at avg
    flat = np.zeros(by * bx * rec_size(has_alpha) + 64, dtype=np.uint8)
    lib().ref_pass1(_vp(img), C.c_size_t(w), C.c_size_t(h), int(has_alpha), _vp(flat))
    del table
    return flat[: by * bx * rec_size(has_alpha)].reshape(by * bx, rec_size(has_alpha)).copy()


def fit(pixels, has_alpha: bool) -> np.ndarray:
    pixels = np.ascontiguousarray(pixels, dtype=np.uint32).ravel()
    out = np.zeros(64, dtype=np.uint8)
    lib().ref_fit(_vp(pixels), C.c_size_t(pixels.size), int(has_alpha), _vp(out))
    return out[: rec_size(has_alpha)].copy()


def matches(has_alpha: bool, a, b) -> bool:
    a = np.ascontiguousarray(a, dtype=np.uint8); b = np.ascontiguousarray(b, dtype=np.uint8)
    return bool(lib().ref_matches(int(has_alpha), _vp(a), _vp(b)))


def project(has_alpha: bool, decomp, pixels):
    pixels = np.ascontiguousarray(pixels, dtype=np.uint32).ravel()
    decomp = np.ascontiguousarray(decomp, dtype=np.uint8)
    n = pixels.size
    a = np.zeros(n, np.uint8); b = np.zeros(n, np.uint8); c = np.zeros(n, np.uint8)
    lib().ref_project(int(has_alpha), _vp(decomp), _vp(pixels), C.c_size_t(n), _vp(a), _vp(b), _vp(c))
    return a, b, c


def trial(has_alpha: bool, error_factor: int, decomp, pixels, fa, fb, fc, shift, block_error_in: int = 0):
    pixels = np.ascontiguousarray(pixels, dtype=np.uint32).ravel()
    decomp = np.ascontiguousarray(decomp, dtype=np.uint8)
    sh = np.ascontiguousarray(shift, dtype=np.uint8)
    be = C.c_uint64(block_error_in)
    ok = lib().ref_trial(int(has_alpha), C.c_uint32(error_factor), _vp(decomp), _vp(pixels), C.c_size_t(pixels.size), _vp(fa), _vp(fb), _vp(fc), _vp(sh), C.byref(be))
    return bool(ok), be.value


def search(has_alpha: bool, error_factor: int, fast: bool, decomp, pixels, fa, fb, fc) -> np.ndarray:
    pixels = np.ascontiguousarray(pixels, dtype=np.uint32).ravel()
    decomp = np.ascontiguousarray(decomp, dtype=np.uint8)
    sh = np.zeros(3, np.uint8)
    lib().ref_search(int(has_alpha), C.c_uint32(error_factor), int(fast), _vp(decomp), _vp(pixels), C.c_size_t(pixels.size), _vp(fa), _vp(fb), _vp(fc), _vp(sh))
    return sh


def dither(shift: int, state: int, factors):
    f = np.ascontiguousarray(factors, dtype=np.uint8).copy()
    new = lib().ref_dither(C.c_uint8(shift), C.c_size_t(f.size), C.c_uint64(state), _vp(f))
    return f, int(new)


def decode(has_alpha: bool, decomp, shift, fa, fb, fc, rx: int, ry: int) -> np.ndarray:
    decomp = np.ascontiguousarray(decomp, dtype=np.uint8)
    sh = np.ascontiguousarray(shift, dtype=np.uint8)
    out = np.zeros((ry, rx), np.uint32)
    lib().ref_decode(int(has_alpha), _vp(out), C.c_size_t(rx), C.c_size_t(rx), C.c_size_t(ry), _vp(fa), _vp(fb), _vp(fc), _vp(decomp), _vp(sh))
    return out


def blocked_trace(img, has_alpha: bool, error_factor: int = 100, fast: bool = True) -> dict:
    """Harness-driven three-stage run that also returns the area list, decompositions, shifts and factor streams."""
    img = _img(img)
    h, w = img.shape
    bx, by = (w + 7) // 8, (h + 7) // 8
    p = alloc_planes(h, w)
    areas = np.zeros(bx * by, dtype=AREA_DTYPE)
    pre = [np.zeros(h * w, np.uint8) for _ in range(3)]
    post = [np.zeros(h * w, np.uint8) for _ in range(3)]
    table = np.zeros(bx * by * rec_size(has_alpha) + 64, np.uint8)
    n = lib().ref_blocked_trace(_vp(img), C.c_size_t(w), C.c_size_t(h), int(has_alpha), C.c_uint32(error_factor), int(fast),
                                _vp(p["pDecoded"]), _vp(p["pFactorsA"]), _vp(p["pFactorsB"]), _vp(p["pFactorsC"]), _vp(p["pBitsPerPixel"]),
                                _vp(p["pShiftABCX"]), _vp(p["pColAMin"]), _vp(p["pColAMax"]), _vp(p["pColBMin"]), _vp(p["pColBMax"]), _vp(p["pColCMin"]), _vp(p["pColCMax"]), _vp(p["pBlockIndex"]),
                                _vp(areas), C.c_size_t(areas.size), _vp(pre[0]), _vp(pre[1]), _vp(pre[2]), _vp(post[0]), _vp(post[1]), _vp(post[2]), _vp(table))
    assert n >= 0
    return {"planes": p, "areas": areas[:n].copy(), "pre": pre, "post": post,
            "pass1": table[: bx * by * rec_size(has_alpha)].reshape(bx * by, rec_size(has_alpha)).copy()}
