"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

from oracle import oracle as lo

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PLANES = ("pDecoded", "pFactorsA", "pFactorsB", "pFactorsC", "pBitsPerPixel", "pShiftABCX",
          "pColAMin", "pColAMax", "pColBMin", "pColBMax", "pColCMin", "pColCMax", "pBlockIndex")


def golden_image_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if "kernel_vectors" not in p)


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def golden_area_decomps(g):
    """golden area_avg/area_dec -> unified 64-byte records"""
    n = g["area_rect"].shape[0]
    out = np.zeros(n, dtype=lo.DECOMP_DTYPE)
    out["avg"] = g["area_avg"]
    for i, name in enumerate(lo.FIELDS):
        out[name] = g["area_dec"][:, i, :]
    return out


def random_pixels(rng, kind, n):
    if kind == 0:
        return rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    base = rng.integers(0, 256, 4)
    if kind == 1:
        v = base[None, :] + rng.normal(0, 6, (n, 4))
    elif kind == 2:
        d = rng.normal(0, 1, 4)
        t = rng.uniform(-40, 40, n)
        v = base[None, :] + t[:, None] * d[None, :] + rng.normal(0, 2, (n, 4))
    else:
        v = base[None, :] + rng.integers(-1, 2, (n, 4))
    v = np.clip(v, 0, 255).astype(np.uint32)
    return v[:, 0] | (v[:, 1] << 8) | (v[:, 2] << 16) | (v[:, 3] << 24)
