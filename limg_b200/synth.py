"""Synthetic inputs of BASELINE.json's configs.

All images are uint32 RGBA8 (R in the low byte, reference layout limg_internal.h:214-223),
row-major, stride = width. The draw order of the numpy Generator calls is fixed; changing it
changes every golden fixture under tests/golden/.
"""
from __future__ import annotations

import numpy as np


def pack_rgba(r, g, b, a=None) -> np.ndarray:
    r = np.asarray(r, dtype=np.uint32)
    g = np.asarray(g, dtype=np.uint32)
    b = np.asarray(b, dtype=np.uint32)
    a = np.full_like(r, 255) if a is None else np.asarray(a, dtype=np.uint32)
    return np.ascontiguousarray(r | (g << 8) | (b << 16) | (a << 24)).astype(np.uint32)


def _clip_u8(x) -> np.ndarray:
    return np.clip(np.rint(x), 0, 255).astype(np.uint32)


def gradient_noise(w: int = 512, h: int = 512, seed: int = 1234, sigma: float = 6.0) -> np.ndarray:
    """C1: R=255x/(W-1), G=255y/(H-1), B=255(x+y)/(W+H-2) + N(0, sigma) per channel."""
    rng = np.random.default_rng(seed)
    x = np.arange(w, dtype=np.float64)[None, :]
    y = np.arange(h, dtype=np.float64)[:, None]
    r = 255.0 * x / max(w - 1, 1) + 0 * y
    g = 255.0 * y / max(h - 1, 1) + 0 * x
    b = 255.0 * (x + y) / max(w + h - 2, 1)
    r = r + rng.normal(0.0, sigma, (h, w))
    g = g + rng.normal(0.0, sigma, (h, w))
    b = b + rng.normal(0.0, sigma, (h, w))
    return pack_rgba(_clip_u8(r), _clip_u8(g), _clip_u8(b))


def photo_like(w: int = 3840, h: int = 2160, seed: int = 1, channels: int = 3, sigma: float = 4.0) -> np.ndarray:
    """C2/C3/C5: per channel 128 + sum_{k<4} U(10,40) sin(2pi(fx x/W + fy y/H) + phi) + N(0, sigma)."""
    rng = np.random.default_rng(seed)
    x = (np.arange(w, dtype=np.float32) / np.float32(w))[None, :]
    y = (np.arange(h, dtype=np.float32) / np.float32(h))[:, None]
    planes = []
    for _ in range(channels):
        acc = np.full((h, w), 128.0, dtype=np.float32)
        for _k in range(4):
            amp = np.float32(rng.uniform(10.0, 40.0))
            fx = np.float32(rng.uniform(0.5, 6.0))
            fy = np.float32(rng.uniform(0.5, 6.0))
            phi = np.float32(rng.uniform(0.0, 2.0 * np.pi))
            acc += amp * np.sin(np.float32(2.0 * np.pi) * (fx * x + fy * y) + phi)
        acc += rng.normal(0.0, sigma, (h, w)).astype(np.float32)
        planes.append(_clip_u8(acc))
    if channels == 3:
        return pack_rgba(planes[0], planes[1], planes[2])
    return pack_rgba(planes[0], planes[1], planes[2], planes[3])


def flat_ui(w: int = 3840, h: int = 2160, seed: int = 2, rects: int = 400) -> np.ndarray:
    """C4: background 240, `rects` uniformly placed rectangles w~U{16..599}, h~U{16..299}, uniform RGB."""
    rng = np.random.default_rng(seed)
    img = np.full((h, w, 3), 240, dtype=np.uint32)
    for _ in range(rects):
        x0 = int(rng.integers(0, w))
        y0 = int(rng.integers(0, h))
        rw = int(rng.integers(16, 600))
        rh = int(rng.integers(16, 300))
        col = rng.integers(0, 256, 3)
        img[y0:min(h, y0 + rh), x0:min(w, x0 + rw), :] = col
    return pack_rgba(img[..., 0], img[..., 1], img[..., 2])


def frame(index: int, w: int = 1920, h: int = 1080) -> np.ndarray:
    """C5: frame `index` of the 1080p batch (photo_like, seed 3 + index)."""
    return photo_like(w, h, seed=3 + index, channels=3)


CONFIGS = {
    "c1_512_gradient": lambda: (gradient_noise(512, 512, 1234), False),
    "c2_4k_photo": lambda: (photo_like(3840, 2160, 1, 3), False),
    "c3_8k_rgba": lambda: (photo_like(7680, 4320, 4, 4), True),
    "c4_4k_flatui": lambda: (flat_ui(3840, 2160, 2), False),
    "c5_1080p_frame0": lambda: (frame(0), False),
}
