"""The C oracle against the compiled reference itself on seeded inputs (runs where oracle/_ref exists and the
host's RSQRTSS equals the committed table; skipped otherwise -- the golden tests still pin the oracle)."""
import numpy as np
import pytest

from limg_b200 import synth
from oracle import oracle as lo
from tests import helpers as H

IMAGES = {
    "c1_512": (lambda: synth.gradient_noise(512, 512, 1234), False),
    "rgba_256": (lambda: synth.photo_like(256, 256, 4, 4), True),
    "odd_301x203": (lambda: synth.photo_like(301, 203, 7, 3), False),
    "odd_rgba_301x203": (lambda: synth.photo_like(301, 203, 7, 4), True),
    "flatui_640x480": (lambda: synth.flat_ui(640, 480, 2, 30), False),
    "noise_rgba_64": (lambda: np.random.default_rng(0).integers(0, 2 ** 32, (64, 64), dtype=np.uint64).astype(np.uint32), True),
    "const_64": (lambda: np.full((64, 64), 0xFF102030, np.uint32), False),
    "tiny_12x20": (lambda: synth.gradient_noise(12, 20, 5), False),
}


@pytest.mark.parametrize("name", sorted(IMAGES))
@pytest.mark.parametrize("mode", ["fast", "accurate", "aes", "ef37"])
def test_full_pipeline(ref_lib, name, mode):
    factory, alpha = IMAGES[name]
    img = factory()
    ef = 37 if mode == "ef37" else 100
    fast = mode != "accurate"
    aes = mode == "aes"
    ref_lib.set_modes(True, aes)
    try:
        r = ref_lib.blocked_encode3d(img, alpha, ef, fast)
    finally:
        ref_lib.set_modes(True, False)
    o = lo.blocked_encode3d(img, alpha, ef, fast, lo.DITHER_AES if aes else lo.DITHER_LCG)
    for k in H.PLANES:
        assert np.array_equal(o["planes"][k], r[k]), k


@pytest.mark.parametrize("alpha", [False, True])
def test_fit_random_blocks(ref_lib, alpha):
    rng = np.random.default_rng(7)
    for t in range(1500):
        px = H.random_pixels(rng, t % 4, int(rng.integers(4, 200)))
        assert lo.fit(px, alpha).tobytes() == lo.decomp_from_ref(ref_lib.fit(px, alpha), alpha).tobytes(), t


@pytest.mark.parametrize("alpha", [False, True])
def test_predicate_all_neighbour_pairs(ref_lib, alpha):
    img = synth.photo_like(256, 192, 31, 4 if alpha else 3)
    rt = ref_lib.pass1(img, alpha)
    t = lo.decomp_from_ref(rt, alpha)
    rng = np.random.default_rng(3)
    for i in range(len(t)):
        for j in (i + 1, i + 32, int(rng.integers(0, len(t)))):
            if j < len(t):
                assert lo.matches(alpha, t[i], t[j]) == ref_lib.matches(alpha, rt[i], rt[j]), (i, j)


def test_1080p_photo(ref_lib):
    img = synth.frame(0)
    r = ref_lib.blocked_encode3d(img, False)
    o = lo.blocked_encode3d(img, False)
    for k in H.PLANES:
        assert np.array_equal(o["planes"][k], r[k]), k


def test_unmerged_encoder(ref_lib):
    img = synth.photo_like(200, 136, 8, 3)
    for threads in (0, 1, 3):
        r = ref_lib.encode3d(img, False, 100, True, threads)
        o = lo.encode3d(img, False, 100, True, lo.DITHER_LCG, threads)
        for k, v in r.items():
            assert np.array_equal(o[k], v), (threads, k)
