"""The C oracle (oracle/limg_oracle.c) against golden vectors produced by the real reference
(tools/make_golden.py). Runs anywhere: no GPU, no /root/reference."""
import numpy as np
import pytest

from oracle import oracle as lo
from tests import helpers as H


@pytest.mark.parametrize("name", H.golden_image_cases())
def test_blocked_encode_matches_reference(name):
    g = H.load_golden(name)
    alpha, ef, fast, aes = bool(g["has_alpha"]), int(g["error_factor"]), bool(g["fast"]), bool(g["aes"])
    o = lo.blocked_encode3d(g["img"], alpha, ef, fast, lo.DITHER_AES if aes else lo.DITHER_LCG)
    a = o["areas"]
    assert len(a) == g["area_rect"].shape[0]
    rect = np.stack([a["ox"], a["oy"], a["rx"], a["ry"], a["stage"]], 1)
    assert np.array_equal(rect, g["area_rect"])
    assert np.array_equal(np.stack([a["px_x"], a["px_y"], a["px_w"], a["px_h"]], 1), g["area_px"])
    assert np.array_equal(a["shift"], g["area_shift"])
    assert a["decomp"].tobytes() == H.golden_area_decomps(g).tobytes()
    assert np.array_equal(np.stack([a["ditherBefore"], a["ditherAfter"]], 1), g["area_dither"])
    for k in H.PLANES:
        assert np.array_equal(o["planes"][k], g["plane_" + k]), k
    assert not o["planes"]["pBlockError"].any()  # never written by the 3D paths (Q12)
    psnr, mse, _ = lo.compare(g["img"], o["planes"]["pDecoded"], alpha)
    assert abs(psnr - float(g["psnr"])) < 1e-9 and abs(mse - float(g["mse"])) < 1e-9


@pytest.mark.parametrize("name", H.golden_image_cases())
def test_pass1_table(name):
    g = H.load_golden(name)
    alpha = bool(g["has_alpha"])
    assert lo.pass1(g["img"], alpha).tobytes() == lo.decomp_from_ref(g["pass1"], alpha).tobytes()


@pytest.mark.parametrize("name", [n for n in H.golden_image_cases() if "enc3d_t0_pDecoded" in H.load_golden(n).files])
@pytest.mark.parametrize("threads", [0, 2])
def test_unmerged_encoder(name, threads):
    g = H.load_golden(name)
    p = lo.encode3d(g["img"], bool(g["has_alpha"]), 100, True, lo.DITHER_LCG, threads)
    for k, v in p.items():
        assert np.array_equal(v, g["enc3d_t%d_%s" % (threads, k)]), k


@pytest.mark.parametrize("name", H.golden_image_cases())
def test_decode_from_reference_streams_is_bit_exact(name):
    """Standalone decode: reference (shifts, int16 decompositions, right-aligned factors) -> reference pDecoded."""
    g = H.load_golden(name)
    alpha = bool(g["has_alpha"])
    n = g["area_rect"].shape[0]
    areas = np.zeros(n, dtype=lo.AREA_DTYPE)
    for i, k in enumerate(("px_x", "px_y", "px_w", "px_h")):
        areas[k] = g["area_px"][:, i]
    areas["shift"] = g["area_shift"]
    areas["decomp"] = H.golden_area_decomps(g)
    h, w = g["img"].shape
    out = lo.decode_areas(alpha, areas, g["post"][0], g["post"][1], g["post"][2], h, w)
    assert np.array_equal(out, g["plane_pDecoded"])


@pytest.mark.parametrize("tag", ["rgb", "rgba"])
def test_kernel_vectors(tag):
    g = H.load_golden("kernel_vectors")
    alpha = tag == "rgba"
    off = g[tag + "_offsets"]
    trials = g[tag + "_trials"]
    for t in range(len(off) - 1):
        px = g[tag + "_pixels"][off[t]:off[t + 1]]
        d = lo.fit(px, alpha)
        assert d.tobytes() == lo.decomp_from_ref(g[tag + "_fit"][t], alpha).tobytes(), t
        fa, fb, fc = lo.project(alpha, d, px)
        assert np.array_equal(fa, g[tag + "_fa"][off[t]:off[t + 1]])
        assert np.array_equal(fb, g[tag + "_fb"][off[t]:off[t + 1]])
        assert np.array_equal(fc, g[tag + "_fc"][off[t]:off[t + 1]])
        assert np.array_equal(lo.search(alpha, 100, True, d, px, fa, fb, fc), g[tag + "_shift_fast"][t])
        assert np.array_equal(lo.search(alpha, 100, False, d, px, fa, fb, fc), g[tag + "_shift_accurate"][t])
        for row in trials[trials[:, 0] == t]:
            ok, be = lo.trial(alpha, 100, d, px, fa, fb, fc, row[1:4].astype(np.uint8), 0xDEAD)
            assert (int(ok), be) == (int(row[4]), int(row[5]))
    table = lo.decomp_from_ref(g[tag + "_pred_table"], alpha)
    for i, j, want in g[tag + "_pred_pairs"]:
        assert lo.matches(alpha, table[i], table[j]) == bool(want), (i, j)


def test_dither_streams():
    g = H.load_golden("kernel_vectors")
    rows = g["dither_rows"]
    k = 0
    for mode in (lo.DITHER_LCG, lo.DITHER_AES):
        state = 0xCA7F00D15BADF00D
        for shift in (1, 3, 5, 7, 2, 4, 6):
            f, state = lo.dither(mode, shift, state, g["dither_in"])
            assert np.array_equal(f, rows[k][:61])
            assert state == int(np.frombuffer(rows[k][61:].tobytes(), np.uint64)[0])
            k += 1
    f, s = lo.dither(lo.DITHER_LCG, 8, 1234, g["dither_in"])
    assert s == 1234 and np.array_equal(f, g["dither_in"])  # shift 8: untouched (limg.cpp:801)


def test_rsqrt_table_function():
    assert lo.rsqrt(1.0) == np.float32(0.999755859375)
    assert lo.rsqrt(4.0) == np.float32(0.999755859375) / 2
    assert lo.rsqrt(0.0) == np.inf and np.isnan(lo.rsqrt(-1.0)) and lo.rsqrt(np.inf) == 0.0
    x = np.float32(3.7)
    assert lo.rsqrt(float(x) * 4.0 ** 9) == lo.rsqrt(float(x)) / 2.0 ** 9


def test_compare_constants():
    a = np.zeros((2, 2), np.uint32)
    b = np.full((2, 2), 0xFFFFFFFF, np.uint32)
    assert lo.compare(a, b, False)[2] == 585225 and lo.compare(a, b, True)[2] == 780300
