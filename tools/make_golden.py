"""Generate tests/golden/*.npz from the REAL reference (oracle/_ref/libref.so, built from /root/reference).

Run in the build container only (the reference sources do not exist on the GPU box):
    python tools/make_golden.py
The fixtures pin the C oracle (tests/test_oracle_golden.py) and, through it, the CUDA path. The reference
is run on its SSE4.1 path with the LCG dither unless a case says "aes".
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from limg_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> (image factory, has_alpha, error_factor, fast, aes)
IMAGE_CASES = {
    "rgb_photo_96x64": (lambda: synth.photo_like(96, 64, 11, 3), False, 100, True, False),
    "rgba_photo_64x64": (lambda: synth.photo_like(64, 64, 12, 4), True, 100, True, False),
    "rgb_flatui_128x96": (lambda: synth.flat_ui(128, 96, 13, 12), False, 100, True, False),
    "rgb_gradient_100x70": (lambda: synth.gradient_noise(100, 70, 14), False, 100, True, False),
    "rgb_photo_96x64_accurate": (lambda: synth.photo_like(96, 64, 11, 3), False, 100, False, False),
    "rgb_photo_96x64_aes": (lambda: synth.photo_like(96, 64, 11, 3), False, 100, True, True),
    "rgb_photo_96x64_ef0": (lambda: synth.photo_like(96, 64, 11, 3), False, 0, True, False),
    "rgb_photo_96x64_ef30": (lambda: synth.photo_like(96, 64, 11, 3), False, 30, True, False),
    "rgba_gradient_72x40_accurate": (lambda: synth.photo_like(72, 40, 15, 4), True, 100, False, False),
    "rgb_smooth_160x120": (lambda: synth.photo_like(160, 120, 16, 3, sigma=1.0), False, 100, True, False),
}


def image_case(name, factory, alpha, ef, fast, aes):
    img = factory()
    ref.set_modes(True, aes)
    tr = ref.blocked_trace(img, alpha, ef, fast)
    real = ref.blocked_encode3d(img, alpha, ef, fast)
    for k, v in real.items():
        if k != "pBlockError":
            assert np.array_equal(v, tr["planes"][k]), (name, k)
    a = tr["areas"]
    out = {"img": img, "has_alpha": alpha, "error_factor": ef, "fast": fast, "aes": aes,
           "pass1": tr["pass1"],
           "area_rect": np.stack([a["ox"], a["oy"], a["rx"], a["ry"], a["stage"]], 1).astype(np.uint32),
           "area_px": np.stack([a["px_x"], a["px_y"], a["px_w"], a["px_h"]], 1).astype(np.uint32),
           "area_shift": a["shift"].copy(), "area_avg": a["avg"].copy(), "area_dec": a["dec"].copy(),
           "area_dither": np.stack([a["ditherBefore"], a["ditherAfter"]], 1),
           "pre": np.stack(tr["pre"]), "post": np.stack(tr["post"])}
    for k, v in tr["planes"].items():
        if k != "pBlockError":
            out["plane_" + k] = v
    psnr, mse, mx = ref.compare(img, tr["planes"]["pDecoded"], alpha)
    out["psnr"] = np.float64(psnr)
    out["mse"] = np.float64(mse)
    # the non-merged encoder on the same input (pool-less and with a 2-thread pool: 8 y-bands)
    if not aes and fast and ef == 100:
        for threads in (0, 2):
            e = ref.encode3d(img, alpha, ef, fast, threads)
            for k, v in e.items():
                out["enc3d_t%d_%s" % (threads, k)] = v
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, img.shape, "areas", len(a), "psnr %.4f" % psnr)


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


def kernel_vectors():
    """Per-function known-answer vectors (fit, projection, trial, search, predicate, dither, decode)."""
    rng = np.random.default_rng(2024)
    ref.set_modes(True, False)
    out = {}
    for alpha in (False, True):
        tag = "rgba" if alpha else "rgb"
        pix, off, fits, fa, fb, fc, shifts, shifts_acc, trials = [], [0], [], [], [], [], [], [], []
        for t in range(96):
            n = 64 if t % 3 == 0 else int(rng.integers(4, 300))
            px = random_pixels(rng, t % 4, n)
            rec = ref.fit(px, alpha)
            a, b, c = ref.project(alpha, rec, px)
            pix.append(px); off.append(off[-1] + n); fits.append(rec)
            fa.append(a); fb.append(b); fc.append(c)
            shifts.append(ref.search(alpha, 100, True, rec, px, a, b, c))
            shifts_acc.append(ref.search(alpha, 100, False, rec, px, a, b, c))
            for sh in ((4, 5, 6), (2, 4, 5), (0, 0, 1), (8, 8, 8), (5, 8, 8), (1, 7, 3)):
                ok, be = ref.trial(alpha, 100, rec, px, a, b, c, sh, 0xDEAD)
                trials.append((t, sh[0], sh[1], sh[2], int(ok), be))
        out[tag + "_pixels"] = np.concatenate(pix); out[tag + "_offsets"] = np.array(off, np.int64)
        out[tag + "_fit"] = np.stack(fits)
        out[tag + "_fa"] = np.concatenate(fa); out[tag + "_fb"] = np.concatenate(fb); out[tag + "_fc"] = np.concatenate(fc)
        out[tag + "_shift_fast"] = np.stack(shifts); out[tag + "_shift_accurate"] = np.stack(shifts_acc)
        out[tag + "_trials"] = np.array(trials, np.uint64)
        # merge predicate on pairs of fits drawn from a real pass-1 table (neighbours + random pairs)
        img = synth.photo_like(160, 120, 21, 4 if alpha else 3)
        table = ref.pass1(img, alpha)
        bx = 20
        pairs = []
        for i in range(table.shape[0]):
            for j in (i + 1, i + bx, int(rng.integers(0, table.shape[0]))):
                if j < table.shape[0]:
                    pairs.append((i, j, int(ref.matches(alpha, table[i], table[j]))))
        out[tag + "_pred_table"] = table
        out[tag + "_pred_pairs"] = np.array(pairs, np.int32)
    # dither streams, both generators
    f = (np.arange(61) * 17 % 256).astype(np.uint8)
    rows = []
    for aes in (False, True):
        ref.set_modes(True, aes)
        state = 0xCA7F00D15BADF00D
        for shift in (1, 3, 5, 7, 2, 4, 6):
            g, state = ref.dither(shift, state, f)
            rows.append(np.concatenate([g, np.frombuffer(np.uint64(state).tobytes(), np.uint8)]))
    out["dither_in"] = f
    out["dither_rows"] = np.stack(rows)
    ref.set_modes(True, False)
    np.savez_compressed(os.path.join(OUT, "kernel_vectors.npz"), **out)
    print("kernel_vectors", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    assert ref.build(), "reference harness could not be built"
    for name, (factory, alpha, ef, fast, aes) in IMAGE_CASES.items():
        image_case(name, factory, alpha, ef, fast, aes)
    kernel_vectors()
