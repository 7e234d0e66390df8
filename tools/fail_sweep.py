"""First-try verification failures of the merge scan over many synthetic frames (different seeds, noise levels, flat-UI densities, sizes)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from limg_b200 import Codec, synth
c = Codec(0)
cases = []
for s in range(12):
    cases.append(("photo 1920x1080 seed %d sigma %.1f" % (s, 1.0 + s % 4 * 2), lambda s=s: synth.photo_like(1920, 1080, 100 + s, 3, sigma=1.0 + s % 4 * 2), False))
for s in range(6):
    cases.append(("rgba 1280x720 seed %d" % s, lambda s=s: synth.photo_like(1280, 720, 200 + s, 4), True))
for s in range(10):
    cases.append(("flat-ui 1920x1080 seed %d rects %d" % (s, 50 + 80 * s), lambda s=s: synth.flat_ui(1920, 1080, 300 + s, 50 + 80 * s), False))
for s in range(4):
    cases.append(("gradient 1024x768 seed %d" % s, lambda s=s: synth.gradient_noise(1024, 768, 400 + s, sigma=2.0 + 3 * s), False))
cases.append(("flat-ui 3840x2160 rects 150", lambda: synth.flat_ui(3840, 2160, 500, 150), False))
cases.append(("photo 3840x2160 smooth", lambda: synth.photo_like(3840, 2160, 501, 3, sigma=1.0), False))
total = 0
ITER = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ONLY = sys.argv[2] if len(sys.argv) > 2 else ""
cases = [k for k in cases if ONLY in k[0]]
for name, make, alpha in cases:
    img = make()
    h, w = img.shape
    d = torch.from_numpy(img.view(np.int32)).cuda()
    codes = [torch.empty((h, w), dtype=torch.uint8, device="cuda") for _ in range(3)]
    st = {"codesA": codes[0].data_ptr(), "codesB": codes[1].data_ptr(), "codesC": codes[2].data_ptr()}
    fails = []
    o = None
    for it in range(ITER):
        c.blocked_encode3d_device(d.data_ptr(), w, h, alpha, 100, True, False, st, None)
        c.sync()
        fails.append(int(c.debug_counters()[24]))
        if fails[-1] and o is None:
            dbg = c.debug_wave()
            o = dbg[104:112]
            why = {1: "corner busy but count differs", 2: "replay emits nothing", 3: "replay emits more", 4: "different rectangle", 5: "a block of the rectangle has another owner", 6: "recorded more"}
            for k in range(min(int(dbg[100]), 8)):
                q = dbg[104 + 8 * k: 112 + 8 * k]
                wx, wy = dbg[232 + 2 * k], dbg[233 + 2 * k]
                if int(q[7]) == 5:
                    bx = (w + 7) // 8
                    own = int(wx)
                    st1 = own >= 0x40000000
                    sd = ((own & 0x3FFFFFFF) >> 3)
                    print("     stage %d seed (%d, %d) rectangle (%d,%d %dx%d): a block is owned by stage %d seed (%d, %d) attempt %d whose first rectangle is (%d,%d %dx%d)" % (
                        q[0] & 255, q[1], q[2], q[5] & 0xFFFF, q[5] >> 16, q[6] & 0xFFFF, q[6] >> 16, int(st1), sd % bx, sd // bx, own & 7, wy & 0xFFFF, wy >> 16, q[4] & 0xFFFF, q[4] >> 16))
                    continue
                print("     stage %d seed (%d, %d) recorded %d replayed %d first recorded (%d,%d %dx%d) replay wants (%d,%d %dx%d): %s" % (
                    q[0] & 255, q[1], q[2], q[3], q[4], q[5] & 0xFFFF, q[5] >> 16, q[6] & 0xFFFF, q[6] >> 16, wx & 0xFFFF, wx >> 16, wy & 0xFFFF, wy >> 16, why.get(int(q[7]), "?")))
    total += sum(1 for f in fails if f)
    if any(fails):
        print("FAIL %s: tries %s; first failing seed: stage %d (%d, %d) recorded %d rect0 ox %d oy %d rx %d ry %d" % (name, fails, o[0] & 255, o[1], o[2], o[3], o[5] & 0xFFFF, o[5] >> 16, o[6] & 0xFFFF, o[6] >> 16))
print("frames", len(cases), "encodes", ITER * len(cases), "encodes whose first try failed", total)
