"""Alternates two images on one codec and reports every merge whose first try failed verification, with the failing seeds."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from limg_b200 import Codec, synth
names = sys.argv[1:] or ["c2_4k_photo", "c4_4k_flatui"]
c = Codec(0)
data = []
for n in names:
    img, alpha = synth.CONFIGS[n]()
    h, w = img.shape
    d = torch.from_numpy(img.view(np.int32)).cuda()
    codes = [torch.empty((h, w), dtype=torch.uint8, device="cuda") for _ in range(3)]
    data.append((n, d, w, h, alpha, {"codesA": codes[0].data_ptr(), "codesB": codes[1].data_ptr(), "codesC": codes[2].data_ptr()}))
nfail = 0
for it in range(40):
    for (n, d, w, h, alpha, stream) in data:
        for rep in range(1 + it % 3):
            c.blocked_encode3d_device(d.data_ptr(), w, h, alpha, 100, True, False, stream, None)
            c.sync()
            cnt = c.debug_counters()
            if cnt[24] > 0:
                nfail += 1
                dbg = c.debug_wave()
                print("FAIL", n, "iter", it, rep, "tries", cnt[24], "which", hex(cnt[31]), "failing seeds", dbg[100])
                for k in range(min(int(dbg[100]), 8)):
                    o = dbg[104 + k * 8: 112 + k * 8]
                    print("    stage %d try %d seed (%d,%d) recorded %d replayed %d first recorded rect ox %d oy %d rx %d ry %d" % (o[0] & 255, o[0] >> 8, o[1], o[2], o[3], o[4], o[5] & 0xFFFF, o[5] >> 16, o[6] & 0xFFFF, o[6] >> 16))
print("failures", nfail)
