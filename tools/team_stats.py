"""Counters of the merge scan with teams of warps per row (LIMGCU_SCAN_TEAM): expansions, re-expansions, expansions on the turn, wasted ones."""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from limg_b200 import Codec, synth
c = Codec(0)
name = sys.argv[1] if len(sys.argv) > 1 else "c2_4k_photo"
img, alpha = synth.CONFIGS[name]()
h, w = img.shape
d = torch.from_numpy(img.view(np.int32)).cuda()
codes = [torch.empty((h, w), dtype=torch.uint8, device="cuda") for _ in range(3)]
stream = {"codesA": codes[0].data_ptr(), "codesB": codes[1].data_ptr(), "codesC": codes[2].data_ptr()}
for it in range(3):
    c.blocked_encode3d_device(d.data_ptr(), w, h, alpha, 100, True, False, stream, None)
    c.sync()
cnt = c.debug_counters()
dbg = c.debug_wave()
s = cnt[8:16]
print("team %s cluster %s %s: stage 0 expansions %d re-expansions %d polls %d | stage 1 expansions %d re-expansions %d polls %d | expanded on the turn %d, covered before the turn %d | failed first tries %d" % (
    os.environ.get("LIMGCU_SCAN_TEAM", "1"), os.environ.get("LIMGCU_SCAN_CLUSTER", "-"), name, s[0], s[1], s[2], s[4], s[5], s[6], dbg[241], dbg[240], cnt[24]))
