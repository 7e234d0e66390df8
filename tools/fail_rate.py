"""How often does the pipelined merge scan fail its verification (and which stage)? Usage: python tools/fail_rate.py config [runs]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from limg_b200 import Codec, synth
name = sys.argv[1] if len(sys.argv) > 1 else "c4_4k_flatui"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 20
c = Codec(0)
c.enable_phase_timing(True)
img, alpha = synth.CONFIGS[name]()
h, w = img.shape
d = torch.from_numpy(img.view(np.int32)).cuda()
codes = [torch.empty((h, w), dtype=torch.uint8, device="cuda") for _ in range(3)]
stream = {"codesA": codes[0].data_ptr(), "codesB": codes[1].data_ptr(), "codesC": codes[2].data_ptr()}
fails, which, ms = 0, [], []
for it in range(runs):
    c.blocked_encode3d_device(d.data_ptr(), w, h, alpha, 100, True, False, stream, None)
    c.sync()
    cnt = c.debug_counters()
    fails += int(cnt[24] > 0)
    which.append(int(cnt[31]))
    ms.append(c.phase_ms()["merge_scan"])
print(name, "runs", runs, "first-try failures", fails, "which (bit0 stage0, bit1 stage1, bit2 wave; <<4 per try):", [hex(x) for x in which], "merge ms median %.2f max %.2f" % (np.median(ms), np.max(ms)))
