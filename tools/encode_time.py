"""Encode alone: median CUDA-event time per config with the L2 flushed between iterations (environment knobs are read by limgcu_create)."""
import sys
sys.path.insert(0, ".")
import os, statistics, numpy as np, torch
from limg_b200 import Codec, synth, AREA_DTYPE
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2_4k_photo", "c4_4k_flatui", "c5_1080p_frame0", "c3_8k_rgba"]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
c = Codec(0)
if len(sys.argv) > 3 and sys.argv[3] == "aes":
    c.set_dither_mode(True)  # the reference's AES-round dither chain, walked on the host (dither_aes_host.cpp)
stream = torch.cuda.ExternalStream(c.stream)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
out = []
for name in names:
    img, alpha = synth.CONFIGS[name]()
    h, w = img.shape
    bx, by = (w + 7) // 8, (h + 7) // 8
    d = torch.from_numpy(img.view(np.int32)).cuda()
    codes = [torch.empty((h, w), dtype=torch.uint8, device="cuda") for _ in range(3)]
    areas = torch.empty(bx * by * AREA_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    bmap = torch.empty(bx * by, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = {"areas": areas.data_ptr(), "area_count": cnt.data_ptr(), "block_to_area": bmap.data_ptr(), "codesA": codes[0].data_ptr(), "codesB": codes[1].data_ptr(), "codesC": codes[2].data_ptr()}
    ev = []
    fails = 0
    with torch.cuda.stream(stream):
        for i in range(iters + 2):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            c.blocked_encode3d_device(d.data_ptr(), w, h, alpha, 100, True, False, st, None)
            e1.record(stream)
            ev.append((e0, e1))
    c.sync()
    fails = int(c.debug_counters()[24])
    ms = [a.elapsed_time(b) for a, b in ev[2:]]
    out.append("%s %.3f ms (min %.3f, failed first tries in the last run %d)" % (name, statistics.median(ms), min(ms), fails))
print(" | ".join(out))
