"""Encodes every row band of a workload (as bench.py --mode rowband cuts it) on one GPU and prints the scan's tries per band."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch, statistics
from limg_b200 import Codec, synth, shard
name = sys.argv[1] if len(sys.argv) > 1 else "c3_8k_rgba"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
img, alpha = synth.CONFIGS[name]()
h, w = img.shape
c = Codec(0)
stream = torch.cuda.ExternalStream(c.stream)
for r, (y0, y1) in enumerate(shard.row_bands(h, world)):
    band = np.ascontiguousarray(img[y0:y1])
    d = torch.from_numpy(band.view(np.int32)).cuda()
    bh = y1 - y0
    codes = [torch.empty((bh, w), dtype=torch.uint8, device="cuda") for _ in range(3)]
    st = {"codesA": codes[0].data_ptr(), "codesB": codes[1].data_ptr(), "codesC": codes[2].data_ptr()}
    ms, tries = [], []
    for it in range(6):
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            c.blocked_encode3d_device(d.data_ptr(), w, bh, alpha, 100, True, False, st, None)
            e1.record(stream)
        c.sync()
        ms.append(e0.elapsed_time(e1))
        tries.append(int(c.debug_counters()[24]))
    print("band %d rows %d..%d: encode ms %s failed tries %s which 0x%x" % (r, y0, y1, [round(m, 2) for m in ms], tries, int(c.debug_counters()[31])))
    if tries[-1]:
        dbg = c.debug_wave()
        n = int(dbg[100])
        print("   verification failures recorded: %d; first ones (stage | attempt << 8, x, y, recorded count, replayed so far, first recorded rect):" % n)
        for s in range(min(n, 8)):
            o = dbg[104 + s * 8: 112 + s * 8]
            print("     stage %d attempt %d seed (%d, %d) recorded %d replayed %d rect0 ox %d oy %d rx %d ry %d" % (o[0] & 0xFF, o[0] >> 8, o[1], o[2], o[3], o[4], o[5] & 0xFFFF, o[5] >> 16, o[6] & 0xFFFF, o[6] >> 16))
