"""Decode kernel alone: CUDA-event time with the L2 flushed between iterations, as GB/s of the 7 algorithmic bytes per pixel."""
import sys
sys.path.insert(0, ".")
import json, numpy as np, torch
from limg_b200 import Codec, synth, AREA_DTYPE
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6545.0
c = Codec(0)
stream = torch.cuda.ExternalStream(c.stream)
for name, (w, h, alpha) in {"4K RGB": (3840, 2160, False), "8K RGB": (7680, 4320, False), "8K RGBA": (7680, 4320, True)}.items():
    if len(sys.argv) > 2 and name not in sys.argv[2].split(","):
        continue
    img = synth.photo_like(w, h, 1, 4 if alpha else 3)
    d = torch.from_numpy(img.view(np.int32)).cuda()
    bx, by = w // 8, h // 8
    codes = [torch.empty((h, w), dtype=torch.uint8, device="cuda") for _ in range(3)]
    areas = torch.empty(bx * by * AREA_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    bmap = torch.empty(bx * by, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    dec = torch.empty((h, w), dtype=torch.int32, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    st = {"areas": areas.data_ptr(), "area_count": cnt.data_ptr(), "block_to_area": bmap.data_ptr(), "codesA": codes[0].data_ptr(), "codesB": codes[1].data_ptr(), "codesC": codes[2].data_ptr()}
    c.blocked_encode3d_device(d.data_ptr(), w, h, alpha, 100, True, False, st, {"pDecoded": dec.data_ptr()})
    c.sync()
    ref = dec.clone()
    for variant in [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["42"])]:
        c.set_decode_variant(variant)
        dec.zero_()
        times = []
        with torch.cuda.stream(stream):
            for i in range(12):
                flush.fill_(i)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                c.decode_device(areas.data_ptr(), bmap.data_ptr(), codes[0].data_ptr(), codes[1].data_ptr(), codes[2].data_ptr(), w, h, alpha, dec.data_ptr())
                e1.record(stream)
                times.append((e0, e1))
        c.sync()
        ms = sorted(a.elapsed_time(b) for a, b in times[2:])
        med = ms[len(ms) // 2]
        gbs = 7.0 * w * h / (med * 1e-3) / 1e9
        print("variant %d %s: decode %.1f us median (min %.1f), %.0f GB/s of 7 B/px = %.1f %% of %.0f GB/s; identical to the in-encoder reconstruction: %s" % (variant, name, med * 1e3, ms[0] * 1e3, gbs, 100 * gbs / peak, peak, bool(torch.equal(ref, dec))))
