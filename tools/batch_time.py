"""Frames of a batch encoded + decoded concurrently on ONE GPU: K independent limgcu contexts (own streams and scratch), one frame each per
step. The area scan is latency bound (IPC 0.14, one CTA per SM), so frames overlap almost for free. Usage: batch_time.py [workload] [lanes,...]"""
import sys
sys.path.insert(0, ".")
import statistics, numpy as np, torch
from limg_b200 import Codec, synth, AREA_DTYPE

name = sys.argv[1] if len(sys.argv) > 1 else "c5_1080p_frame0"
lanes_list = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "2", "4", "8"])]
img, alpha = synth.CONFIGS[name]()
h, w = img.shape
bx, by = (w + 7) // 8, (h + 7) // 8
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
main = torch.cuda.current_stream()


class Lane:
    def __init__(self, seed):
        self.codec = Codec(0)
        self.stream = torch.cuda.ExternalStream(self.codec.stream)
        frame = synth.photo_like(w, h, 100 + seed, 4 if alpha else 3) if "photo" in name or "frame" in name or "rgba" in name else img
        self.src = torch.from_numpy(frame.view(np.int32)).cuda()
        self.codes = [torch.empty((h, w), dtype=torch.uint8, device="cuda") for _ in range(3)]
        self.areas = torch.empty(bx * by * AREA_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        self.bmap = torch.empty(bx * by, dtype=torch.int32, device="cuda")
        self.cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.dec = torch.empty((h, w), dtype=torch.int32, device="cuda")
        self.st = {"areas": self.areas.data_ptr(), "area_count": self.cnt.data_ptr(), "block_to_area": self.bmap.data_ptr(),
                   "codesA": self.codes[0].data_ptr(), "codesB": self.codes[1].data_ptr(), "codesC": self.codes[2].data_ptr()}

    def step(self):
        self.codec.blocked_encode3d_device(self.src.data_ptr(), w, h, alpha, 100, True, False, self.st, None)
        self.codec.decode_device(self.areas.data_ptr(), self.bmap.data_ptr(), self.codes[0].data_ptr(), self.codes[1].data_ptr(), self.codes[2].data_ptr(), w, h, alpha, self.dec.data_ptr())


all_lanes = [Lane(i) for i in range(max(lanes_list))]
for lanes in lanes_list:
    ls = all_lanes[:lanes]
    times = []
    for it in range(8):
        flush.fill_(it)
        start, ends = torch.cuda.Event(enable_timing=True), []
        torch.cuda.synchronize()
        start.record(main)
        for l in ls:
            l.stream.wait_event(start)
        for l in ls:
            l.step()
            e = torch.cuda.Event(enable_timing=True)
            e.record(l.stream)
            ends.append(e)
        torch.cuda.synchronize()
        times.append(max(start.elapsed_time(e) for e in ends))
    ms = statistics.median(times[2:])
    fails = [int(l.codec.debug_counters()[24]) for l in ls]
    flags = [l.codec.debug_counters()[24:29].tolist() for l in ls]
    print("%s lanes %d: %.3f ms per step of %d frames -> %.0f Mpx/s (%.2f ms per frame); tries that failed in the last step, per lane: %s; step times %s" % (name, lanes, ms, lanes, lanes * w * h / ms / 1e3, ms / lanes, flags, [round(t, 2) for t in times]))
