import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from limg_b200 import Codec, synth, AREA_DTYPE
for name in ("c5_1080p_frame0", "c2_4k_photo"):
    img, alpha = synth.CONFIGS[name]()
    h, w = img.shape
    bx, by = w // 8, h // 8
    c = Codec(0)
    src = torch.from_numpy(img.view(np.int32)).cuda()
    codes = [torch.empty((h, w), dtype=torch.uint8, device="cuda") for _ in range(3)]
    areas = torch.empty(bx * by * AREA_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    bmap = torch.empty(bx * by, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = {"areas": areas.data_ptr(), "area_count": cnt.data_ptr(), "block_to_area": bmap.data_ptr(), "codesA": codes[0].data_ptr(), "codesB": codes[1].data_ptr(), "codesC": codes[2].data_ptr()}
    for _ in range(3):
        c.blocked_encode3d_device(src.data_ptr(), w, h, alpha, 100, True, False, st, None)
    c.sync()
    l0 = c.launch_count()
    t0 = time.perf_counter()
    c.blocked_encode3d_device(src.data_ptr(), w, h, alpha, 100, True, False, st, None)
    t1 = time.perf_counter()
    c.sync()
    t2 = time.perf_counter()
    print(name, "host enqueue %.3f ms, then sync %.3f ms, kernels per encode %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, c.launch_count() - l0))
    # back to back without sync: GPU-side time per encode when the host is ahead
    t0 = time.perf_counter()
    for _ in range(20):
        c.blocked_encode3d_device(src.data_ptr(), w, h, alpha, 100, True, False, st, None)
    t1 = time.perf_counter()
    c.sync()
    t2 = time.perf_counter()
    print("   20 back to back: host %.3f ms per call, total %.3f ms per encode" % ((t1 - t0) * 1e3 / 20, (t2 - t0) * 1e3 / 20))
