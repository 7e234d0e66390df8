"""Host-buffer entry points separately (pinned memory, wall clock): limgcu_host_encode_stream and limgcu_host_decode of one 4K frame."""
import sys, time
sys.path.insert(0, ".")
import ctypes as C, numpy as np, torch
from limg_b200 import Codec, synth, AREA_DTYPE
img, alpha = synth.CONFIGS["c2_4k_photo"]()
h, w = img.shape
c = Codec(0)
h_src = torch.from_numpy(img.view(np.int32)).pin_memory()
h_codes = [torch.empty((h, w), dtype=torch.uint8).pin_memory() for _ in range(3)]
h_areas = torch.empty((w // 8) * (h // 8) * AREA_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
h_dec = torch.empty((h, w), dtype=torch.int32).pin_memory()
n = C.c_uint32(0)
def enc():
    assert c.lib.limgcu_host_encode_stream(c.h, h_src.data_ptr(), w, h, 0, 100, 1, h_areas.data_ptr(), C.byref(n), h_codes[0].data_ptr(), h_codes[1].data_ptr(), h_codes[2].data_ptr(), None) == 0
def dec():
    assert c.lib.limgcu_host_decode(c.h, h_areas.data_ptr(), n.value, h_codes[0].data_ptr(), h_codes[1].data_ptr(), h_codes[2].data_ptr(), w, h, 0, h_dec.data_ptr()) == 0
for f, name in ((enc, "limgcu_host_encode_stream"), (dec, "limgcu_host_decode")):
    for _ in range(3): f()
    t = []
    for _ in range(10):
        t0 = time.perf_counter(); f(); t.append(time.perf_counter() - t0)
    print("%s: median %.3f ms (min %.3f)" % (name, sorted(t)[5] * 1e3, min(t) * 1e3))
# raw copies for comparison
d = torch.empty((h, w), dtype=torch.int32, device="cuda")
for name, fn in (("H2D 33 MB", lambda: d.copy_(h_src, non_blocking=True)), ("D2H 33 MB", lambda: h_dec.copy_(d, non_blocking=True))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); print("%s: %.3f ms" % (name, (time.perf_counter() - t0) * 100))
