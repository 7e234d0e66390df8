"""Design tool: how often does a seed's masked growth equal its mask-free growth in the reference-order scan? (tools/scan_stats.c)"""
import ctypes as C, os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from limg_b200 import synth
from oracle import oracle as lo
HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_scan_stats.so")
subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", SO, os.path.join(HERE, "scan_stats.c"), os.path.join(HERE, "..", "oracle", "liblimg_oracle.so"), "-Wl,-rpath," + os.path.join(HERE, "..", "oracle")], check=True)
lo.lib()
m = C.CDLL(SO)
for name in sys.argv[1:] or ["c5_1080p_frame0"]:
    img, alpha = synth.CONFIGS[name]()
    h, w = img.shape
    bx, by = (w + 7) // 8, (h + 7) // 8
    table = lo.pass1(img, alpha)
    out = np.zeros(64)
    out[63] = int(os.environ.get("ROUNDS", "1"))  # extra rounds of the plan filter model
    m.scan_stats(table.ctypes.data_as(C.c_void_p), bx, by, 4 if alpha else 3, out.ctypes.data_as(C.c_void_p))
    print("%s plan filter (%d rounds): stage-0 candidates %d, declared alive %d, emitting seeds declared dead %d of %d" % (name, out[62], out[60], out[61], out[14], out[0] - out[1]))
    for s in range(2):
        o = out[s * 32:(s + 1) * 32]
        print("%s stage %d: expansions %d, emitting %d | R == mask-free R %d, mask-free R free %d (bad %d), R inside 8x8 %d, either %d | four-way %d: C == mask-free C %d, both predicted rects free %d (bad %d), C within +-8 of the centre %d, regrowth wins %d" % (
            name, s, o[0], o[0] - o[1], o[2], o[3], o[4], o[5], o[12], o[6], o[7], o[8], o[9], o[10], o[11]))
