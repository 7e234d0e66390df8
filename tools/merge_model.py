"""Design tool: runs tools/merge_model.c (CPU model of the row-pipelined merge scan) on the synthetic configs.
Usage: python tools/merge_model.py [config ...] [--workers N] [--margin M]"""
import argparse, ctypes as C, os, subprocess, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from limg_b200 import synth
from oracle import oracle as lo

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_merge_model.so")


def build():
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", SO, os.path.join(HERE, "merge_model.c"),
                    os.path.join(HERE, "..", "oracle", "liblimg_oracle.so"), "-Wl,-rpath," + os.path.join(HERE, "..", "oracle")], check=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c1_512_gradient", "c5_1080p_frame0"])
    ap.add_argument("--workers", type=int, default=128)
    ap.add_argument("--margin", type=int, nargs="*", default=[4, 8])
    ap.add_argument("--sequential", type=int, default=0)
    ap.add_argument("--predict", type=float, default=0.0)
    ap.add_argument("--wide", type=float, default=8.0)
    ap.add_argument("--depth", type=float, default=0.0)
    ap.add_argument("--jitter", type=float, default=0.0)
    ap.add_argument("--runrule", type=float, default=0.0)
    ap.add_argument("--poll", type=float, default=0.4)
    ap.add_argument("--exp", type=float, default=1.0)
    a = ap.parse_args()
    build()
    lo.lib()
    m = C.CDLL(SO)
    m.model_run.restype = C.c_size_t
    costs = (C.c_double * 12)(1.0, a.poll, 0.1, a.exp, 0.8, 0.3, 0.1, a.predict, a.wide, a.depth, a.jitter, a.runrule)
    for name in a.configs:
        img, alpha = synth.CONFIGS[name]()
        h, w = img.shape
        bx, by = (w + 7) // 8, (h + 7) // 8
        table = lo.pass1(img, alpha)
        ref, _ = lo.merge(table, bx, by, alpha)
        ref = ref[ref["stage"] < 2]
        for margin in a.margin:
            areas = np.zeros(bx * by, dtype=lo.AREA_DTYPE)
            stats = np.zeros(64)
            t = time.time()
            n = m.model_run(table.ctypes.data_as(C.c_void_p), bx, by, 4 if alpha else 3, a.workers, margin, a.sequential, costs,
                            areas.ctypes.data_as(C.c_void_p), stats.ctypes.data_as(C.c_void_p))
            got = areas[:n]
            same = n == len(ref) and all(np.array_equal(got[k], ref[k]) for k in ("ox", "oy", "rx", "ry", "stage"))
            print("%s margin %d workers %d: %d areas (ref %d) identical=%s  [model took %.1fs]" % (name, margin, a.workers, n, len(ref), same, time.time() - t))
            for s in range(2):
                o = stats[s * 32:(s + 1) * 32]
                print("  stage %d: expansions %d re-expansions %d polls %d conflicts %d emitted %d VERIFY-FAIL %d makespan %.0f us (preds %d)" % (s, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]))
                print("     left-ext hist", o[8:20].astype(int).tolist(), " box-width hist", o[20:32].astype(int).tolist())


if __name__ == "__main__":
    main()
