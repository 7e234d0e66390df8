"""Dump the host's RSQRTSS table (through oracle/_ref/libref.so) into oracle/rsqrt_lut.h and
limg_b200/csrc/rsqrt_lut.h, or just check the host against the committed table (--check).

The x86 reciprocal-square-root approximation is a pure table function of (exponent parity, top 10
mantissa bits); Intel and AMD parts use different tables, so the table the GPU kernels and the C
oracle use is data, not code. The committed table was measured on an Intel Xeon (this container).
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402


def host_table() -> np.ndarray:
    L = ref.lib()
    m = np.arange(1 << 23, dtype=np.uint32)
    tab = np.zeros(2048, np.uint32)
    for par, e in ((0, 127), (1, 128)):
        x = ((np.uint32(e) << 23) | m).view(np.float32)
        out = np.empty_like(x)
        L.ref_rsqrtss_many(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(x.size))
        r = out.view(np.uint32).reshape(1024, 1 << 13)
        assert (r == r[:, :1]).all(), "host RSQRTSS is not a function of the top 10 mantissa bits"
        tab[par * 1024:(par + 1) * 1024] = r[:, 0]
    assert ((tab >> 23) == 126).all() and ((tab & 0x7FF) == 0).all()
    return ((tab >> 11) & 0xFFF).astype(np.uint16)


def committed_table() -> np.ndarray:
    txt = open(os.path.join(ROOT, "oracle", "rsqrt_lut.h")).read()
    body = txt[txt.index("{") + 1: txt.rindex("}")]
    return np.array([int(v, 16) for v in body.replace("\n", " ").split(",") if v.strip()], dtype=np.uint16)


if __name__ == "__main__":
    host = host_table()
    same = np.array_equal(host, committed_table())
    print("host RSQRTSS table matches committed table:", same)
    sys.exit(0 if same or "--check" not in sys.argv else 1)
