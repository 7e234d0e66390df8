"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/limgcu.h
declares, the C++ drop-in symbols of include/limg_dropin.h are present, and the product path fails loudly
without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "limg_b200", "liblimgcu.so")):
        g.build()
    from limg_b200 import _lib
    return _lib.load()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "limgcu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(limgcu_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(lib):
    from limg_b200 import _lib
    names = declared_symbols()
    assert set(names) == set(_lib.SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n


def test_cxx_dropin_symbols_exported():
    out = subprocess.run(["nm", "-DC", os.path.join(ROOT, "limg_b200", "liblimgcu.so")], capture_output=True, text=True, check=True).stdout
    for n in ("limg_blocked_encode3d_test(", "limg_encode3d_test(", "limg_encode3d_test_perf(", "limg_compare(", "limg_encode_test(",
              "limg_thread_pool_new(", "limg_thread_pool_destroy(", "limg_threading_max_threads(", "limg_b200_set_device(", "limg_b200_set_dither_mode("):
        assert n in out, n


def test_struct_layouts_match_header():
    from limg_b200 import _lib
    assert _lib.DECOMP_DTYPE.itemsize == 64 and _lib.AREA_DTYPE.itemsize == 120
    assert ctypes.sizeof(_lib.Planes) == 14 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_lib.Stream) == 6 * ctypes.sizeof(ctypes.c_void_p)
    assert _lib.AREA_DTYPE.fields["decomp"][1] == 56 and _lib.AREA_DTYPE.fields["ditherBefore"][1] == 40


def test_no_cpu_fallback(lib):
    """Without a CUDA device the context cannot be created and nothing is computed on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    assert lib.limgcu_create(0, ctypes.byref(h)) == 200  # LIMGCU_ERROR_NO_DEVICE
    from limg_b200 import Codec, LimgError
    with pytest.raises(LimgError):
        Codec(0)


def test_product_does_not_import_the_oracle():
    """limg_b200/ never references oracle/ (the oracle is the checker, not the product)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "limg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, f)
