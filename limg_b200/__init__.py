"""limg_b200 -- B200-native (sm_100a) implementation of limg's encode/decode hot path.

The product is limg_b200/liblimgcu.so (C ABI: include/limgcu.h, C++ drop-in: include/limg_dropin.h); this package is
the thin Python mirror of the reference's operator interface used by the tests and the benchmark.
"""
from . import synth  # noqa: F401
from ._lib import AREA_DTYPE, DECOMP_DTYPE, LimgError  # noqa: F401
from .api import Codec  # noqa: F401
from .batch import BatchCodec  # noqa: F401
