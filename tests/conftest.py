import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref_lib():
    """The real reference (oracle/_ref/libref.so) -- only where it was built and the host's RSQRTSS matches the committed table."""
    from oracle import ref
    if not ref.available():
        ref.build()
    if not ref.available():
        pytest.skip("oracle/_ref/libref.so not built (reference sources absent)")
    import numpy as np
    from tools.dump_rsqrt_lut import committed_table, host_table
    if not np.array_equal(host_table(), committed_table()):
        pytest.skip("host RSQRTSS table differs from the committed (Intel) table: reference results are host dependent")
    ref.set_modes(True, False)
    return ref
