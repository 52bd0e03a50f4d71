import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the shared libraries are build artefacts (git-ignored); build them if this checkout has none
    needed = [os.path.join(ROOT, "tileconv_b200", "libtcb200.so"), os.path.join(ROOT, "tileconv_b200", "libtcb200_host.so"),
              os.path.join(ROOT, "oracle", "libtco.so")]
    if not all(os.path.exists(p) for p in needed):
        import __graft_entry__
        __graft_entry__.build()


def _has_cuda() -> bool:
    try:
        from tileconv_b200 import capi
        return capi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.oracle()
    return pyoracle


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "encode_golden.npz")
    return dict(np.load(path))


GOLDEN_CONFIGS = [(t, q) for t in (1, 2, 3) for q in (0, 2, 3, 4, 5, 6, 7, 9)]
# one quality per equivalence class of ConverterDxt::getFlags (converter_dxt.cpp:191-208)
QUALITY_CLASSES = (2, 4, 6, 9)
