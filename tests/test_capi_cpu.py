"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/tcb200.h declares,
and fails loudly (never falls back to a CPU path) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

from tileconv_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tcb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(capi.library_path())
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(n for n, _, _ in capi.SYMBOLS) == names


def test_record_size_matches_reference_formula(oracle):
    for t in (1, 2, 3):
        for (w, h) in ((64, 64), (1, 1), (4, 4), (5, 3), (63, 64), (37, 22)):
            assert capi.record_size(t, w, h) == oracle.record_size(t, w, h)
    assert capi.record_size(0, 64, 64) == 0 and capi.record_size(1, 0, 1) == 0 and capi.record_size(1, 65, 1) == 0


def test_no_cpu_fallback_without_a_device():
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.TcbError) as e:
        capi.Encoder(1, 9)
    assert "no usable CUDA device" in str(e.value)


def test_create_rejects_bad_arguments():
    lib = capi.load_library()
    for args in ((0, 0, 9, 0), (0, 4, 9, 0), (0, 1, 10, 0), (0, 1, -1, 0), (0, 1, 9, -5)):
        assert not lib.tcb200_create(*args)
        assert lib.tcb200_last_error(None)


def test_product_sources_never_touch_the_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "tileconv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in text and "libtco" not in text and "squish_oracle" not in text, f


def test_candidate_tables_hold_libsquish_loop_order_and_the_layout_the_kernels_assume():
    """tcb200_selftest_tables (host only): list contents = the reference's (i,j[,k]) loops, field encoding, read-ahead room."""
    assert capi.load_library().tcb200_selftest_tables() == 0
