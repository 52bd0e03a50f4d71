"""tileconv_b200 -- B200 (sm_100a) implementation of tileconv's TIS/MOS -> TBC/MBC encode hot path.

The product is the C-ABI shared library ``libtcb200.so`` (``include/tcb200.h``; sources in
``tileconv_b200/csrc``) plus the C++ host layer in ``tileconv_b200/host`` that mirrors the
reference's ``Converter`` / ``TileThreadPool`` seams.  This Python package is only a thin ctypes
binding used by ``tests/``, ``bench.py`` and ``__graft_entry__.py``; it never computes anything on
the CPU and raises if the CUDA library is missing.
"""
from .capi import (  # noqa: F401
    Encoder,
    TcbError,
    device_count,
    encode_multi,
    library_path,
    load_library,
    record_size,
)

__all__ = ["Encoder", "TcbError", "device_count", "encode_multi", "library_path", "load_library", "record_size"]
