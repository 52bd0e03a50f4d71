"""ctypes binding of include/tcb200.h (libtcb200.so).  No CPU fallback: every compute call needs a
CUDA device and raises TcbError otherwise."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

TILE_RECORD = 5120
_LIB: Optional[ctypes.CDLL] = None

_u8p = ctypes.POINTER(ctypes.c_uint8)
_u16p = ctypes.POINTER(ctypes.c_uint16)



class Options(ctypes.Structure):
    """tcb200_options (include/tcb200.h): the libsquish-version knobs."""
    _fields_ = [("struct_size", ctypes.c_uint32), ("metric", ctypes.c_float * 3), ("reciprocal", ctypes.c_int32)]


RECIP_IEEE, RECIP_SSE = 0, 1
PERCEPTUAL = (0.2126, 0.7152, 0.0722)   # libsquish <= 1.10's default metric

# (name, restype, argtypes) for every symbol include/tcb200.h declares.
SYMBOLS = [
    ("tcb200_device_count", ctypes.c_int, []),
    ("tcb200_record_size", ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    ("tcb200_create", ctypes.c_void_p, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    ("tcb200_destroy", None, [ctypes.c_void_p]),
    ("tcb200_default_options", None, [ctypes.POINTER(Options)]),
    ("tcb200_create_ex", ctypes.c_void_p, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(Options)]),
    ("tcb200_get_options", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Options)]),
    ("tcb200_rejected_tiles", ctypes.c_longlong, [ctypes.c_void_p]),
    ("tcb200_selftest_rcpps", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    ("tcb200_last_error", ctypes.c_char_p, [ctypes.c_void_p]),
    ("tcb200_type", ctypes.c_int, [ctypes.c_void_p]),
    ("tcb200_quality", ctypes.c_int, [ctypes.c_void_p]),
    ("tcb200_device", ctypes.c_int, [ctypes.c_void_p]),
    ("tcb200_record_stride", ctypes.c_int, [ctypes.c_void_p]),
    ("tcb200_encode", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    ("tcb200_encode_multi", ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    ("tcb200_encode_device", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    ("tcb200_slab_count", ctypes.c_int, [ctypes.c_void_p]),
    ("tcb200_slab_capacity", ctypes.c_int, [ctypes.c_void_p]),
    ("tcb200_slab_input", ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int]),
    ("tcb200_slab_sizes", ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int]),
    ("tcb200_slab_output", ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int]),
    ("tcb200_slab_submit", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    ("tcb200_slab_wait", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    ("tcb200_slab_ready", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    ("tcb200_host_alloc", ctypes.c_void_p, [ctypes.c_size_t]),
    ("tcb200_host_free", None, [ctypes.c_void_p]),
    ("tcb200_pal_to_argb", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    ("tcb200_decode", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]),
    ("tcb200_time_decode_ms", ctypes.c_float, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    ("tcb200_time_kernel_ms", ctypes.c_float, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]),
    ("tcb200_fp32_peak", ctypes.c_double, [ctypes.c_void_p, ctypes.c_int]),
    ("tcb200_selftest_reciprocal", ctypes.c_longlong, [ctypes.c_void_p]),
    ("tcb200_launch_count", ctypes.c_longlong, [ctypes.c_void_p]),
    ("tcb200_selftest_tables", ctypes.c_int, []),
]


class TcbError(RuntimeError):
    pass


def library_path() -> str:
    # TCB200_LIB selects an alternative build of the same library (kernel tuning sweeps)
    return os.environ.get("TCB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtcb200.so")


def load_library() -> ctypes.CDLL:
    """Loads libtcb200.so (built in-tree by ``__graft_entry__.build()``) and types its symbols."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise TcbError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = ctypes.CDLL(path)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


def device_count() -> int:
    return int(load_library().tcb200_device_count())


def record_size(type_: int, width: int = 64, height: int = 64) -> int:
    return int(load_library().tcb200_record_size(type_, width, height))


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


class PinnedBuffer:
    """numpy view over cudaHostAlloc'ed memory (tcb200_host_alloc)."""

    def __init__(self, nbytes: int):
        self._lib = load_library()
        self.nbytes = int(nbytes)
        self.ptr = self._lib.tcb200_host_alloc(self.nbytes)
        if not self.ptr:
            raise TcbError(f"tcb200_host_alloc({nbytes}) failed")
        buf = (ctypes.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.uint8)

    def free(self) -> None:
        if self.ptr:
            self.array = None
            self._lib.tcb200_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def encode_multi(encoders, tiles_ptr: int, wh_ptr: Optional[int], n: int, out_ptr: int) -> None:
    """tcb200_encode_multi over the contexts of `encoders` (one per device): host pointers in, host pointers out."""
    lib = load_library()
    arr = (ctypes.c_void_p * len(encoders))(*[e.handle for e in encoders])
    rc = lib.tcb200_encode_multi(arr, len(encoders), tiles_ptr, wh_ptr, n, out_ptr)
    if rc != 0:
        msg = lib.tcb200_last_error(encoders[0].handle) if encoders and encoders[0].handle else b""
        raise TcbError(f"tcb200_encode_multi failed ({rc}): {(msg or b'').decode()}")


class Encoder:
    """One tcb200_ctx: device x pixel type (1,2,3 = BC1,BC2,BC3) x encode quality (0..9)."""

    def __init__(self, type_: int = 1, quality: int = 9, device: int = 0, max_batch_tiles: int = 0,
                 metric=None, sse: bool = False):
        """metric / sse: the libsquish-version knobs of tcb200_create_ex (None / False = tcb200_create)."""
        self._lib = load_library()
        if metric is None and not sse:
            self._ctx = self._lib.tcb200_create(device, type_, quality, max_batch_tiles)
        else:
            opt = Options()
            self._lib.tcb200_default_options(ctypes.byref(opt))
            if metric is not None:
                opt.metric[0], opt.metric[1], opt.metric[2] = (float(v) for v in metric)
            opt.reciprocal = RECIP_SSE if sse else RECIP_IEEE
            self._ctx = self._lib.tcb200_create_ex(device, type_, quality, max_batch_tiles, ctypes.byref(opt))
        if not self._ctx:
            msg = self._lib.tcb200_last_error(None)
            raise TcbError((msg or b"tcb200_create failed").decode())
        self.type = type_
        self.quality = quality
        self.device = device
        self.stride = int(self._lib.tcb200_record_stride(self._ctx))

    # -- lifetime -------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.tcb200_destroy(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._ctx

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self._lib.tcb200_last_error(self._ctx)
            raise TcbError(f"{what} failed ({rc}): {(msg or b'').decode()}")

    # -- host-to-host ---------------------------------------------------------------------------
    def encode(self, tiles: np.ndarray, wh: Optional[np.ndarray] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        """tiles: (n, 5120) uint8; wh: optional (n, 2) uint16; returns (n, stride) uint8."""
        tiles = np.ascontiguousarray(tiles, dtype=np.uint8).reshape(-1, TILE_RECORD)
        n = tiles.shape[0]
        if out is None:
            out = np.zeros((n, self.stride), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.size >= n * self.stride and out.flags.c_contiguous
        whp = None
        if wh is not None:
            wh = np.ascontiguousarray(wh, dtype=np.uint16).reshape(n, 2)
            whp = _ptr(wh)
        self._check(self._lib.tcb200_encode(self._ctx, _ptr(tiles) if n else None, whp, n, _ptr(out) if n else None), "tcb200_encode")
        return out

    def encode_ptr(self, tiles_ptr: int, wh_ptr: Optional[int], n: int, out_ptr: int) -> None:
        self._check(self._lib.tcb200_encode(self._ctx, tiles_ptr, wh_ptr, n, out_ptr), "tcb200_encode")

    # -- device-resident ------------------------------------------------------------------------
    def encode_device(self, d_tiles: int, d_wh: Optional[int], n: int, d_out: int, stream: Optional[int] = None) -> None:
        self._check(self._lib.tcb200_encode_device(self._ctx, d_tiles, d_wh, n, d_out, stream), "tcb200_encode_device")

    def time_kernel_ms(self, d_tiles: int, d_wh: Optional[int], n: int, d_out: int, reps: int) -> float:
        ms = float(self._lib.tcb200_time_kernel_ms(self._ctx, d_tiles, d_wh, n, d_out, reps))
        if ms < 0:
            self._check(int(ms), "tcb200_time_kernel_ms")
        return ms

    def time_decode_ms(self, d_records: int, n: int, d_bgra: int, reps: int, reference_bc3_alpha: bool = False) -> float:
        ms = float(self._lib.tcb200_time_decode_ms(self._ctx, d_records, n, d_bgra, 1 if reference_bc3_alpha else 0, reps))
        if ms < 0:
            self._check(int(ms), "tcb200_time_decode_ms")
        return ms

    def fp32_peak(self, fused: bool) -> float:
        v = float(self._lib.tcb200_fp32_peak(self._ctx, 1 if fused else 0))
        if v < 0:
            self._check(int(v), "tcb200_fp32_peak")
        return v

    def selftest_reciprocal(self) -> int:
        v = int(self._lib.tcb200_selftest_reciprocal(self._ctx))
        if v < 0:
            self._check(v, "tcb200_selftest_reciprocal")
        return v

    def launch_count(self) -> int:
        return int(self._lib.tcb200_launch_count(self._ctx))

    def options(self) -> Options:
        opt = Options()
        self._check(self._lib.tcb200_get_options(self._ctx, ctypes.byref(opt)), "tcb200_get_options")
        return opt

    def rejected_tiles(self) -> int:
        v = int(self._lib.tcb200_rejected_tiles(self._ctx))
        if v < 0:
            self._check(v, "tcb200_rejected_tiles")
        return v

    def selftest_rcpps(self, values: np.ndarray) -> np.ndarray:
        values = np.ascontiguousarray(values, dtype=np.float32).ravel()
        out = np.zeros_like(values)
        self._check(self._lib.tcb200_selftest_rcpps(self._ctx, _ptr(values), values.size, _ptr(out)), "tcb200_selftest_rcpps")
        return out

    # -- expand stage ---------------------------------------------------------------------------
    def pal_to_argb(self, tiles: np.ndarray) -> np.ndarray:
        tiles = np.ascontiguousarray(tiles, dtype=np.uint8).reshape(-1, TILE_RECORD)
        n = tiles.shape[0]
        out = np.zeros((n, 4096, 4), dtype=np.uint8)
        self._check(self._lib.tcb200_pal_to_argb(self._ctx, _ptr(tiles) if n else None, n, _ptr(out) if n else None), "tcb200_pal_to_argb")
        return out

    def decode(self, records: np.ndarray, reference_bc3_alpha: bool = False) -> np.ndarray:
        """(n, stride) encoded records -> (n, 4096, 4) B,G,R,A pixels at row pitch w of each tile."""
        records = np.ascontiguousarray(records, dtype=np.uint8).reshape(-1, self.stride)
        n = records.shape[0]
        out = np.zeros((n, 4096, 4), dtype=np.uint8)
        self._check(self._lib.tcb200_decode(self._ctx, _ptr(records) if n else None, n, _ptr(out) if n else None,
                                            1 if reference_bc3_alpha else 0), "tcb200_decode")
        return out

    # -- slabs ----------------------------------------------------------------------------------
    def slab_count(self) -> int:
        return int(self._lib.tcb200_slab_count(self._ctx))

    def slab_capacity(self) -> int:
        return int(self._lib.tcb200_slab_capacity(self._ctx))

    def _slab_view(self, ptr: int, nbytes: int, dtype) -> np.ndarray:
        if not ptr:
            self._check(-3, "tcb200_slab buffer")
        buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype)

    def slab_input(self, slab: int) -> np.ndarray:
        cap = self.slab_capacity()
        return self._slab_view(self._lib.tcb200_slab_input(self._ctx, slab), cap * TILE_RECORD, np.uint8).reshape(cap, TILE_RECORD)

    def slab_sizes(self, slab: int) -> np.ndarray:
        cap = self.slab_capacity()
        return self._slab_view(self._lib.tcb200_slab_sizes(self._ctx, slab), cap * 4, np.uint16).reshape(cap, 2)

    def slab_output(self, slab: int) -> np.ndarray:
        cap = self.slab_capacity()
        return self._slab_view(self._lib.tcb200_slab_output(self._ctx, slab), cap * self.stride, np.uint8).reshape(cap, self.stride)

    def slab_submit(self, slab: int, n: int) -> None:
        self._check(self._lib.tcb200_slab_submit(self._ctx, slab, n), "tcb200_slab_submit")

    def slab_wait(self, slab: int) -> None:
        self._check(self._lib.tcb200_slab_wait(self._ctx, slab), "tcb200_slab_wait")

    def slab_ready(self, slab: int) -> bool:
        rc = self._lib.tcb200_slab_ready(self._ctx, slab)
        if rc < 0:
            self._check(rc, "tcb200_slab_ready")
        return rc == 1
