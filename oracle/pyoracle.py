"""oracle/pyoracle.py -- TEST INFRASTRUCTURE.  ctypes binding of the CPU oracle (oracle/libtco.so =
squish_oracle.cpp + tile_oracle.cpp) and, when present, of the reference-glue build
(oracle/_ref/libtileconv_ref.so).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_TCO: Optional[ctypes.CDLL] = None
_REF: Optional[ctypes.CDLL] = None
_TUNED: Optional[ctypes.CDLL] = None


class Counters(ctypes.Structure):
    _fields_ = [(n, ctypes.c_ulonglong) for n in
                ("blocks", "single_blocks", "range_blocks", "cluster_blocks", "cand3", "cand4", "sweeps3", "sweeps4", "points")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(ref: bool = True) -> None:
    """Compiles the oracle (and, where /root/reference exists, the reference-glue build)."""
    subprocess.run(["make", "-C", _HERE, "libtco.so"], check=True, capture_output=True)
    if ref and os.path.isdir("/root/reference/tileconv"):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


def oracle() -> ctypes.CDLL:
    global _TCO
    if _TCO is None:
        path = os.path.join(_HERE, "libtco.so")
        if not os.path.exists(path):
            build(ref=False)
        _TCO = ctypes.CDLL(path)
        _TCO.tco_encode_tile.restype = ctypes.c_int
        _TCO.tco_required_space.restype = ctypes.c_int
        _TCO.tco_decode_tile.restype = ctypes.c_int
        _TCO.tco_pal_to_argb.restype = ctypes.c_int
        _TCO.tco_squish_flags.restype = ctypes.c_int
    return _TCO


def tuned() -> Optional[ctypes.CDLL]:
    """libtco_avx2.so: the restatement built with -O3 -mavx2 (oracle/Makefile), or None if it was not built or this
    CPU has no AVX2.  Same results bit for bit; only bench.py's cpu_baseline.tuned_port and its test use it."""
    global _TUNED
    if _TUNED is None:
        path = os.path.join(_HERE, "libtco_avx2.so")
        try:
            has_avx2 = " avx2 " in open("/proc/cpuinfo").read().replace("\n", " ")
        except OSError:
            has_avx2 = False
        if not has_avx2 or not os.path.exists(path):
            return None
        _TUNED = ctypes.CDLL(path)
    return _TUNED


def tuned_encode_tiles(tiles: np.ndarray, type_: int, quality: int, threads: int = 1) -> Optional[np.ndarray]:
    lib = tuned()
    if lib is None:
        return None
    tiles = np.ascontiguousarray(tiles, np.uint8).reshape(-1, 5120)
    out = np.zeros((tiles.shape[0], record_size(type_)), np.uint8)
    lib.tco_encode_tiles(_p(tiles), tiles.shape[0], type_, quality, _p(out), threads)
    return out


def reference() -> Optional[ctypes.CDLL]:
    """The reference's own glue linked with the restated squish, or None if it was never built."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libtileconv_ref.so")
        if not os.path.exists(path):
            return None
        _REF = ctypes.CDLL(path)
        _REF.tcref_encode_tile.restype = ctypes.c_int
        _REF.tcref_encode_tiles.restype = ctypes.c_int
        _REF.tcref_selftest.restype = ctypes.c_int
        if _REF.tcref_selftest() != 1:
            raise RuntimeError("oracle/_ref: Converter::ReorderColors mis-compiled (SURVEY.md F3, known answer KA1)")
    return _REF


def _p(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def record_size(type_: int, w: int = 64, h: int = 64) -> int:
    s = oracle().tco_required_space(type_, w, h)
    return s + 4 if s else 0


def encode_tile(palette: np.ndarray, indexed: np.ndarray, w: int, h: int, type_: int, quality: int) -> np.ndarray:
    palette = np.ascontiguousarray(palette, np.uint8); indexed = np.ascontiguousarray(indexed, np.uint8)
    out = np.zeros(4 + 4096, np.uint8)
    n = oracle().tco_encode_tile(_p(palette), _p(indexed), w, h, type_, quality, _p(out))
    return out[:n].copy()


def encode_tiles(tiles: np.ndarray, type_: int, quality: int, threads: int = 1) -> np.ndarray:
    tiles = np.ascontiguousarray(tiles, np.uint8).reshape(-1, 5120)
    n = tiles.shape[0]
    out = np.zeros((n, record_size(type_)), np.uint8)
    oracle().tco_encode_tiles(_p(tiles), n, type_, quality, _p(out), threads)
    return out


def encode_tiles_wh(tiles: np.ndarray, wh: np.ndarray, type_: int, quality: int) -> np.ndarray:
    """Per-tile sizes; output records padded to the 64x64 stride like libtcb200's."""
    tiles = np.ascontiguousarray(tiles, np.uint8).reshape(-1, 5120)
    n = tiles.shape[0]
    out = np.zeros((n, record_size(type_)), np.uint8)
    for i in range(n):
        w, h = int(wh[i][0]), int(wh[i][1])
        rec = encode_tile(tiles[i, :1024], tiles[i, 1024:1024 + w * h], w, h, type_, quality)
        out[i, :len(rec)] = rec
    return out


def compress_block(rgba: np.ndarray, flags: int) -> np.ndarray:
    rgba = np.ascontiguousarray(rgba, np.uint8).reshape(64)
    out = np.zeros(16, np.uint8)
    oracle().tco_compress_block(_p(rgba), flags, _p(out))
    return out[:8] if (flags & 1) else out


def squish_flags(type_: int, quality: int) -> int:
    return int(oracle().tco_squish_flags(type_, quality))


def gather_block(palette, indexed, w, h, bx, by) -> np.ndarray:
    palette = np.ascontiguousarray(palette, np.uint8); indexed = np.ascontiguousarray(indexed, np.uint8)
    out = np.zeros(64, np.uint8)
    oracle().tco_gather_block(_p(palette), _p(indexed), w, h, bx, by, _p(out))
    return out


def pal_to_argb(indexed: np.ndarray, palette: np.ndarray) -> np.ndarray:
    palette = np.ascontiguousarray(palette, np.uint8); indexed = np.ascontiguousarray(indexed, np.uint8)
    out = np.zeros((indexed.size, 4), np.uint8)
    oracle().tco_pal_to_argb(_p(indexed), _p(palette), _p(out), ctypes.c_uint32(indexed.size))
    return out


def decode_tile(record: np.ndarray, type_: int, reference_quirk: bool = False) -> np.ndarray:
    """(h, w, 4) pixels in memory order B,G,R,A.  reference_quirk mimics the reference's BC3 alpha decode."""
    record = np.ascontiguousarray(record, np.uint8)
    w = int(record[0]) | (int(record[1]) << 8); h = int(record[2]) | (int(record[3]) << 8)
    out = np.zeros((h, w, 4), np.uint8)
    fn = oracle().tco_decode_tile_ex
    fn.restype = ctypes.c_int
    n = fn(_p(record), type_, _p(out), 1 if reference_quirk else 0)
    assert n == w * h * 4
    return out


def ref_decode_tile(record: np.ndarray, type_: int) -> np.ndarray:
    """The reference's own block decoders (oracle/_ref): (h, w, 4) pixels in memory order R,G,B,A."""
    lib = reference()
    record = np.ascontiguousarray(record, np.uint8)
    w = int(record[0]) | (int(record[1]) << 8); h = int(record[2]) | (int(record[3]) << 8)
    out = np.zeros((h, w, 4), np.uint8)
    lib.tcref_decode_tile.restype = ctypes.c_int
    n = lib.tcref_decode_tile(_p(record), int(record.size), type_, _p(out), int(out.size))
    assert n == w * h * 4, n
    return out


def totals_reset() -> None:
    oracle().tco_totals_reset()


def totals() -> dict:
    c = Counters()
    oracle().tco_totals_get(ctypes.byref(c))
    return c.as_dict()


def ref_totals_reset() -> None:
    reference().tcref_totals_reset()


def ref_totals() -> dict:
    c = Counters()
    reference().tcref_totals_get(ctypes.byref(c))
    return c.as_dict()


def single_colour_table(bits: int, colours: int) -> np.ndarray:
    out = np.zeros(256 * 6, np.uint8)
    oracle().squish_oracle_single_colour_table(bits, colours, _p(out))
    return out.reshape(256, 2, 3)


def ref_encode_tile(palette, indexed, w, h, type_, quality) -> np.ndarray:
    lib = reference()
    palette = np.ascontiguousarray(palette, np.uint8); indexed = np.ascontiguousarray(indexed, np.uint8)
    out = np.zeros(8192, np.uint8)
    n = lib.tcref_encode_tile(_p(palette), _p(indexed), w, h, type_, quality, _p(out), 8192)
    return out[:n].copy()


def ref_encode_tiles(tiles: np.ndarray, type_: int, quality: int, threads: int) -> np.ndarray:
    lib = reference()
    tiles = np.ascontiguousarray(tiles, np.uint8).reshape(-1, 5120)
    n = tiles.shape[0]
    rec = record_size(type_)
    out = np.zeros((n, rec), np.uint8)
    failed = lib.tcref_encode_tiles(_p(tiles), n, type_, quality, _p(out), rec, threads)
    if failed:
        raise RuntimeError(f"reference failed on {failed} tiles")
    return out


PERCEPTUAL = (0.2126, 0.7152, 0.0722)   # libsquish <= 1.10's default metric (kColourMetricPerceptual)


def bound_audit(tiles: np.ndarray, type_: int, quality: int, scale: float = 0.0) -> dict:
    """Encodes `tiles` on the calling thread with the bound audit of squish_oracle.cpp switched on: every 4-colour
    candidate's error is compared with the lower bound the CUDA path prunes by (restated in binary32).  scale = 0
    uses the CUDA path's margin scale (2^-17)."""
    lib = oracle()
    lib.squish_oracle_bound_audit.argtypes = [ctypes.c_int, ctypes.c_float]
    lib.squish_oracle_bound_audit(1, scale)
    try:
        encode_tiles(tiles, type_, quality, threads=1)
        out = (ctypes.c_double * 5)()
        lib.squish_oracle_bound_audit_get(out)
    finally:
        lib.squish_oracle_bound_audit(0, 0.0)
    return {"tested": int(out[0]), "skipped": int(out[1]), "wrongly_skipped": int(out[2]), "below_raw_bound": int(out[3]),
            "worst_shortfall_over_margin": float(out[4])}


def set_simd(on: bool) -> None:
    """Cluster solve on 4-wide SSE2 vectors (identical results, ~2.3x faster): what a libsquish built with SQUISH_USE_SSE
    does with its Vec4.  Off by default (the scalar code is the definition); bench.py's CPU arms switch it on.  Applies
    to libtco.so, libtco_avx2.so and the copy linked into oracle/_ref."""
    for lib in (oracle(), tuned(), reference()):
        if lib is not None:
            lib.squish_oracle_set_simd(1 if on else 0)


def set_variant(metric=None, sse: bool = False) -> None:
    """libsquish-version knobs of the restated squish (squish_oracle.cpp header): the metric used when
    the caller passes none, and the SQUISH_USE_SSE arithmetic.  Applies to libtco.so and, when built,
    to the copy linked into oracle/_ref.  set_variant() restores the oracle's definition."""
    m = (ctypes.c_float * 3)(*metric) if metric is not None else None
    oracle().squish_oracle_set_variant(m, 1 if sse else 0)
    ref = reference()
    if ref is not None:
        ref.squish_oracle_set_variant(m, 1 if sse else 0)


class variant:
    """with pyoracle.variant(metric=pyoracle.PERCEPTUAL, sse=True): ..."""

    def __init__(self, metric=None, sse: bool = False):
        self.metric, self.sse = metric, sse

    def __enter__(self):
        set_variant(self.metric, self.sse)
        return self

    def __exit__(self, *exc):
        set_variant(None, False)
        return False


def rcp_estimate(values: np.ndarray) -> np.ndarray:
    """The host CPU's rcpps estimate of each float32 in `values`."""
    fn = oracle().squish_oracle_rcp_estimate
    fn.restype = ctypes.c_float
    fn.argtypes = [ctypes.c_float]
    return np.array([fn(float(v)) for v in np.asarray(values, np.float32).ravel()], np.float32)
