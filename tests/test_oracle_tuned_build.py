"""oracle/libtco_avx2.so (the restatement built with -O3 -mavx2, bench.py's cpu_baseline.tuned_port) must be the same
function as the reference-flags build: no contraction, no reassociation, identical records."""
import numpy as np
import pytest

from oracle import pyoracle
from tileconv_b200 import synth


def test_tuned_build_is_bit_identical():
    if pyoracle.tuned() is None:
        pytest.skip("no AVX2 on this CPU or libtco_avx2.so not built")
    tiles = np.concatenate([synth.uniform_tiles(4, 21), synth.smooth_tiles(4, 22), synth.keyed_tiles(4, 23), synth.corner_case_tiles(4, 24)])
    for type_, quality in ((1, 9), (1, 2), (2, 6), (3, 4), (3, 9)):
        want = pyoracle.encode_tiles(tiles, type_, quality, threads=4)
        got = pyoracle.tuned_encode_tiles(tiles, type_, quality, threads=4)
        assert np.array_equal(got, want), (type_, quality)


def test_simd_cluster_solve_is_bit_identical():
    """squish_oracle_set_simd(1): the cluster solve on SSE2 vectors (bench.py's CPU arms) gives the scalar code's records."""
    tiles = np.concatenate([synth.uniform_tiles(6, 31), synth.smooth_tiles(6, 32), synth.keyed_tiles(6, 33), synth.corner_case_tiles(6, 34)])
    for type_, quality in ((1, 9), (1, 4), (2, 6), (3, 4), (3, 9)):
        want = pyoracle.encode_tiles(tiles, type_, quality, threads=4)
        pyoracle.set_simd(True)
        try:
            got = pyoracle.encode_tiles(tiles, type_, quality, threads=4)
            ref = pyoracle.ref_encode_tiles(tiles, type_, quality, 4) if pyoracle.reference() is not None else got
        finally:
            pyoracle.set_simd(False)
        assert np.array_equal(got, want) and np.array_equal(ref, want), (type_, quality)
    # the perceptual metric goes through the same vector code (the SSE-arithmetic knob keeps its scalar emulation)
    with pyoracle.variant(metric=pyoracle.PERCEPTUAL):
        want = pyoracle.encode_tiles(tiles, 3, 9, threads=4)
        pyoracle.set_simd(True)
        try:
            got = pyoracle.encode_tiles(tiles, 3, 9, threads=4)
        finally:
            pyoracle.set_simd(False)
        assert np.array_equal(got, want)
