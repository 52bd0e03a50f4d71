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
