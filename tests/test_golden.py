"""Golden fixture tests/golden/encode_golden.npz (made by tests/golden/make_golden.py from the
reference's own glue + restated squish): the oracle must reproduce it on the CPU, the CUDA path on
the GPU, bit for bit, every pixel type and every quality class, full and ragged tiles."""
import numpy as np
import pytest

from conftest import GOLDEN_CONFIGS


@pytest.mark.parametrize("type_,quality", GOLDEN_CONFIGS)
def test_oracle_reproduces_golden(oracle, golden, type_, quality):
    got = oracle.encode_tiles(golden["tiles"], type_, quality, threads=2)
    assert np.array_equal(got, golden[f"full_t{type_}_q{quality}"])
    got = oracle.encode_tiles_wh(golden["ragged_tiles"], golden["ragged_wh"], type_, quality)
    assert np.array_equal(got, golden[f"ragged_t{type_}_q{quality}"])


@pytest.mark.gpu
@pytest.mark.parametrize("type_,quality", GOLDEN_CONFIGS)
def test_cuda_reproduces_golden(oracle, golden, type_, quality):
    from tileconv_b200 import Encoder
    with Encoder(type_, quality) as enc:
        got = enc.encode(golden["tiles"])
        assert np.array_equal(got, golden[f"full_t{type_}_q{quality}"])
        wh = golden["ragged_wh"]
        got = enc.encode(golden["ragged_tiles"], wh)
        want = golden[f"ragged_t{type_}_q{quality}"]
        for i in range(len(wh)):
            size = oracle.record_size(type_, int(wh[i, 0]), int(wh[i, 1]))
            assert np.array_equal(got[i, :size], want[i, :size]), (i, wh[i])


PERCEPTUAL_CONFIGS = [(1, 2), (1, 9), (2, 4), (3, 7)]


@pytest.mark.parametrize("type_,quality", PERCEPTUAL_CONFIGS)
def test_oracle_reproduces_perceptual_metric_golden(oracle, golden, type_, quality):
    """Records the reference glue produced with the restated squish switched to libsquish <= 1.10's default metric."""
    with oracle.variant(metric=oracle.PERCEPTUAL):
        got = oracle.encode_tiles(golden["tiles"], type_, quality, threads=2)
    assert np.array_equal(got, golden[f"perceptual_t{type_}_q{quality}"])
    assert not np.array_equal(got, golden[f"full_t{type_}_q{quality}"])


@pytest.mark.gpu
@pytest.mark.parametrize("type_,quality", PERCEPTUAL_CONFIGS)
def test_cuda_reproduces_perceptual_metric_golden(golden, type_, quality):
    from tileconv_b200 import Encoder, capi
    with Encoder(type_, quality, metric=capi.PERCEPTUAL) as enc:
        assert np.array_equal(enc.encode(golden["tiles"]), golden[f"perceptual_t{type_}_q{quality}"])
