"""The restated tile path (oracle/tile_oracle.cpp) against the reference's own glue code compiled from
/root/reference (oracle/_ref), byte for byte.  Skipped where the reference build is absent; the
committed golden fixture (made by the same reference build) is checked in test_golden.py."""
import numpy as np
import pytest

from tileconv_b200 import synth
from conftest import QUALITY_CLASSES


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.reference() is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle


@pytest.mark.parametrize("type_", (1, 2, 3))
@pytest.mark.parametrize("quality", QUALITY_CLASSES)
def test_full_tiles(ref, type_, quality):
    tiles = np.concatenate([synth.uniform_tiles(1, 21), synth.smooth_tiles(2, 22), synth.keyed_tiles(2, 23)])
    assert np.array_equal(ref.encode_tiles(tiles, type_, quality, threads=4),
                          ref.ref_encode_tiles(tiles, type_, quality, threads=4))


@pytest.mark.parametrize("type_", (1, 2, 3))
def test_ragged_tiles(ref, type_):
    tiles, wh = synth.ragged_tiles(8, 24)
    for i in range(len(tiles)):
        w, h = int(wh[i, 0]), int(wh[i, 1])
        for q in (2, 9):
            a = ref.encode_tile(tiles[i, :1024], tiles[i, 1024:1024 + w * h], w, h, type_, q)
            b = ref.ref_encode_tile(tiles[i, :1024], tiles[i, 1024:1024 + w * h], w, h, type_, q)
            assert np.array_equal(a, b), (i, w, h, q)


def test_nan_axis_blocks(ref):
    """Degenerate covariance (NaN principal axis): the restated path must still follow the reference glue."""
    from test_cuda_parity import nan_axis_tile
    t = nan_axis_tile()
    for type_ in (1, 2, 3):
        for q in (4, 9):
            assert np.array_equal(ref.encode_tiles(t, type_, q), ref.ref_encode_tiles(t, type_, q, threads=1))


@pytest.mark.parametrize("type_", (1, 2, 3))
def test_restated_decoder_matches_reference_decoders(ref, type_):
    """The reference's decodeBlockDxt1/3/5 (reached through ConverterDxt::convert in the decode direction;
    the stubbed quantiser hands the decoded pixels back) against oracle/tile_oracle.cpp's decoder.  BC3
    needs reference_quirk=True: the reference reads the alpha half two bytes late (converter_dxt.cpp:351)."""
    tiles, wh = synth.ragged_tiles(8, 25)
    for i in range(len(tiles)):
        w, h = int(wh[i, 0]), int(wh[i, 1])
        rec = ref.encode_tile(tiles[i, :1024], tiles[i, 1024:1024 + w * h], w, h, type_, 4)
        mine = ref.decode_tile(rec, type_, reference_quirk=True)[:, :, [2, 1, 0, 3]]     # B,G,R,A -> R,G,B,A
        assert np.array_equal(mine, ref.ref_decode_tile(rec, type_)), (i, w, h)
        if type_ != 3:
            assert np.array_equal(ref.decode_tile(rec, type_), ref.decode_tile(rec, type_, reference_quirk=True))
