"""Independent format check: Pillow's DDS reader (its own BC1/BC2/BC3 decoder, unrelated to the reference)
must read the oracle's blocks as the picture the oracle's decoder reads.  Catches endianness / index-order /
block-order mistakes that a self-consistent encoder+decoder pair would hide."""
import io
import struct

import numpy as np
import pytest

from tileconv_b200 import synth

PIL = pytest.importorskip("PIL.Image")

FOURCC = {1: b"DXT1", 2: b"DXT3", 3: b"DXT5"}


def dds_bytes(blocks: bytes, type_: int, w: int = 64, h: int = 64) -> bytes:
    # DDS_HEADER: size 124, flags CAPS|HEIGHT|WIDTH|PIXELFORMAT|LINEARSIZE, pixel format FOURCC
    flags = 0x1 | 0x2 | 0x4 | 0x1000 | 0x80000
    header = struct.pack("<7I", 124, flags, h, w, len(blocks), 0, 0) + b"\0" * 44
    header += struct.pack("<2I4s5I", 32, 0x4, FOURCC[type_], 0, 0, 0, 0, 0)
    header += struct.pack("<5I", 0x1000, 0, 0, 0, 0)
    return b"DDS " + header + blocks


@pytest.mark.parametrize("type_", (1, 2, 3))
def test_pillow_reads_the_same_picture(oracle, type_):
    tiles = np.concatenate([synth.smooth_tiles(2, 121), synth.keyed_tiles(2, 122), synth.uniform_tiles(1, 123)])
    records = oracle.encode_tiles(tiles, type_, 4, threads=4)
    for rec in records:
        img = PIL.open(io.BytesIO(dds_bytes(rec[4:].tobytes(), type_)))
        img.load()
        got = np.asarray(img.convert("RGBA")).astype(int)                     # R,G,B,A
        want = oracle.decode_tile(rec, type_)[:, :, [2, 1, 0, 3]].astype(int)  # oracle gives B,G,R,A
        assert got.shape == want.shape == (64, 64, 4)
        assert np.array_equal(got[:, :, 3], want[:, :, 3])                    # alpha exactly
        opaque = want[:, :, 3] > 0
        # interpolated colours may round differently by one step between decoders
        assert np.abs(got[:, :, :3] - want[:, :, :3])[opaque].max() <= 1


def test_cluster_fit_beats_an_independent_simple_encoder(oracle):
    """Sanity of the restated fitters' QUALITY: Pillow ships its own (simple) BC1 encoder; on the same pixels
    the oracle's RangeFit must not be worse than it by more than a whisker and ClusterFit must be better."""
    tiles = np.concatenate([synth.smooth_tiles(3, 131), synth.uniform_tiles(2, 132)])
    for t in tiles:
        src = oracle.pal_to_argb(t[1024:], t[:1024]).reshape(64, 64, 4)[:, :, [2, 1, 0, 3]]
        buf = io.BytesIO()
        try:
            PIL.fromarray(src.astype(np.uint8), "RGBA").save(buf, format="DDS", pixel_format="DXT1")
        except Exception as e:       # older Pillow: no DDS writer
            pytest.skip(f"Pillow cannot write DXT1: {e}")
        theirs = np.asarray(PIL.open(io.BytesIO(buf.getvalue())).convert("RGBA")).astype(float)
        rmse_pillow = np.sqrt(((theirs[:, :, :3] - src[:, :, :3]) ** 2).mean())
        rmse = {}
        for q in (2, 4, 9):
            dec = oracle.decode_tile(oracle.encode_tiles(t[None], 1, q)[0], 1)[:, :, [2, 1, 0]].astype(float)
            rmse[q] = np.sqrt(((dec - src[:, :, :3]) ** 2).mean())
        assert rmse[9] <= rmse[4] + 1e-9 <= rmse[2] + 1e-9
        assert rmse[4] < rmse_pillow and rmse[2] < rmse_pillow * 1.05
