import numpy as np

from tileconv_b200 import synth


def test_generators_are_deterministic():
    for gen in (synth.uniform_tiles, synth.smooth_tiles, synth.keyed_tiles, synth.mixed_tiles):
        a, b = gen(3), gen(3)
        assert a.shape == (3, 5120) and a.dtype == np.uint8 and np.array_equal(a, b)


def test_keyed_tiles_have_the_key_entry_and_transparent_pixels():
    t = synth.keyed_tiles(3)
    assert (t[:, 0:4] == synth.KEY_ENTRY).all()
    assert 0.2 < (t[:, 1024:] == 0).mean() < 0.5
    assert not ((synth.smooth_tiles(3)[:, 0:4] == synth.KEY_ENTRY).all(axis=1)).any()


def test_rotated_palettes_keep_the_pixels(oracle):
    t = synth.mixed_tiles(4)
    r = synth.rotate_palettes(t, 37)
    assert not np.array_equal(t, r)
    for a, b in zip(t, r):
        assert np.array_equal(oracle.pal_to_argb(a[1024:], a[:1024]), oracle.pal_to_argb(b[1024:], b[:1024]))


def test_ragged_tiles_pack_rows_at_tile_pitch():
    tiles, wh = synth.ragged_tiles(6)
    assert wh.shape == (6, 2) and wh.min() >= 1 and wh.max() <= 64
    for t, (w, h) in zip(tiles, wh):
        assert (t[1024 + int(w) * int(h):] == 0).all()
