"""Known answers KA1-KA7 of SURVEY.md Appendix B, checked on the CPU oracle (and on the reference's own
glue where oracle/_ref was built)."""
import numpy as np
import pytest

K_DXT1, K_DXT3, K_DXT5 = 1, 2, 4
K_CLUSTER, K_RANGE, K_WEIGHT, K_ITER = 8, 16, 128, 256
ALL_FITS = (K_RANGE, K_CLUSTER, K_CLUSTER | K_WEIGHT, K_ITER | K_WEIGHT)


def block(pixels):
    return np.array(pixels, np.uint8).reshape(16, 4)


def test_ka1_swizzle_of_reference_build(oracle):
    ref = oracle.reference()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    assert ref.tcref_selftest() == 1      # 11 22 33 44 -> 33 22 11 44 (converter.cpp:127)


def test_ka2_pal_to_argb(oracle):
    pal = np.zeros(1024, np.uint8)
    pal[0:4] = [0x00, 0xFF, 0x00, 0x00]
    pal[4 * 9: 4 * 9 + 4] = [1, 2, 3, 77]
    idx = np.array([0, 9], np.uint8)
    out = oracle.pal_to_argb(idx, pal)
    assert out[0].tolist() == [0, 0, 0, 0]
    assert out[1].tolist() == [1, 2, 3, 255]
    pal[3] = 1                               # 00 FF 00 01 is NOT the key
    out = oracle.pal_to_argb(idx, pal)
    assert out[0].tolist() == [0x00, 0xFF, 0x00, 0xFF]


@pytest.mark.parametrize("fit", ALL_FITS)
def test_ka3_bc1_all_transparent(oracle, fit):
    out = oracle.compress_block(np.zeros((16, 4), np.uint8), K_DXT1 | fit)
    assert out.tolist() == [0, 0, 0, 0, 0xFF, 0xFF, 0xFF, 0xFF]


@pytest.mark.parametrize("fit", ALL_FITS)
def test_ka4_bc1_solid_white_and_black(oracle, fit):
    # SingleColourFit at every level (App. A.2).  lookup_5_3[255] source 0 is {start 31, end 0, error 0}
    # (the squishgen rule keeps the first (start,end) pair in lexicographic order, i.e. end = 0), so
    # white is stored as endpoints (white, black) with every pixel on the white one; WriteColourBlock3
    # swaps them to keep a <= b and flips the codes 0 -> 1.  SURVEY.md's hand-derived KA4 assumed
    # start == end; the decoded pixels are what matters and they are exactly white.
    white = block([[255, 255, 255, 255]] * 16)
    black = block([[0, 0, 0, 255]] * 16)
    assert oracle.compress_block(white, K_DXT1 | fit).tolist() == [0x00, 0x00, 0xFF, 0xFF, 0x55, 0x55, 0x55, 0x55]
    assert oracle.compress_block(black, K_DXT1 | fit).tolist() == [0] * 8
    for colour, blk in (((255, 255, 255), white), ((0, 0, 0), black)):
        rec = np.concatenate([np.array([4, 0, 4, 0], np.uint8), oracle.compress_block(blk, K_DXT1 | fit)])
        px = oracle.decode_tile(rec, 1)
        assert (px[:, :, :3] == np.array(colour[::-1], np.uint8)).all() and (px[:, :, 3] == 255).all()


@pytest.mark.parametrize("method", (K_DXT3, K_DXT5))
@pytest.mark.parametrize("fit", ALL_FITS)
def test_ka5_bc23_all_transparent_colour_half(oracle, method, fit):
    out = oracle.compress_block(np.zeros((16, 4), np.uint8), method | fit)
    assert out[8:].tolist() == [0] * 8


def test_ka6_bc2_alpha_half(oracle):
    opaque = block([[9, 8, 7, 255]] * 16)
    assert oracle.compress_block(opaque, K_DXT3 | K_RANGE)[:8].tolist() == [0xFF] * 8
    assert oracle.compress_block(np.zeros((16, 4), np.uint8), K_DXT3 | K_RANGE)[:8].tolist() == [0] * 8
    mixed = block([[9, 8, 7, 255], [0, 0, 0, 0]] * 8)
    assert oracle.compress_block(mixed, K_DXT3 | K_RANGE)[:8].tolist() == [0x0F] * 8


def test_ka7_bc3_alpha_half(oracle):
    opaque = block([[9, 8, 7, 255]] * 16)
    assert oracle.compress_block(opaque, K_DXT5 | K_RANGE)[:8].tolist() == [0, 5] + [0xFF] * 6
    assert oracle.compress_block(np.zeros((16, 4), np.uint8), K_DXT5 | K_RANGE)[:8].tolist() == [0, 5] + [0] * 6
    # mixed: transparent -> code 0, opaque -> code 7 (codes5[7] == 255)
    mixed = block([[0, 0, 0, 0]] * 8 + [[9, 8, 7, 255]] * 8)
    out = oracle.compress_block(mixed, K_DXT5 | K_RANGE)[:8]
    bits = int.from_bytes(bytes(out[2:8].tolist()), "little")
    codes = [(bits >> (3 * i)) & 7 for i in range(16)]
    assert out[0] == 0 and out[1] == 5 and codes == [0] * 8 + [7] * 8


def test_quality_to_flags_map(oracle):
    # ConverterDxt::getFlags, converter_dxt.cpp:179-212
    for t, m in ((1, K_DXT1), (2, K_DXT3), (3, K_DXT5)):
        assert oracle.squish_flags(t, 0) == m | K_RANGE
        assert oracle.squish_flags(t, 2) == m | K_RANGE
        assert oracle.squish_flags(t, 3) == m | K_CLUSTER
        assert oracle.squish_flags(t, 4) == m | K_CLUSTER
        assert oracle.squish_flags(t, 5) == m | K_CLUSTER | K_WEIGHT
        assert oracle.squish_flags(t, 6) == m | K_CLUSTER | K_WEIGHT
        assert oracle.squish_flags(t, 7) == m | K_ITER | K_WEIGHT
        assert oracle.squish_flags(t, 9) == m | K_ITER | K_WEIGHT


def test_required_space(oracle):
    # ConverterDxt::getRequiredSpace, converter_dxt.cpp:41-56
    assert oracle.record_size(1, 64, 64) == 2052
    assert oracle.record_size(2, 64, 64) == 4100
    assert oracle.record_size(3, 64, 64) == 4100
    assert oracle.record_size(1, 1, 1) == 12
    assert oracle.record_size(3, 5, 3) == 4 + 2 * 16
    assert oracle.record_size(1, 0, 4) == 0
    assert oracle.record_size(4, 64, 64) == 0


def test_partial_block_padding_is_transparent_black(oracle):
    # PadBlock(useCopy=false): converter.cpp:62-97; a 5x3 tile has a 1x3 block at bx=1
    pal = np.zeros(1024, np.uint8)
    pal[4:8] = [10, 20, 30, 0]
    idx = np.ones(15, np.uint8)
    rgba = oracle.gather_block(pal, idx, 5, 3, 1, 0).reshape(4, 4, 4)
    assert rgba[0, 0].tolist() == [30, 20, 10, 255]       # swizzled to R,G,B,A
    assert rgba[0, 1:].max() == 0 and rgba[3].max() == 0  # padding
    assert rgba[2, 0].tolist() == [30, 20, 10, 255]
