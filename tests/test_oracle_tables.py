"""SingleColourFit lookup tables: spot values and brute-force optimality."""
import numpy as np
import pytest


def expand(v, bits):
    return (v << (8 - bits)) | (v >> (2 * bits - 8))


def test_spot_values_lookup_5_3(oracle):
    t = oracle.single_colour_table(5, 3).tolist()
    assert t[3] == [[0, 0, 3], [0, 1, 1]]
    assert t[4] == [[0, 0, 4], [0, 1, 0]]
    assert t[5] == [[1, 0, 3], [0, 1, 1]]
    assert t[7] == [[1, 0, 1], [0, 2, 1]]
    assert t[8] == [[1, 0, 0], [0, 2, 0]]
    assert t[255][0] == [31, 0, 0]


def test_spot_values_lookup_6_3(oracle):
    t = oracle.single_colour_table(6, 3).tolist()
    assert t[:5] == [[[0, 0, 0], [0, 0, 0]], [[0, 0, 1], [0, 1, 1]], [[0, 0, 2], [0, 1, 0]],
                     [[1, 0, 1], [0, 2, 1]], [[1, 0, 0], [0, 2, 0]]]


@pytest.mark.parametrize("bits", (5, 6))
@pytest.mark.parametrize("colours", (3, 4))
def test_tables_are_error_optimal_and_consistent(oracle, bits, colours):
    t = oracle.single_colour_table(bits, colours).astype(int)
    n = 1 << bits
    s, e = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a, b = expand(s, bits), expand(e, bits)
    codes = [a, (a + b) // 2 if colours == 3 else (2 * a + b) // 3]
    for target in range(256):
        for k in range(2):
            start, end, err = t[target, k]
            best = np.abs(codes[k] - target).min()
            assert err == best, (target, k)
            got = codes[k][start, end]
            assert abs(int(got) - target) == err, (target, k)
