"""The lower bound the CUDA path prunes 4-colour cluster candidates by (tileconv_b200/csrc/bc_device.cuh run_sweep,
DESIGN.md s4 "Pruned sweep"), checked on the CPU: oracle/squish_oracle.cpp restates the bound test in binary32 and
compares it with the error it computes for EVERY candidate, using the tightest incumbent a skip could ever see.  A
candidate that would be skipped although its error is not above the incumbent would be a candidate the CUDA path
could wrongly drop.  (The GPU suite runs the same audit inside the kernels: test_pruned_sweeps_never_skip_...)"""
import numpy as np
import pytest

from oracle import pyoracle
from tileconv_b200 import synth

CLASSES = {
    "S-U": lambda n, seed: synth.uniform_tiles(n, seed),
    "S-G": lambda n, seed: synth.smooth_tiles(n, seed),
    "S-K": lambda n, seed: synth.keyed_tiles(n, seed),
    "S-C": lambda n, seed: synth.corner_case_tiles(n, seed),
}


@pytest.mark.parametrize("name", sorted(CLASSES))
@pytest.mark.parametrize("type_,quality", ((1, 9), (3, 4), (3, 9)))      # BC1 iterative, BC3 plain and alpha-weighted iterative
def test_bound_never_exceeds_the_error_it_bounds(name, type_, quality):
    tiles = CLASSES[name](3, 4242 + type_ * 10 + quality)
    a = pyoracle.bound_audit(tiles, type_, quality)
    assert a["tested"] > 10000
    assert a["wrongly_skipped"] == 0
    # rounding can put a binary32 error below the raw bound; the margin must dwarf the worst case
    assert a["worst_shortfall_over_margin"] < 0.25
    if name != "S-C":
        assert a["skipped"] > 0.4 * a["tested"]          # and the bound is tight enough to matter


def test_bound_needs_its_margin_only_for_rounding():
    """With the margin scaled down 1000x the audit may start to find binary32 errors below the bound -- that is what
    the margin is for -- but with the shipped scale the worst shortfall stays under 2 % of it on a larger sample."""
    tiles = np.concatenate([synth.smooth_tiles(6, 99), synth.keyed_tiles(6, 98)])
    a = pyoracle.bound_audit(tiles, 3, 9)
    assert a["wrongly_skipped"] == 0 and a["worst_shortfall_over_margin"] < 0.02


@pytest.mark.parametrize("metric,sse", ((pyoracle.PERCEPTUAL, False), (None, True), (pyoracle.PERCEPTUAL, True), ((3.0, 0.5, 2.0), False)))
def test_bound_holds_for_the_libsquish_version_knobs(metric, sse):
    """The VARIANT kernels prune too: per-channel metric weights enter the bound terms and widen the margin; the SSE
    arithmetic only changes how the endpoints are found, which the bound does not depend on."""
    tiles = np.concatenate([synth.uniform_tiles(2, 71), synth.smooth_tiles(3, 72), synth.keyed_tiles(3, 73), synth.corner_case_tiles(2, 74)])
    pyoracle.set_variant(metric, sse)
    try:
        for type_, quality in ((1, 9), (3, 9)):
            a = pyoracle.bound_audit(tiles, type_, quality)
            assert a["tested"] > 10000 and a["skipped"] > 0.3 * a["tested"]
            assert a["wrongly_skipped"] == 0
            assert a["worst_shortfall_over_margin"] < 0.25
    finally:
        pyoracle.set_variant(None, False)
