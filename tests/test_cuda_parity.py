"""GPU parity tests proper: the CUDA path, called through the C ABI (include/tcb200.h via ctypes),
against the CPU oracle on the same seeded inputs.  The bar is bit-exact for EVERY fitter -- the
cluster sweeps keep the reference's summation order (prefix sums built left to right, arg-min with
lowest-candidate tie-break), so the north_star's 99.9 % allowance for ClusterFit is not needed:
the tests demand 100 % identical blocks and therefore zero RMSE drift."""
import os

import numpy as np
import pytest

from conftest import QUALITY_CLASSES
from tileconv_b200 import synth

pytestmark = pytest.mark.gpu

HOST_THREADS = os.cpu_count() or 8
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def encoder(type_, quality, **kw):
    from tileconv_b200 import Encoder
    return Encoder(type_, quality, **kw)


def assert_blocks_equal(got, want, type_, what):
    bs = 8 if type_ == 1 else 16
    n = got.shape[0]
    assert np.array_equal(got[:, :4], want[:, :4]), f"{what}: tile headers differ"
    g = got[:, 4:].reshape(n, -1, bs); w = want[:, 4:].reshape(n, -1, bs)
    diff = int((g != w).any(axis=2).sum())
    assert diff == 0, f"{what}: {diff} of {g.shape[0] * g.shape[1]} blocks differ from the oracle"


@pytest.mark.parametrize("type_", (1, 2, 3))
@pytest.mark.parametrize("quality", tuple(range(10)))     # every -q level, incl. 1 and 8
def test_bit_exact_on_seeded_tiles(oracle, type_, quality):
    sets = {"S-U": synth.uniform_tiles(6, 41), "S-G": synth.smooth_tiles(10, 42), "S-K": synth.keyed_tiles(8, 43)}
    with encoder(type_, quality) as enc:
        for name, tiles in sets.items():
            got = enc.encode(tiles)
            want = oracle.encode_tiles(tiles, type_, quality, threads=8)
            assert_blocks_equal(got, want, type_, f"BC{type_} -q{quality} {name}")


@pytest.mark.parametrize("type_", (1, 2, 3))
@pytest.mark.parametrize("quality", (2, 9))
def test_bit_exact_on_ragged_tiles(oracle, type_, quality):
    tiles, wh = synth.ragged_tiles(24, 44)
    with encoder(type_, quality) as enc:
        got = enc.encode(tiles, wh)
    want = oracle.encode_tiles_wh(tiles, wh, type_, quality)
    for i in range(len(tiles)):
        size = oracle.record_size(type_, int(wh[i, 0]), int(wh[i, 1]))
        assert np.array_equal(got[i, :size], want[i, :size]), (i, wh[i].tolist())


def test_config1_bc1_q9_1024_tiles(oracle):
    """BASELINE configs[0]: 1024-tile headerless TIS, BC1 -q 9 (half S-U, half S-G)."""
    tiles = synth.mixed_tiles(1024, 1234)
    with encoder(1, 9) as enc:
        got = enc.encode(tiles)
    want = oracle.encode_tiles(tiles, 1, 9, threads=HOST_THREADS)
    assert_blocks_equal(got, want, 1, "config 1 (all 1024 tiles)")
    # and a size-independent property: byte-different copies with identical pixels
    rot = synth.rotate_palettes(tiles, 91)
    with encoder(1, 9) as enc:
        assert np.array_equal(enc.encode(rot), got)


def test_config2_bc1_q2_bit_exact_1024_tiles(oracle):
    """BASELINE configs[1]: same TIS at BC1 -q 2 (RangeFit): bit-exact over all 1024 tiles."""
    tiles = synth.mixed_tiles(1024, 1234)
    with encoder(1, 2) as enc:
        got = enc.encode(tiles)
    assert_blocks_equal(got, oracle.encode_tiles(tiles, 1, 2, threads=16), 1, "config 2")


def test_config3_bc3_q7_keyed(oracle):
    """BASELINE configs[2]: alpha-keyed palette entry, BC3 with alpha-weighted iterative ClusterFit."""
    tiles = synth.keyed_tiles(1024, 3456)
    with encoder(3, 7) as enc:
        got = enc.encode(tiles)
    assert_blocks_equal(got, oracle.encode_tiles(tiles, 3, 7, threads=HOST_THREADS), 3, "config 3 (all 1024 tiles)")
    # KA7: every BC3 alpha half starts 00 05
    blocks = got[:, 4:].reshape(len(tiles), 256, 16)
    assert (blocks[:, :, 0] == 0).all() and (blocks[:, :, 1] == 5).all()


def test_config4_large_mos_bc2_q9(oracle):
    """BASELINE configs[3]: 5120x3840 area map = 80x60 = 4800 full tiles, BC2 -q 9.  All 4800 tiles are
    compared with the oracle, plus properties of the whole map (determinism across chunkings, alpha
    halves from the closed form KA6)."""
    tiles = np.concatenate([synth.smooth_tiles(3600, 4001), synth.keyed_tiles(1200, 4002)])
    with encoder(2, 9, max_batch_tiles=1000) as enc:
        got = enc.encode(tiles)
    with encoder(2, 9, max_batch_tiles=4800) as enc:
        again = enc.encode(tiles)
    assert np.array_equal(got, again)
    assert_blocks_equal(got, oracle.encode_tiles(tiles, 2, 9, threads=HOST_THREADS), 2, "config 4 (all 4800 tiles)")
    # KA6 on the whole map: alpha nibble F for opaque, 0 for keyed pixels
    px_alpha = np.ones((4800, 64, 64), bool)
    px_alpha[3600:] = tiles[3600:, 1024:].reshape(-1, 64, 64) != 0
    nib = np.where(px_alpha, 0xF, 0).astype(np.uint8)
    blocks = nib.reshape(4800, 16, 4, 16, 4).transpose(0, 1, 3, 2, 4).reshape(4800, 256, 16)
    want_alpha = (blocks[:, :, 0::2] | (blocks[:, :, 1::2] << 4)).astype(np.uint8)
    assert np.array_equal(got[:, 4:].reshape(4800, 256, 16)[:, :, :8], want_alpha)


@pytest.mark.parametrize("type_,quality", ((1, 9), (3, 5)))
def test_rmse_matches_oracle_exactly(oracle, type_, quality):
    """north_star tolerance: per-tile RGB RMSE within 0.05 of the reference's.  Identical blocks give
    identical RMSE; the test states the tolerance anyway."""
    from test_oracle_properties import tile_rmse
    tiles = np.concatenate([synth.uniform_tiles(4, 51), synth.keyed_tiles(4, 52)])
    with encoder(type_, quality) as enc:
        got = enc.encode(tiles)
    want = oracle.encode_tiles(tiles, type_, quality, threads=8)
    delta = np.abs(tile_rmse(oracle, tiles, got, type_) - tile_rmse(oracle, tiles, want, type_))
    assert delta.max() <= 0.05


def test_edge_cases(oracle):
    with encoder(1, 9) as enc:
        assert enc.encode(np.zeros((0, 5120), np.uint8)).shape == (0, 2052)
        # fully transparent tile (KA3) and a solid tile
        t = np.zeros((2, 5120), np.uint8)
        t[0, 0:4] = synth.KEY_ENTRY
        t[1, 0:1024] = np.tile(np.array([1, 2, 3, 0], np.uint8), 256)
        got = enc.encode(t)
        assert np.array_equal(got, oracle.encode_tiles(t, 1, 9))
        assert (got[0, 4:].reshape(256, 8) == np.array([0, 0, 0, 0, 255, 255, 255, 255], np.uint8)).all()
        # 1x1 .. 4x4 tiles
        wh = np.array([[1, 1], [4, 4]], np.uint16)
        assert np.array_equal(enc.encode(t, wh)[:, :12], oracle.encode_tiles_wh(t, wh, 1, 9)[:, :12])
    from tileconv_b200 import TcbError
    with encoder(1, 9) as enc:
        with pytest.raises(TcbError):
            enc.encode(np.zeros((1, 5120), np.uint8), np.array([[0, 64]], np.uint16))
        with pytest.raises(TcbError):
            enc.encode(np.zeros((1, 5120), np.uint8), np.array([[65, 1]], np.uint16))


def nan_axis_tile():
    """Blocks whose principal axis is NaN: two opaque pixels (255,0,0) and (0,255,0), everything else
    keyed transparent.  Their covariance is [[.5,-.5,0],[-.5,.5,0],[0,0,0]], the power iteration starts
    from (1,1,1), gets the zero vector and divides by its maximum: 0 * (1/0) = NaN (App. A.4).  The
    reference then sorts NaN projections, i.e. leaves the order alone; the CUDA path has a dedicated
    branch for non-finite projections."""
    t = np.zeros((1, 5120), np.uint8)
    t[0, 0:4] = synth.KEY_ENTRY
    t[0, 4:8] = [0, 0, 255, 0]        # index 1: red   (B,G,R,x)
    t[0, 8:12] = [0, 255, 0, 0]       # index 2: green
    t[0, 12:16] = [255, 0, 0, 0]      # index 3: blue
    idx = t[0, 1024:].reshape(64, 64)
    rng = np.random.default_rng(5)
    for by in range(16):
        for bx in range(16):
            cells = rng.permutation(16)[:2]
            pair = [(1, 2), (2, 1), (1, 3), (3, 2)][(by * 16 + bx) % 4]
            for c, v in zip(cells, pair):
                idx[4 * by + c // 4, 4 * bx + c % 4] = v
    return t


@pytest.mark.parametrize("type_,quality", ((1, 4), (1, 9), (3, 9), (2, 6)))
def test_nan_principal_axis_blocks(oracle, type_, quality):
    t = nan_axis_tile()
    with encoder(type_, quality) as enc:
        got = enc.encode(t)
    assert_blocks_equal(got, oracle.encode_tiles(t, type_, quality), type_, "NaN-axis blocks")


def black_and_transparent_tiles():
    """Opaque black pixels and keyed-transparent pixels share the colour (0,0,0) in BC2/BC3, so with
    kWeightColourByAlpha their weights add up as (256*opaque + transparent)/256 on ONE point (App. A.3);
    in BC1 the transparent ones leave the set.  Mixed with a few other colours, every ratio 0..16."""
    rng = np.random.default_rng(17)
    t = np.zeros((3, 5120), np.uint8)
    for k in range(3):
        pal = rng.integers(0, 256, (256, 4), dtype=np.uint8); pal[:, 3] = 0
        pal[0] = synth.KEY_ENTRY          # index 0: transparent
        pal[1] = [0, 0, 0, 0]             # index 1: opaque black
        pal[2] = [0, 0, 0, 0]             # index 2: a second palette entry with the same colour
        t[k, :1024] = pal.reshape(-1)
        idx = t[k, 1024:].reshape(64, 64)
        for by in range(16):
            for bx in range(16):
                n_t = (by + k) % 17 if (by + k) % 17 <= 16 else 0
                cells = rng.permutation(16)
                vals = np.empty(16, np.uint8)
                vals[:] = rng.integers(3, 3 + 1 + bx // 2, 16)          # 1..8 other colours
                vals[cells[:min(n_t, 16)]] = 0
                n_b = bx % 9
                vals[cells[min(n_t, 16):min(n_t, 16) + n_b]] = rng.integers(1, 3, min(n_b, 16 - min(n_t, 16)))
                idx[4 * by:4 * by + 4, 4 * bx:4 * bx + 4] = vals.reshape(4, 4)
    return t


@pytest.mark.parametrize("type_,quality", ((1, 9), (2, 4), (2, 6), (3, 5), (3, 9), (3, 2)))
def test_black_merges_with_transparent_weights(oracle, type_, quality):
    t = black_and_transparent_tiles()
    with encoder(type_, quality) as enc:
        got = enc.encode(t)
    assert_blocks_equal(got, oracle.encode_tiles(t, type_, quality, threads=4), type_, "black+transparent")


@pytest.mark.parametrize("type_,quality", ((1, 2), (1, 4), (1, 9), (3, 9)))
def test_corner_case_blocks(oracle, type_, quality):
    """Two-colour, collinear, on-grid, one-LSB-apart, outlier, grey and duplicated-palette blocks."""
    t = synth.corner_case_tiles(24, 5678)
    with encoder(type_, quality) as enc:
        got = enc.encode(t)
    assert_blocks_equal(got, oracle.encode_tiles(t, type_, quality, threads=16), type_, "corner-case blocks")


def test_inlined_reciprocal_is_ieee_exact():
    """The cluster solve replaces `1.0f/x` by MUFU.RCP + one Newton step without the compiler's range
    check; every binary32 value of the range the solve can produce must give the IEEE result."""
    with encoder(1, 9) as enc:
        assert enc.selftest_reciprocal() == 0


def test_pal_to_argb_matches_oracle(oracle):
    tiles = np.concatenate([synth.uniform_tiles(3, 61), synth.keyed_tiles(3, 62)])
    with encoder(1, 2) as enc:
        got = enc.pal_to_argb(tiles)
    for t, g in zip(tiles, got):
        assert np.array_equal(g, oracle.pal_to_argb(t[1024:], t[:1024]))


def test_device_slab_and_sharded_paths_agree(oracle):
    """encode() == encode_device() == slab interface == two tile-range shards concatenated."""
    import torch
    from tileconv_b200.sharding import tile_range
    tiles = synth.mixed_tiles(300, 71)
    with encoder(1, 9, max_batch_tiles=128) as enc:
        want = enc.encode(tiles)
        d_in = torch.from_numpy(tiles.reshape(-1)).cuda()
        d_out = torch.zeros(300 * enc.stride, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()          # the context launches on its own (non-blocking) stream
        enc.encode_device(d_in.data_ptr(), None, 300, d_out.data_ptr())
        torch.cuda.synchronize()
        # the kernel ran on the context's own stream: wait for it through the context
        ms = enc.time_kernel_ms(d_in.data_ptr(), None, 300, d_out.data_ptr(), 1)
        assert ms > 0
        assert np.array_equal(d_out.cpu().numpy().reshape(300, -1), want)
        # slabs: three batches in flight
        outs = []
        cap = enc.slab_capacity()
        assert cap == 128 and enc.slab_count() >= 2
        for s, lo in enumerate(range(0, 300, cap)):
            m = min(cap, 300 - lo)
            enc.slab_input(s)[:m] = tiles[lo:lo + m]
            enc.slab_sizes(s)[:m] = 64
            enc.slab_submit(s, m)
        for s, lo in enumerate(range(0, 300, cap)):
            m = min(cap, 300 - lo)
            enc.slab_wait(s)
            assert enc.slab_ready(s)
            outs.append(enc.slab_output(s)[:m].copy())
        assert np.array_equal(np.concatenate(outs), want)
        # two shards
        parts = []
        for rank in range(2):
            lo, hi = tile_range(rank, 2, 300)
            parts.append(enc.encode(tiles[lo:hi]))
        assert np.array_equal(np.concatenate(parts), want)
        assert enc.launch_count() >= 8


@pytest.mark.parametrize("type_", (1, 2, 3))
def test_decode_matches_restated_reference_decoder(oracle, type_):
    """Next-row N3: BCn decode on the GPU against the oracle's restatement of decodeBlockDxt1/3/5 (which
    tests/test_oracle_vs_reference.py pins to the reference's own decoders), full and ragged tiles,
    both BC3 alpha modes."""
    tiles, wh = synth.ragged_tiles(12, 111)
    full = np.concatenate([synth.keyed_tiles(3, 112), synth.uniform_tiles(2, 113)])
    with encoder(type_, 4) as enc:
        rec_r = enc.encode(tiles, wh)
        rec_f = enc.encode(full)
        for quirk in (False, True):
            px_r = enc.decode(rec_r, quirk)
            px_f = enc.decode(rec_f, quirk)
            for i in range(len(tiles)):
                w, h = int(wh[i, 0]), int(wh[i, 1])
                want = oracle.decode_tile(rec_r[i], type_, quirk)
                assert np.array_equal(px_r[i, :w * h].reshape(h, w, 4), want), (i, w, h, quirk)
            for i in range(len(full)):
                assert np.array_equal(px_f[i].reshape(64, 64, 4), oracle.decode_tile(rec_f[i], type_, quirk))


# ---- libsquish-version knobs (tcb200_create_ex) ---------------------------------------------------------
VARIANTS = {"perceptual": dict(metric=(0.2126, 0.7152, 0.0722)), "sse": dict(sse=True),
            "perceptual+sse": dict(metric=(0.2126, 0.7152, 0.0722), sse=True), "odd-metric": dict(metric=(0.5, 1.0, 0.25))}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
@pytest.mark.parametrize("type_,quality", ((1, 2), (1, 4), (1, 9), (2, 9), (3, 6), (3, 8)))
def test_variants_bit_exact_with_the_oracle_in_the_same_variant(oracle, variant, type_, quality):
    """The metric and SSE-reciprocal knobs select another libsquish build's arithmetic; the CUDA path must
    follow the oracle configured the same way, block for block (the SSE estimate is the host CPU's own
    rcpps on both sides)."""
    kw = VARIANTS[variant]
    sets = {"S-U": synth.uniform_tiles(6, 141), "S-G": synth.smooth_tiles(8, 142), "S-K": synth.keyed_tiles(8, 143),
            "S-C": synth.corner_case_tiles(6, 144), "NaN-axis": nan_axis_tile()}
    with encoder(type_, quality, **kw) as enc:
        opt = enc.options()
        assert opt.reciprocal == (1 if kw.get("sse") else 0)
        with oracle.variant(**kw):
            for name, tiles in sets.items():
                got = enc.encode(tiles)
                want = oracle.encode_tiles(tiles, type_, quality, threads=8)
                assert_blocks_equal(got, want, type_, f"{variant} BC{type_} -q{quality} {name}")
    # and the knobs must actually change something on the GPU side too (S-U at a cluster level)
    if quality >= 3 and "metric" in kw:
        with encoder(type_, quality) as enc:
            base = enc.encode(sets["S-U"])
        with encoder(type_, quality, **kw) as enc:
            assert not np.array_equal(enc.encode(sets["S-U"]), base)


def test_default_options_are_the_plain_create(oracle):
    tiles = synth.mixed_tiles(8, 151)
    with encoder(1, 9) as a, encoder(1, 9, metric=(1.0, 1.0, 1.0)) as b:
        assert np.array_equal(a.encode(tiles), b.encode(tiles))
        assert list(b.options().metric) == [1.0, 1.0, 1.0] and b.options().reciprocal == 0


def test_rcpps_estimate_matches_the_host_instruction(oracle):
    """TCB200_RECIP_SSE replaces the host's rcpps by a table measured at create time; compare it with the
    instruction itself over every exponent, a dense mantissa sample, both signs and the special values."""
    rng = np.random.default_rng(7)
    bits = np.concatenate([
        rng.integers(0, 1 << 32, 200000, dtype=np.uint64).astype(np.uint32),
        (np.arange(0, 256, dtype=np.uint32)[:, None] << 23 | rng.integers(0, 1 << 23, (256, 64), dtype=np.uint32)).ravel(),
        (np.uint32(0x3f800000) | np.arange(0, 1 << 23, 257, dtype=np.uint32)),
        np.array([0, 0x80000000, 0x7f800000, 0xff800000, 1, 0x007fffff, 0x00800000, 0x7f7fffff, 0x7e800000, 0x7e800001], np.uint32)])
    vals = bits.view(np.float32)
    with encoder(1, 9, sse=True) as enc:
        got = enc.selftest_rcpps(vals)
    want = oracle.rcp_estimate(vals)
    nan = np.isnan(vals)
    assert np.isnan(got[nan]).all() and np.isnan(want[nan]).all()
    assert np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32))


def test_encode_device_rejects_bad_tile_sizes(oracle):
    """d_wh lives on the device, so tcb200_encode_device cannot validate it: the kernel must neither read nor
    write out of bounds for w/h outside 1..64, and the context must report the rejected records."""
    import torch
    tiles = synth.keyed_tiles(6, 161)
    wh = np.array([[64, 64], [0, 7], [65, 64], [64, 65535], [3, 5], [64, 64]], np.uint16)
    for type_, quality in ((1, 9), (3, 2)):
        with encoder(type_, quality) as enc:
            good = [0, 4, 5]
            want = enc.encode(tiles[good], wh[good])
            d_in = torch.from_numpy(tiles.reshape(-1)).cuda()
            d_wh = torch.from_numpy(wh.view(np.int16).reshape(-1)).cuda()
            d_out = torch.full((6 * enc.stride,), 0xAB, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            enc.encode_device(d_in.data_ptr(), d_wh.data_ptr(), 6, d_out.data_ptr())
            assert enc.rejected_tiles() == 3
            out = d_out.cpu().numpy().reshape(6, enc.stride)
            for k, i in enumerate(good):
                size = oracle.record_size(type_, int(wh[i, 0]), int(wh[i, 1]))
                assert np.array_equal(out[i, :size], want[k, :size])
            for i in (1, 2, 3):
                assert (out[i, :4] == 0).all() and (out[i, 4:] == 0xAB).all()
            enc.encode_device(d_in.data_ptr(), None, 6, d_out.data_ptr())
            assert enc.rejected_tiles() == 3


# ---- tcb200_encode_multi: the in-process multi-GPU entry point -------------------------------------------
def test_encode_multi_shards_match_single_context(oracle):
    """Contiguous tile-range shards over several contexts == one context, for an uneven split and ragged tiles.
    On a one-GPU box the contexts share device 0 (the sharding and the disjoint output slices are the same
    code); with more devices visible each context sits on its own GPU."""
    from tileconv_b200 import capi, device_count, encode_multi
    ndev = device_count()
    tiles, wh = synth.ragged_tiles(1001, 171)
    wh = np.ascontiguousarray(wh, np.uint16)
    for type_, quality, nctx in ((1, 9, 3), (3, 2, 2), (2, 4, max(2, min(ndev, 8)))):
        encs = [encoder(type_, quality, device=g % ndev, max_batch_tiles=128) for g in range(nctx)]
        try:
            want = encs[0].encode(tiles, wh)
            out = np.zeros_like(want)
            encode_multi(encs, tiles.ctypes.data, wh.ctypes.data, len(tiles), out.ctypes.data)
            for i in range(len(tiles)):
                size = oracle.record_size(type_, int(wh[i, 0]), int(wh[i, 1]))
                assert np.array_equal(out[i, :size], want[i, :size]), (type_, quality, i)
            # n smaller than the number of contexts, and n = 0
            out2 = np.zeros((1, encs[0].stride), np.uint8)
            encode_multi(encs, tiles.ctypes.data, None, 1, out2.ctypes.data)
            assert np.array_equal(out2, encs[0].encode(tiles[:1]))
            encode_multi(encs, None, None, 0, None)
        finally:
            for e in encs:
                e.close()


def test_encode_multi_rejects_mismatched_contexts():
    from tileconv_b200 import TcbError, encode_multi
    tiles = synth.uniform_tiles(4, 181)
    out = np.zeros((4, 2052), np.uint8)
    with encoder(1, 9) as a, encoder(1, 4) as b, encoder(1, 9, metric=(0.5, 1.0, 1.0)) as c:
        for bad in ([a, b], [a, c], [a, a]):
            with pytest.raises(TcbError):
                encode_multi(bad, tiles.ctypes.data, None, 4, out.ctypes.data)
        with pytest.raises(TcbError):       # bad sizes are caught by the shard's own validation
            encode_multi([a], tiles.ctypes.data, np.array([[0, 1]] * 4, np.uint16).ctypes.data, 4, out.ctypes.data)


def test_pruned_sweeps_never_skip_a_candidate_that_could_win():
    """bc_device.cuh run_sweep skips candidate partitions whose lower bound proves they cannot beat the incumbent.
    The audit build of the library (-DTCB_PRUNE_AUDIT=1, made by __graft_entry__.build()) evaluates every skipped
    candidate as well: none may have an error at or below the incumbent, and the rounding shortfall of the bound must
    stay far inside the margin.  Runs in its own process (the audit build replaces the library)."""
    import json, subprocess, sys
    lib = os.path.join(ROOT, "tileconv_b200", "libtcb200_audit.so")
    assert os.path.exists(lib), "libtcb200_audit.so is missing: run __graft_entry__.build()"
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prune_audit.py"), "24", "31"], capture_output=True, text=True, timeout=600)
    rows = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and rows, res.stderr[-800:]
    assert rows[-1] == {"wrongly_skipped_total": 0}
    per_class = [r for r in rows if "class" in r]
    assert len(per_class) == 48 and sum(r["variant"] != "default" for r in per_class) == 12      # default and VARIANT kernels
    assert all(r["wrongly_skipped"] == 0 for r in per_class)
    assert max(r["worst_shortfall_over_margin"] for r in per_class) < 0.25
    # the bound does real work: most 4-colour candidates of ordinary tiles are skipped
    assert all(r["skipped_frac"] > 0.4 for r in per_class if r["class"] in ("S-U", "S-G", "S-K"))
