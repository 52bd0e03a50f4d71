"""The C++ host layer (tileconv_b200/host): ConverterDxtCuda behind ConverterFactory, TileBatchQueue
behind createThreadPool()'s interface, and the TIS/MOS -> TBC/MBC drivers.  GPU tests; expected bytes
come from the CPU oracle and, where the reference CLI was built (oracle/_ref/tileconv_ref), from the
reference program itself."""
import ctypes
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from tileconv_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTLIB = os.path.join(ROOT, "tileconv_b200", "libtcb200_host.so")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "tileconv_ref")
CUDA_CLI = os.path.join(ROOT, "oracle", "_ref", "tileconv_cuda")
TBC_ENCODE = os.path.join(ROOT, "tileconv_b200", "tbc_encode")


def test_host_library_exports():
    lib = ctypes.CDLL(HOSTLIB)
    for name in ("tcbh_convert_tile", "tcbh_queue_encode", "tcbh_queue_encode_mixed", "tcbh_encode_file", "tcbh_last_error"):
        assert hasattr(lib, name)


@pytest.fixture(scope="module")
def host():
    lib = ctypes.CDLL(HOSTLIB)
    lib.tcbh_last_error.restype = ctypes.c_char_p
    return lib


def p(a):
    return ctypes.c_void_p(a.ctypes.data)


# ---- file helpers --------------------------------------------------------------------------------
def write_tis(path, tiles, header=True):
    with open(path, "wb") as f:
        if header:
            f.write(b"TIS V1  " + struct.pack("<IIII", len(tiles), 0x1400, 0x18, 0x40))
        f.write(tiles.tobytes())


def make_mos(width, height, seed, compressed=False):
    """MOS V1: header, palettes, offset table, tile data with row pitch = tile width (graphics.cpp:296-342)."""
    cols, rows = (width + 63) // 64, (height + 63) // 64
    n = cols * rows
    base = synth.keyed_tiles(n, seed)
    pal = base[:, :1024].tobytes()
    table, data, wh, packed = [], b"", [], np.zeros((n, 5120), np.uint8)
    for i in range(n):
        r, c = divmod(i, cols)
        w, h = min(64, width - c * 64), min(64, height - r * 64)
        img = base[i, 1024:].reshape(64, 64)[:h, :w].reshape(-1)
        table.append(len(data)); data += img.tobytes(); wh.append((w, h))
        packed[i, :1024] = base[i, :1024]; packed[i, 1024:1024 + w * h] = img
    mos = b"MOS V1  " + struct.pack("<HHHHII", width, height, cols, rows, 64, 24) + pal + struct.pack(f"<{n}I", *table) + data
    if compressed:
        mos = b"MOSCV1  " + struct.pack("<I", len(mos)) + zlib.compress(mos, 9)
    return mos, packed, np.array(wh, np.uint16)


def expected_stream(oracle, header, tiles, wh, type_, quality, deflate):
    out = header
    for i in range(len(tiles)):
        w, h = (64, 64) if wh is None else (int(wh[i, 0]), int(wh[i, 1]))
        rec = oracle.encode_tile(tiles[i, :1024], tiles[i, 1024:1024 + w * h], w, h, type_, quality).tobytes()
        if deflate:
            rec = zlib.compress(rec, 9)
        out += struct.pack("<I", len(rec)) + rec
    return out


# ---- type 0 (raw passthrough) and the queue's host lane: no GPU involved ---------------------------
def test_raw_passthrough_through_queue_and_plugin(host, oracle):
    """"-t 0": the tile record is u16 w, u16 h, palette, indices (converter_raw.cpp:48-62).  These tiles
    never touch the GPU: the queue runs the tile's own functor on host workers, like the reference pool."""
    tiles, wh = synth.ragged_tiles(20, 90)
    stride = 4 + 1024 + 4096
    out = np.zeros((20, stride * 2), np.uint8); sizes = np.zeros(20, np.int32)
    assert host.tcbh_queue_encode(0, 9, 0, p(tiles), p(wh), 20, p(out), stride * 2, p(sizes), 8, 0, 3) == 0
    ref = oracle.reference()
    for i in range(20):
        w, h = int(wh[i, 0]), int(wh[i, 1])
        want = bytes([w, 0, h, 0]) + tiles[i, :1024 + w * h].tobytes()
        assert sizes[i] == len(want) and out[i, :sizes[i]].tobytes() == want
        one = np.zeros(stride, np.uint8)
        pal, idx = tiles[i, :1024].copy(), tiles[i, 1024:].copy()
        assert host.tcbh_convert_tile(0, 9, p(pal), p(idx), w, h, p(one), stride) == len(want)
        assert one[:len(want)].tobytes() == want
        if ref is not None:
            assert oracle.ref_encode_tile(tiles[i, :1024], tiles[i, 1024:1024 + w * h], w, h, 0, 9).tobytes() == want
    # deflated variant: zlib level 9, one shot (tiledata.cpp:139-148)
    assert host.tcbh_queue_encode(0, 9, 1, p(tiles), p(wh), 20, p(out), stride * 2, p(sizes), 8, 0, 2) == 0
    for i in range(20):
        w, h = int(wh[i, 0]), int(wh[i, 1])
        want = bytes([w, 0, h, 0]) + tiles[i, :1024 + w * h].tobytes()
        assert out[i, :sizes[i]].tobytes() == zlib.compress(want, 9)


# ---- per-tile plugin surface --------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("type_,quality", ((1, 9), (2, 2), (3, 7)))
def test_converter_plugin_single_tiles(host, oracle, type_, quality):
    tiles, wh = synth.ragged_tiles(6, 91)
    for i in range(len(tiles)):
        w, h = int(wh[i, 0]), int(wh[i, 1])
        out = np.zeros(8192, np.uint8)
        pal, idx = tiles[i, :1024].copy(), tiles[i, 1024:].copy()      # keep the buffers alive across the call
        n = host.tcbh_convert_tile(type_, quality, p(pal), p(idx), w, h, p(out), 8192)
        want = oracle.encode_tile(tiles[i, :1024], tiles[i, 1024:1024 + w * h], w, h, type_, quality)
        assert n == len(want) == oracle.record_size(type_, w, h)
        assert np.array_equal(out[:n], want)
    # error behaviour: 0, never a crash (converter.h:83)
    out = np.zeros(8192, np.uint8)
    pal, idx = tiles[0, :1024].copy(), tiles[0, 1024:].copy()
    assert host.tcbh_convert_tile(type_, quality, p(pal), p(idx), 0, 5, p(out), 8192) == 0
    assert host.tcbh_convert_tile(7, quality, p(pal), p(idx), 4, 4, p(out), 8192) == 0


# ---- batch queue behind the TileThreadPool interface ----------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("slab,n", ((4096, 300), (64, 300), (7, 50)))
def test_batch_queue_ordered_results(host, oracle, slab, n):
    tiles = synth.mixed_tiles(n, 92)
    stride = oracle.record_size(1)
    out = np.zeros((n, stride), np.uint8); sizes = np.zeros(n, np.int32)
    rc = host.tcbh_queue_encode(1, 9, 0, p(tiles), None, n, p(out), stride, p(sizes), slab, 0, 2)
    assert rc == 0, host.tcbh_last_error()
    assert (sizes == stride).all()
    from tileconv_b200 import Encoder
    with Encoder(1, 9) as enc:
        assert np.array_equal(out, enc.encode(tiles))
    assert np.array_equal(out[:16], oracle.encode_tiles(tiles[:16], 1, 9, threads=8))


@pytest.mark.gpu
def test_batch_queue_tiles_with_different_options(host, oracle):
    """One pool, tiles that carry different Options (the reference's pool runs whatever functor each
    tile brings, tilethreadpool_posix.cpp:119-145): the slabs stay bound to the first tile's
    (type, level); the others go through the per-tile lane and still come back in index order."""
    n = 90
    tiles = synth.keyed_tiles(n, 97)
    stride = oracle.record_size(3)
    out = np.zeros((n, stride), np.uint8); sizes = np.zeros(n, np.int32)
    rc = host.tcbh_queue_encode_mixed(3, 2, 1, 5, 3, p(tiles), n, p(out), stride, p(sizes), 16, 2)
    assert rc == 0, host.tcbh_last_error()
    alt = np.arange(n) % 3 == 2
    assert (sizes[~alt] == stride).all() and (sizes[alt] == oracle.record_size(1)).all()
    assert np.array_equal(out[~alt], oracle.encode_tiles(tiles[~alt], 3, 2, threads=8))
    assert np.array_equal(out[alt][:, :oracle.record_size(1)], oracle.encode_tiles(tiles[alt], 1, 5, threads=8))


@pytest.mark.gpu
def test_batch_queue_shards_slabs_over_all_gpus(host, oracle):
    """Multi-GPU path of the queue: consecutive slabs go to consecutive devices (tile-range sharding at
    slab granularity, no collective); the ordered result stream must not depend on the device count."""
    from tileconv_b200 import capi
    gpus = capi.device_count()
    n = 700
    tiles = synth.mixed_tiles(n, 191)
    stride = oracle.record_size(3)
    outs = []
    # one context, one per visible device, and three contexts whatever the device count (on a one-GPU box they share
    # the device: same feeder / harvester / slot rotation code)
    for use in (1, gpus, 3):
        out = np.zeros((n, stride), np.uint8); sizes = np.zeros(n, np.int32)
        rc = host.tcbh_queue_encode(3, 9, 0, p(tiles), None, n, p(out), stride, p(sizes), 48, use, 2)
        assert rc == 0, host.tcbh_last_error()
        outs.append(out)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    assert np.array_equal(outs[1][:8], oracle.encode_tiles(tiles[:8], 3, 9, threads=8))
    # and the C ABI on every device gives the same records
    from tileconv_b200 import Encoder
    for d in range(gpus):
        with Encoder(3, 9, device=d) as enc:
            assert np.array_equal(enc.encode(tiles[:64]), outs[0][:64])


@pytest.mark.gpu
@pytest.mark.parametrize("slab,threads,deflate", ((5, 1, 0), (33, 4, 1), (257, 3, 0)))
def test_batch_queue_stress_many_small_slabs(host, oracle, slab, threads, deflate):
    """Feeder / harvester hand-over under pressure: thousands of tiles through tiny slabs (every slot is
    recycled hundreds of times), results must come back complete, in order and identical to one big batch."""
    n = 3000
    tiles = synth.rotate_palettes(synth.mixed_tiles(n, 192), 3)
    stride = 4096
    out = np.zeros((n, stride), np.uint8); sizes = np.zeros(n, np.int32)
    rc = host.tcbh_queue_encode(1, 2, deflate, p(tiles), None, n, p(out), stride, p(sizes), slab, 0, threads)
    assert rc == 0, host.tcbh_last_error()
    from tileconv_b200 import Encoder
    with Encoder(1, 2) as enc:
        want = enc.encode(tiles)
    for i in range(n):
        got = out[i, :sizes[i]].tobytes()
        if deflate:
            got = zlib.decompress(got)
        assert got == want[i].tobytes(), i


@pytest.mark.gpu
def test_batch_queue_ragged_and_deflated(host, oracle):
    tiles, wh = synth.ragged_tiles(40, 93)
    stride = 8192
    out = np.zeros((40, stride), np.uint8); sizes = np.zeros(40, np.int32)
    rc = host.tcbh_queue_encode(3, 5, 1, p(tiles), p(wh), 40, p(out), stride, p(sizes), 16, 0, 3)
    assert rc == 0, host.tcbh_last_error()
    for i in range(40):
        w, h = int(wh[i, 0]), int(wh[i, 1])
        raw = zlib.decompress(out[i, :sizes[i]].tobytes())
        assert raw == oracle.encode_tile(tiles[i, :1024], tiles[i, 1024:1024 + w * h], w, h, 3, 5).tobytes()
        # same zlib parameters as the reference (deflateInit level 9, one shot): identical bytes
        assert out[i, :sizes[i]].tobytes() == zlib.compress(raw, 9)


@pytest.mark.gpu
def test_batch_queue_reports_bad_tiles(host):
    tiles = synth.mixed_tiles(3, 94)
    wh = np.array([[64, 64], [0, 64], [64, 64]], np.uint16)
    out = np.zeros((3, 2052), np.uint8); sizes = np.zeros(3, np.int32)
    rc = host.tcbh_queue_encode(1, 2, 0, p(tiles), p(wh), 3, p(out), 2052, p(sizes), 2, 0, 1)
    assert rc == 1 and sizes.tolist() == [2052, 0, 2052]


# ---- container drivers ----------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("header", (True, False))
@pytest.mark.parametrize("deflate", (0, 1))
def test_tis_to_tbc_file(host, oracle, tmp_path, header, deflate):
    tiles = synth.mixed_tiles(40, 95)
    src, dst = str(tmp_path / "a.tis"), str(tmp_path / "a.tbc")
    write_tis(src, tiles, header)
    assert host.tcbh_encode_file(src.encode(), dst.encode(), 1, 9, deflate, 0 if header else 1, 2) == 0, host.tcbh_last_error()
    code = 1 | (0 if deflate else 0x100)
    want = expected_stream(oracle, b"TBC V1.0" + struct.pack("<II", code, 40), tiles, None, 1, 9, deflate)
    assert open(dst, "rb").read() == want


@pytest.mark.gpu
def test_ka8_framing_1024_tiles(host, tmp_path):
    """KA8: -T -u -t 1 on 1024 tiles -> 16 + 1024*(4+4+2048) bytes with the stated prefix."""
    tiles = synth.mixed_tiles(1024, 96)
    src, dst = str(tmp_path / "k.tis"), str(tmp_path / "k.tbc")
    write_tis(src, tiles, header=False)
    assert host.tcbh_encode_file(src.encode(), dst.encode(), 1, 2, 0, 1, 2) == 0
    blob = open(dst, "rb").read()
    assert len(blob) == 2105360
    assert blob[:16] == bytes.fromhex("54424320 56312E30 01010000 00040000".replace(" ", ""))
    assert blob[16:24] == bytes.fromhex("0408000040004000")
    # headerless input without -T is refused and leaves no output behind
    os.remove(dst)
    assert host.tcbh_encode_file(src.encode(), dst.encode(), 1, 2, 0, 0, 2) != 0
    assert not os.path.exists(dst)


@pytest.mark.gpu
@pytest.mark.parametrize("compressed", (False, True))
def test_mos_to_mbc_file_with_edge_tiles(host, oracle, tmp_path, compressed):
    mos, packed, wh = make_mos(150, 70, 97, compressed)        # 3x2 tiles, right/bottom edges partial
    src, dst = str(tmp_path / "m.mos"), str(tmp_path / "m.mbc")
    open(src, "wb").write(mos)
    assert host.tcbh_encode_file(src.encode(), dst.encode(), 2, 9, 0, 0, 2) == 0, host.tcbh_last_error()
    want = expected_stream(oracle, b"MBC V1.0" + struct.pack("<III", 2 | 0x100, 150, 70), packed, wh, 2, 9, 0)
    assert open(dst, "rb").read() == want


# ---- whole programs: the reference CLI vs the same CLI with the plugin dropped in -----------------
@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF_CLI) and os.path.exists(CUDA_CLI)), reason="oracle/_ref CLIs not built")
@pytest.mark.parametrize("args", (["-T", "-u", "-t", "1", "-q", "-9"], ["-T", "-t", "3", "-q", "-7"], ["-u", "-t", "2", "-q", "-2"],
                                  ["-T", "-t", "0"], ["-u", "-t", "0"]))
def test_reference_cli_with_plugin_writes_identical_files(tmp_path, args):
    if "-T" in args:
        src = str(tmp_path / "in.tis")
        write_tis(src, np.concatenate([synth.mixed_tiles(70, 98), synth.keyed_tiles(30, 99)]), header=False)
        ext = "tbc"
    else:
        src = str(tmp_path / "in.mos")
        open(src, "wb").write(make_mos(200, 130, 100)[0])
        ext = "mbc"
    outs = []
    for exe, name in ((REF_CLI, "ref"), (CUDA_CLI, "cuda")):
        dst = str(tmp_path / f"{name}.{ext}")
        res = subprocess.run([exe, "-s", "-j", "4", "-o", dst] + args + [src], capture_output=True, text=True, timeout=600)
        assert res.returncode == 0 and os.path.exists(dst), res.stdout + res.stderr
        outs.append(open(dst, "rb").read())
    assert outs[0] == outs[1]
    # and this repo's own driver agrees with both
    dst = str(tmp_path / f"own.{ext}")
    res = subprocess.run([TBC_ENCODE, "-s", "-o", dst] + args + [src], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    assert open(dst, "rb").read() == outs[0]


# ---- the streaming fast path (slab_pipeline) vs the TileData route --------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("type_,quality,deflate", ((1, 9, 0), (1, 2, 1), (3, 6, 0), (2, 4, 1)))
def test_slab_pipeline_and_tiledata_route_write_the_same_files(host, tmp_path, monkeypatch, type_, quality, deflate):
    """tbc_encode streams tile records through the pinned slabs; TCB200_TILEDATA_ROUTE=1 selects the reference-style
    caller loop (one TileData per tile through createThreadPool).  Same bytes, for a TIS with several slabs per
    device (small slabs forced), a one-tile TIS and a MOS with ragged edge tiles."""
    tiles = np.concatenate([synth.mixed_tiles(700, 191), synth.keyed_tiles(333, 192)])
    cases = []
    src = str(tmp_path / "big.tis"); write_tis(src, tiles, header=True); cases.append((src, 0))
    src = str(tmp_path / "one.tis"); write_tis(src, tiles[:1], header=False); cases.append((src, 1))
    src = str(tmp_path / "edge.mos"); open(src, "wb").write(make_mos(333, 197, 193)[0]); cases.append((src, 0))
    for src, assume in cases:
        outs = []
        for route, slab in (("0", "97"), ("1", "")):
            monkeypatch.setenv("TCB200_TILEDATA_ROUTE", route)
            if slab:
                monkeypatch.setenv("TCB200_SLAB_TILES", slab)
            else:
                monkeypatch.delenv("TCB200_SLAB_TILES", raising=False)
            dst = src + ".out" + route
            assert host.tcbh_encode_file(src.encode(), dst.encode(), type_, quality, deflate, assume, 3) == 0, host.tcbh_last_error()
            outs.append(open(dst, "rb").read())
        assert outs[0] == outs[1], src
    monkeypatch.delenv("TCB200_TILEDATA_ROUTE", raising=False)
    monkeypatch.delenv("TCB200_SLAB_TILES", raising=False)


@pytest.mark.gpu
def test_slab_pipeline_failures_leave_no_output(host, tmp_path):
    src, dst = str(tmp_path / "short.tis"), str(tmp_path / "short.tbc")
    with open(src, "wb") as f:
        f.write(b"TIS V1  " + struct.pack("<IIII", 10, 0x1400, 0x18, 0x40) + bytes(5120 * 3))     # claims 10 tiles, holds 3
    assert host.tcbh_encode_file(src.encode(), dst.encode(), 1, 2, 0, 0, 2) != 0
    assert b"Incomplete" in host.tcbh_last_error() and not os.path.exists(dst)
    # a MOS whose offset table points outside the file
    mos = bytearray(make_mos(100, 64, 194)[0])
    table = 24 + 2 * 1024
    mos[table + 4:table + 8] = struct.pack("<I", 1 << 30)
    src, dst = str(tmp_path / "bad.mos"), str(tmp_path / "bad.mbc")
    open(src, "wb").write(bytes(mos))
    assert host.tcbh_encode_file(src.encode(), dst.encode(), 1, 2, 0, 0, 2) != 0
    assert not os.path.exists(dst)


# ---- tileconv -I (N4): container facts printed exactly like TileConv::showInfo ---------------------------------
INFO_EXPECTED = """Parsing "{d}/h.tis"...
File type:       TIS
File subtype:    Paletted
Number of tiles: 3

Parsing "{d}/n.tis"...
File type:       TIS (assumed)
File subtype:    Paletted
Number of tiles: 3

Parsing "{d}/a.mos"...
File type:       MOS (compressed)
File subtype:    Paletted (V1)
Width:           150
Height:          70
Columns:         3
Rows:            2

Parsing "{d}/c.mos"...
File type:       MOS (compressed)
File subtype:    Paletted (V1)
Width:           150
Height:          70
Columns:         3
Rows:            2

Parsing "{d}/a.tbc"...
File type:       TBC
TBC version:     1.0
Compression:     0x0101 - BC1/DXT1 (uncompressed)
Number of tiles: 3

Parsing "{d}/a.mbc"...
File type:       MBC
MBC version:     1.0
Compression:     0x0002 - BC2/DXT3 (zlib-compressed)
Width:           150
Height:          70
Number of tiles: 6

Parsing "{d}/a.tiz"...
File type:       TIZ
Format version:  0
Number of tiles: 7

Parsing "{d}/a.moz"...
File type:       MOZ
Format version:  0
Width:           130
Height:          65
Number of tiles: 6

Parsing "{d}/x.bin"...
File type:       Unknown

Parsing "{d}/missing"...

"""


def test_info_dump_matches_the_reference(tmp_path):
    """`-I` needs no GPU.  The expected text was captured from the reference CLI (oracle/_ref/tileconv_ref -I -T ...); where
    that binary exists the two programs are also compared directly.  (The reference reports every MOS as "compressed":
    its isMosc flag is never cleared, tileconv.cpp:281-284.)"""
    d = str(tmp_path)
    tiles = synth.mixed_tiles(3, 1)
    write_tis(os.path.join(d, "h.tis"), tiles, header=True)
    write_tis(os.path.join(d, "n.tis"), tiles, header=False)
    open(os.path.join(d, "a.mos"), "wb").write(make_mos(150, 70, 5)[0])
    open(os.path.join(d, "c.mos"), "wb").write(make_mos(150, 70, 5, compressed=True)[0])
    open(os.path.join(d, "a.tbc"), "wb").write(b"TBC V1.0" + struct.pack("<II", 0x101, 3))
    open(os.path.join(d, "a.mbc"), "wb").write(b"MBC V1.0" + struct.pack("<III", 2, 150, 70))
    open(os.path.join(d, "a.tiz"), "wb").write(b"TIZ0" + struct.pack(">H", 7))
    open(os.path.join(d, "a.moz"), "wb").write(b"MOZ0" + struct.pack(">HH", 130, 65))
    open(os.path.join(d, "x.bin"), "wb").write(b"hello world")
    names = ["h.tis", "n.tis", "a.mos", "c.mos", "a.tbc", "a.mbc", "a.tiz", "a.moz", "x.bin", "missing"]
    cmd = ["-I", "-T"] + [os.path.join(d, n) for n in names]
    own = subprocess.run([TBC_ENCODE] + cmd, capture_output=True, text=True, timeout=60)
    assert own.returncode == 0
    assert own.stdout == INFO_EXPECTED.format(d=d)
    if os.path.exists(REF_CLI):
        ref = subprocess.run([REF_CLI] + cmd, capture_output=True, text=True, timeout=60)
        assert ref.stdout == own.stdout


# ---- container validation and the no-fallback rule need no GPU --------------------------------------------------------
def test_driver_rejects_malformed_containers_before_touching_the_gpu(host, tmp_path):
    """Header checks of readTIS / readMOS (tileconv/graphics.cpp:773-973): every malformed input is refused with the
    reference's message and leaves no output file."""
    tiles = synth.mixed_tiles(2, 7)
    body = tiles.tobytes()
    cases = [
        (b"TIS V9  " + struct.pack("<IIII", 2, 0x1400, 0x18, 0x40) + body, 0, b"Invalid TIS version"),
        (b"TIS V1  " + struct.pack("<IIII", 0, 0x1400, 0x18, 0x40) + body, 0, b"No tiles found"),
        (b"TIS V1  " + struct.pack("<IIII", 2, 0x000c, 0x18, 0x40) + body, 0, b"Invalid tile size"),
        (b"TIS V1  " + struct.pack("<IIII", 2, 0x1400, 0x10, 0x40) + body, 0, b"Invalid header size"),
        (b"TIS V1  " + struct.pack("<IIII", 2, 0x1400, 0x18, 0x20) + body, 0, b"Invalid tile dimensions"),
        (b"TIS V1  " + struct.pack("<IIII", 3, 0x1400, 0x18, 0x40) + body, 0, b"Incomplete TIS file"),
        (b"TIS V1  " + struct.pack("<II", 2, 0x1400), 0, b"Invalid TIS header"),
        (body[:-1], 1, b"Unsupported input file"),                   # headerless but not a multiple of 5120 bytes
        (body, 0, b"Unsupported input file"),                        # headerless without -T
        (b"MOS V2  " + bytes(40), 0, b"Unsupported MOS version"),
        (b"MOS V1  " + struct.pack("<HHHHII", 0, 64, 1, 1, 0x40, 24) + bytes(5120), 0, b"Invalid MOS dimensions"),
        (b"MOS V1  " + struct.pack("<HHHHII", 64, 64, 1, 1, 0x20, 24) + bytes(5120), 0, b"Invalid tile dimensions"),
        (b"MOS V1  " + struct.pack("<HHHHII", 64, 64, 1, 1, 0x40, 24) + bytes(100), 0, b"Incomplete or corrupted MOS file"),
        (b"MOSC" + b"V1  " + struct.pack("<I", 5000) + b"not a zlib stream", 0, b"Error while decompressing MOSC input file"),
    ]
    for k, (blob, assume, message) in enumerate(cases):
        src, dst = str(tmp_path / f"bad{k}.bin"), str(tmp_path / f"bad{k}.out")
        open(src, "wb").write(blob)
        assert host.tcbh_encode_file(src.encode(), dst.encode(), 1, 2, 0, assume, 2) != 0, k
        assert message in host.tcbh_last_error(), (k, host.tcbh_last_error())
        assert not os.path.exists(dst), k


def test_driver_has_no_cpu_fallback(host, tmp_path):
    from tileconv_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    src, dst = str(tmp_path / "ok.tis"), str(tmp_path / "ok.tbc")
    write_tis(src, synth.mixed_tiles(3, 8), header=True)
    for route in ("0", "1"):
        os.environ["TCB200_TILEDATA_ROUTE"] = route
        try:
            assert host.tcbh_encode_file(src.encode(), dst.encode(), 1, 9, 0, 0, 2) != 0
        finally:
            del os.environ["TCB200_TILEDATA_ROUTE"]
        assert not os.path.exists(dst)
