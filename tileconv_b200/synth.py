"""Synthetic tile generators.  Every tile record is 5120 bytes: a 1024-byte palette of
256 x {B,G,R,0} followed by 4096 palette indices -- the in-memory layout the reference reads from a
headerless TIS (tileconv/graphics.cpp:101-119, :832-843).

  S-U "uniform"  palette entries and indices uniformly random: ~16 distinct colours per 4x4 block,
                 no transparency.  Worst case for the cluster fitters.
  S-G "smooth"   a random colour plane (2-D gradient) + N(0,4) noise, quantised to a 256-entry
                 palette built from the tile itself (equal-population bins along luminance):
                 blocks with few colours, the occasional single-colour block.
  S-K "keyed"    S-G with palette[0] := 00 FF 00 00 (the reference's transparency key,
                 tileconv/colors.cpp:47) and ~25 % of the pixels set to index 0 in random discs:
                 BC1 3-colour+transparent blocks, BC2/BC3 alpha, weight-by-alpha.

All generators are deterministic functions of (n, seed).
"""
from __future__ import annotations

import numpy as np

TILE_RECORD = 5120
KEY_ENTRY = np.array([0x00, 0xFF, 0x00, 0x00], dtype=np.uint8)


def uniform_tiles(n: int, seed: int = 1234) -> np.ndarray:
    rng = np.random.default_rng(seed)
    tiles = np.zeros((n, TILE_RECORD), dtype=np.uint8)
    pal = rng.integers(0, 256, size=(n, 256, 4), dtype=np.uint8)
    pal[:, :, 3] = 0
    # keep index 0 opaque: never emit the key entry by accident
    clash = (pal[:, 0, 0] == 0) & (pal[:, 0, 1] == 255) & (pal[:, 0, 2] == 0)
    pal[clash, 0, 1] = 254
    tiles[:, :1024] = pal.reshape(n, 1024)
    tiles[:, 1024:] = rng.integers(0, 256, size=(n, 4096), dtype=np.uint8)
    return tiles


def smooth_tiles(n: int, seed: int = 2345, chunk: int = 256) -> np.ndarray:
    rng = np.random.default_rng(seed)
    tiles = np.zeros((n, TILE_RECORD), dtype=np.uint8)
    yy, xx = np.mgrid[0:64, 0:64]
    xx = (xx.reshape(1, 4096, 1) / 63.0).astype(np.float32)
    yy = (yy.reshape(1, 4096, 1) / 63.0).astype(np.float32)
    lum = np.array([1.0, 2.0, 1.0], dtype=np.float32)
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        corners = rng.uniform(0.0, 255.0, size=(m, 3, 3)).astype(np.float32)      # c00, c10, c01 (r,g,b)
        img = (corners[:, 0:1, :] + (corners[:, 1:2, :] - corners[:, 0:1, :]) * xx
               + (corners[:, 2:3, :] - corners[:, 0:1, :]) * yy
               + rng.normal(0.0, 4.0, size=(m, 4096, 3)).astype(np.float32))
        img = np.clip(np.rint(img), 0, 255)
        key = img @ lum
        order = np.argsort(key, axis=1, kind="stable")                             # (m, 4096)
        rank = np.empty_like(order)
        np.put_along_axis(rank, order, np.broadcast_to(np.arange(4096), (m, 4096)).copy(), axis=1)
        index = (rank // 16).astype(np.uint8)                                      # 256 bins x 16 pixels
        rep = order[:, 8::16]                                                      # bin representative
        pal_rgb = np.take_along_axis(img, rep[:, :, None], axis=1).astype(np.uint8)  # (m,256,3) r,g,b
        pal = np.zeros((m, 256, 4), dtype=np.uint8)
        pal[:, :, 0] = pal_rgb[:, :, 2]
        pal[:, :, 1] = pal_rgb[:, :, 1]
        pal[:, :, 2] = pal_rgb[:, :, 0]
        clash = (pal[:, 0, 0] == 0) & (pal[:, 0, 1] == 255) & (pal[:, 0, 2] == 0)
        pal[clash, 0, 1] = 254
        tiles[lo:lo + m, :1024] = pal.reshape(m, 1024)
        tiles[lo:lo + m, 1024:] = index
    return tiles


def keyed_tiles(n: int, seed: int = 3456) -> np.ndarray:
    tiles = smooth_tiles(n, seed)
    rng = np.random.default_rng(seed + 1)
    yy, xx = np.mgrid[0:64, 0:64]
    for t in range(n):
        tiles[t, 0:4] = KEY_ENTRY
        mask = np.zeros((64, 64), dtype=bool)
        while mask.mean() < 0.25:
            cx, cy = rng.uniform(0, 64, size=2)
            r = rng.uniform(3, 14)
            mask |= (xx - cx) ** 2 + (yy - cy) ** 2 <= r * r
        idx = tiles[t, 1024:].reshape(64, 64)
        idx[mask] = 0
    return tiles


def mixed_tiles(n: int, seed: int = 777) -> np.ndarray:
    """Half S-U, half S-G, interleaved (the BASELINE config-5 pool)."""
    a = uniform_tiles((n + 1) // 2, seed)
    b = smooth_tiles(n // 2, seed + 1)
    tiles = np.zeros((n, TILE_RECORD), dtype=np.uint8)
    tiles[0::2] = a
    tiles[1::2] = b
    return tiles


def rotate_palettes(tiles: np.ndarray, shift: int) -> np.ndarray:
    """Byte-different copies with identical pixels: palette rolled by `shift` entries, indices
    shifted to match.  Not for keyed tiles (index 0 is special there)."""
    n = tiles.shape[0]
    out = np.empty_like(tiles)
    pal = tiles[:, :1024].reshape(n, 256, 4)
    out[:, :1024] = np.roll(pal, shift, axis=1).reshape(n, 1024)
    out[:, 1024:] = (tiles[:, 1024:].astype(np.int32) + shift).astype(np.uint8)
    return out


def ragged_tiles(n: int, seed: int = 4567):
    """MOS-style edge tiles: random sizes 1..64 with index rows packed at pitch = width
    (tileconv/graphics.cpp:322-342).  Returns (tiles, wh)."""
    rng = np.random.default_rng(seed)
    base = keyed_tiles(n, seed)
    wh = rng.integers(1, 65, size=(n, 2)).astype(np.uint16)
    wh[0] = (64, 64); wh[1 % n] = (1, 1); wh[2 % n] = (63, 5); wh[3 % n] = (4, 64)
    tiles = np.zeros_like(base)
    tiles[:, :1024] = base[:, :1024]
    for t in range(n):
        w, h = int(wh[t, 0]), int(wh[t, 1])
        img = base[t, 1024:].reshape(64, 64)[:h, :w]
        tiles[t, 1024:1024 + w * h] = img.reshape(-1)
    return tiles, wh


def corner_case_tiles(n: int, seed: int = 5678) -> np.ndarray:
    """Blocks built to stress the fitters' numerics rather than their throughput: per 4x4 block one of
    two colours only / colours on a line in RGB / colours already on the 5:6:5 grid / colours one LSB
    apart / a single colour with one outlier / 16 greys / duplicated palette entries.  No transparency."""
    rng = np.random.default_rng(seed)
    tiles = np.zeros((n, TILE_RECORD), dtype=np.uint8)
    for t in range(n):
        pal = np.zeros((256, 3), dtype=np.int64)      # r,g,b
        idx = np.zeros((64, 64), dtype=np.uint8)
        used = 0
        for by in range(16):
            for bx in range(16):
                kind = int(rng.integers(0, 7))
                k = int(rng.integers(2, 9))
                if kind == 0:      # two colours
                    cols = rng.integers(0, 256, (2, 3))
                elif kind == 1:    # collinear
                    a, d = rng.integers(0, 200, 3), rng.integers(-6, 7, 3)
                    cols = np.clip(a[None, :] + np.arange(k)[:, None] * d[None, :], 0, 255)
                elif kind == 2:    # on the 565 grid
                    r5, g6, b5 = rng.integers(0, 32, k), rng.integers(0, 64, k), rng.integers(0, 32, k)
                    cols = np.stack([(r5 << 3) | (r5 >> 2), (g6 << 2) | (g6 >> 4), (b5 << 3) | (b5 >> 2)], axis=1)
                elif kind == 3:    # one LSB apart
                    a = rng.integers(1, 255, 3)
                    cols = np.clip(a[None, :] + rng.integers(-1, 2, (k, 3)), 0, 255)
                elif kind == 4:    # one colour + outlier
                    cols = np.stack([rng.integers(0, 256, 3), rng.integers(0, 256, 3)])
                elif kind == 5:    # greys
                    g = np.sort(rng.integers(0, 256, k))
                    cols = np.stack([g, g, g], axis=1)
                else:              # duplicated colours under different palette indices
                    a = rng.integers(0, 256, (max(1, k // 2), 3))
                    cols = np.concatenate([a, a])
                m = len(cols)
                if used + m > 256:
                    used = 0       # start overwriting: earlier blocks keep their indices but may change colour (fine)
                pal[used:used + m] = cols
                choice = rng.integers(0, m, 16)
                if kind == 4:
                    choice[:] = 0; choice[int(rng.integers(0, 16))] = 1
                idx[4 * by:4 * by + 4, 4 * bx:4 * bx + 4] = (used + choice).reshape(4, 4)
                used += m
        tiles[t, 0:1024:4] = pal[:, 2]; tiles[t, 1:1024:4] = pal[:, 1]; tiles[t, 2:1024:4] = pal[:, 0]
        if tiles[t, 0] == 0 and tiles[t, 1] == 255 and tiles[t, 2] == 0:
            tiles[t, 1] = 254
        tiles[t, 1024:] = idx.reshape(-1)
    return tiles
