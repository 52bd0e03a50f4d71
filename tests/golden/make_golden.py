"""Generates tests/golden/encode_golden.npz from the REFERENCE's own per-tile path (oracle/_ref:
the reference's ConverterDxt/TileData glue compiled from /root/reference, linked with the restated
squish).  Run in the container that has /root/reference:

    python tests/golden/make_golden.py

The fixture holds the inputs and, for every (type, quality) pair, the exact records the reference
path produced; tests compare both the oracle (CPU suite) and the CUDA path (GPU suite) with it.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle  # noqa: E402
from tileconv_b200 import synth  # noqa: E402

CONFIGS = [(t, q) for t in (1, 2, 3) for q in (0, 2, 3, 4, 5, 6, 7, 9)]
# the libsquish <= 1.10 default metric (squish_oracle_set_variant): host independent, unlike the SSE reciprocal
PERCEPTUAL_CONFIGS = [(1, 2), (1, 9), (2, 4), (3, 7)]


def main():
    pyoracle.build(ref=True)
    if pyoracle.reference() is None:
        raise SystemExit("oracle/_ref is not available: /root/reference is needed to make the goldens")
    tiles = np.concatenate([synth.uniform_tiles(2, 101), synth.smooth_tiles(2, 102), synth.keyed_tiles(2, 103)])
    # hand-made edge tiles: all index 0 with the key entry (fully transparent), a solid tile, a two-colour tile
    extra = np.zeros((3, 5120), np.uint8)
    extra[0, 0:4] = synth.KEY_ENTRY
    extra[1, 0:1024] = np.tile(np.array([10, 200, 30, 0], np.uint8), 256); extra[1, 1024:] = 7
    extra[2, 0:1024] = np.random.default_rng(5).integers(0, 256, 1024, dtype=np.uint8)
    extra[2, 1024:] = (np.indices((64, 64)).sum(axis=0) % 2).astype(np.uint8).reshape(-1) + 3
    tiles = np.concatenate([tiles, extra])
    rag, wh = synth.ragged_tiles(6, 104)
    out = {"tiles": tiles, "ragged_tiles": rag, "ragged_wh": wh}
    for t, q in CONFIGS:
        out[f"full_t{t}_q{q}"] = pyoracle.ref_encode_tiles(tiles, t, q, threads=4)
        recs = np.zeros((len(rag), pyoracle.record_size(t)), np.uint8)
        for i in range(len(rag)):
            w, h = int(wh[i, 0]), int(wh[i, 1])
            r = pyoracle.ref_encode_tile(rag[i, :1024], rag[i, 1024:1024 + w * h], w, h, t, q)
            recs[i, :len(r)] = r
        out[f"ragged_t{t}_q{q}"] = recs
    with pyoracle.variant(metric=pyoracle.PERCEPTUAL):
        for t, q in PERCEPTUAL_CONFIGS:
            out[f"perceptual_t{t}_q{q}"] = pyoracle.ref_encode_tiles(tiles, t, q, threads=4)
    path = os.path.join(ROOT, "tests", "golden", "encode_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
