"""How far apart are the plausible libsquish builds?  

For each tile class (S-U, S-G, S-K), pixel type and -q level: the share of 4x4 blocks whose encoding
differs from the default build's (libsquish 1.15 scalar, metric (1,1,1)) and the per-tile RGB RMSE
(8-bit units, decoded through the GPU BCn decoder, opaque source pixels only) of each variant:
  sse             SQUISH_USE_SSE build: rcpps + Newton reciprocal, minps/maxps, cvttps (host CPU's rcpps)
  perceptual      libsquish <= 1.10 default metric (0.2126, 0.7152, 0.0722)
  perceptual+sse  both
Runs on the CUDA path (the GPU parity tests pin each variant to the oracle configured the same way).
usage: python tools/variant_sensitivity.py [tiles_per_class] > profiles/r2_variant_sensitivity.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from tileconv_b200 import Encoder, capi, synth  # noqa: E402

VARIANTS = {"default": {}, "sse": dict(sse=True), "perceptual": dict(metric=capi.PERCEPTUAL),
            "perceptual+sse": dict(metric=capi.PERCEPTUAL, sse=True)}


def rmse_per_tile(enc, tiles, records):
    src = enc.pal_to_argb(tiles).astype(np.float64)            # (n, 4096, 4) B,G,R,A
    dec = enc.decode(records).astype(np.float64)
    opaque = src[:, :, 3] > 0
    d2 = ((src[:, :, :3] - dec[:, :, :3]) ** 2).sum(axis=2) * opaque
    cnt = np.maximum(opaque.sum(axis=1), 1)
    return np.sqrt(d2.sum(axis=1) / (3.0 * cnt))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    sets = {"S-U": synth.uniform_tiles(n, 9001), "S-G": synth.smooth_tiles(n, 9002), "S-K": synth.keyed_tiles(n, 9003)}
    rows = []
    for type_ in (1, 3):
        bs = 8 if type_ == 1 else 16
        for quality in (2, 4, 9):
            enc = {name: Encoder(type_, quality, **kw) for name, kw in VARIANTS.items()}
            for cls, tiles in sets.items():
                recs = {name: e.encode(tiles) for name, e in enc.items()}
                rm = {name: rmse_per_tile(enc["default"], tiles, r) for name, r in recs.items()}
                base = recs["default"][:, 4:].reshape(n, 256, bs)
                for name in ("sse", "perceptual", "perceptual+sse"):
                    diff = (recs[name][:, 4:].reshape(n, 256, bs) != base).any(axis=2)
                    delta = rm[name] - rm["default"]
                    rows.append({"type": type_, "quality": quality, "class": cls, "variant": name, "tiles": n,
                                 "blocks_differing_pct": 100.0 * float(diff.mean()),
                                 "tiles_differing_pct": 100.0 * float(diff.any(axis=1).mean()),
                                 "rmse_default_mean": float(rm["default"].mean()), "rmse_variant_mean": float(rm[name].mean()),
                                 "rmse_delta_mean": float(delta.mean()), "rmse_delta_max_abs": float(np.abs(delta).max())})
            for e in enc.values():
                e.close()
    print(json.dumps({"tiles_per_class": n, "rmse": "per-tile RGB RMSE, 8-bit units, plain (unweighted) RGB, opaque source pixels",
                      "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
