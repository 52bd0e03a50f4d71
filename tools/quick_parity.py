"""Quick GPU-vs-oracle parity sweep (development aid; the real tests live in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tileconv_b200 import Encoder, synth
from oracle import pyoracle

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # 0 = the generators' default seeds
    if seed:
        sets = {"S-U": synth.uniform_tiles(n, seed), "S-G": synth.smooth_tiles(n, seed + 1), "S-K": synth.keyed_tiles(n, seed + 2),
                "S-C": synth.corner_case_tiles(n, seed + 3)}
        rag, wh = synth.ragged_tiles(n, seed + 4)
    else:
        sets = {"S-U": synth.uniform_tiles(n), "S-G": synth.smooth_tiles(n), "S-K": synth.keyed_tiles(n),
                "S-C": synth.corner_case_tiles(n)}
        rag, wh = synth.ragged_tiles(n)
    print("tiles per class", n, "seed", seed)
    bad = 0
    variants = {"default": {}, "perceptual+sse": dict(metric=pyoracle.PERCEPTUAL, sse=True)}
    for vname, kw in variants.items():
      pyoracle.set_variant(kw.get("metric"), kw.get("sse", False))
      print("== variant", vname)
      for t in (1, 2, 3):
        for q in ((0, 1, 2, 3, 5, 7, 8, 9) if vname == "default" else (2, 4, 9)):
            with Encoder(t, q, **kw) as enc:
                for name, tiles in sets.items():
                    t0 = time.time(); got = enc.encode(tiles); dt = time.time() - t0
                    want = pyoracle.encode_tiles(tiles, t, q, threads=16)
                    bs = 8 if t == 1 else 16
                    g = got[:, 4:].reshape(n, 256, bs); w_ = want[:, 4:].reshape(n, 256, bs)
                    diff = (g != w_).any(axis=2).sum()
                    hdr = (got[:, :4] != want[:, :4]).any()
                    print(f"type {t} q {q} {name}: blocks differing {diff}/{n*256} header_bad {hdr} gpu {dt*1e3:.1f} ms")
                    bad += int(diff) + int(hdr)
                got = enc.encode(rag, wh)
                want = pyoracle.encode_tiles_wh(rag, wh, t, q)
                nd = 0
                for i in range(n):
                    sz = pyoracle.record_size(t, int(wh[i, 0]), int(wh[i, 1]))
                    nd += int((got[i, :sz] != want[i, :sz]).any())
                print(f"type {t} q {q} ragged: tiles differing {nd}/{n}")
                bad += nd
    pyoracle.set_variant(None, False)
    print("TOTAL BAD", bad)
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
