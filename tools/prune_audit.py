"""Audit of the pruned cluster sweeps (bc_device.cuh run_sweep): an alternative build of libtcb200.so with
-DTCB_PRUNE_AUDIT=1 evaluates every candidate the bound test skips and counts
  * skipped candidates whose error was NOT above the incumbent (must be zero: such a candidate could have won),
  * skipped candidates whose binary32 error lies below the raw bound -(sum Q) -- the rounding the margin exists for --
    and the worst such shortfall as a fraction of the margin (must stay below 1; the smaller, the more headroom).
Build:  __graft_entry__.build() makes tileconv_b200/libtcb200_audit.so next to the product library.
Run:    python tools/prune_audit.py [tiles per class] [seed]         (on the GPU box; JSON lines on stdout)"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TCB200_LIB"] = os.path.join(ROOT, "tileconv_b200", "libtcb200_audit.so")
import numpy as np
from tileconv_b200 import Encoder, capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib = capi.load_library()
lib.tcb200_prune_audit.argtypes = [ctypes.POINTER(ctypes.c_double)]
sets = {"S-U": synth.uniform_tiles(n, seed), "S-G": synth.smooth_tiles(n, seed + 1), "S-K": synth.keyed_tiles(n, seed + 2),
        "S-C": synth.corner_case_tiles(n, seed + 3)}
out = (ctypes.c_double * 5)()
total_bad = 0
PERCEPTUAL = (0.2126, 0.7152, 0.0722)       # libsquish <= 1.10's default metric
runs = [("default", {}, t, q) for t in (1, 2, 3) for q in (4, 6, 9)]
runs += [("perceptual+sse", dict(metric=PERCEPTUAL, sse=True), t, q) for (t, q) in ((1, 9), (3, 4), (3, 9))]   # the VARIANT kernels
for variant, kw, t, q in runs:
    with Encoder(t, q, **kw) as enc:
        for name, tiles in sets.items():
            lib.tcb200_prune_audit(out)          # reset
            enc.encode(tiles)
            lib.tcb200_prune_audit(out)
            row = {"variant": variant, "type": t, "quality": q, "class": name, "tiles": n, "tested": int(out[0]), "skipped": int(out[1]),
                   "skipped_frac": round(out[1] / max(1.0, out[0]), 4), "wrongly_skipped": int(out[2]),
                   "below_raw_bound": int(out[3]), "worst_shortfall_over_margin": round(out[4], 5)}
            total_bad += int(out[2])
            print(json.dumps(row), flush=True)
print(json.dumps({"wrongly_skipped_total": total_bad}))
sys.exit(1 if total_bad else 0)
