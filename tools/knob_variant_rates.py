"""Kernel rate of the VARIANT kernels (libsquish-version knobs: perceptual metric + SSE arithmetic) for two builds of the
library -- by default the product library against an exhaustive-sweep build (-DTCB_PRUNE=0) -- with the records of both
compared hash for hash.  Development aid.   usage: python tools/knob_variant_rates.py [libA.so libB.so]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, json, hashlib
sys.path.insert(0, %r)
import numpy as np, torch
import bench
from tileconv_b200 import Encoder, synth
B = 16384
pool = bench.make_pool()
mix = np.zeros((B, 5120), np.uint8); bench.fill_batch(mix, pool, 0)
sets = {"mix": mix, "S-K": np.concatenate([synth.keyed_tiles(512, 3)] * (B // 512))}
out = {}
for (t, q) in ((1, 9), (3, 9)):
    enc = Encoder(t, q, metric=(0.2126, 0.7152, 0.0722), sse=True)
    for name, tiles in sets.items():
        n = len(tiles)
        d_in = torch.from_numpy(tiles.reshape(-1)).cuda(); d_out = torch.zeros(n * enc.stride, dtype=torch.uint8, device="cuda")
        enc.encode_device(d_in.data_ptr(), None, n, d_out.data_ptr()); torch.cuda.synchronize()
        ms = min(enc.time_kernel_ms(d_in.data_ptr(), None, n, d_out.data_ptr(), 3) for _ in range(2))
        out[f"bc{t}_{name}"] = [n / (ms * 1e-3), hashlib.sha1(d_out.cpu().numpy().tobytes()).hexdigest()[:10]]
    enc.close()
print(json.dumps(out))
''' % ROOT
libs = sys.argv[1:3] if len(sys.argv) >= 3 else [os.path.join(ROOT, "tools", "_bin", "variants", "lib_noprune_full.so"), os.path.join(ROOT, "tileconv_b200", "libtcb200.so")]
ref = None
for lib in libs:
    res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, TCB200_LIB=lib), capture_output=True, text=True)
    try:
        r = json.loads(res.stdout.strip().splitlines()[-1])
    except Exception:
        print(os.path.basename(lib), "FAILED", res.stderr[-400:]); continue
    if ref is None: ref = r
    print(f"{os.path.basename(lib):24s} " + " | ".join(f"{k} {v[0]/1e3:8.1f}k ({100*(v[0]/ref[k][0]-1):+5.1f}%){'' if v[1] == ref[k][1] else ' MISMATCH'}" for k, v in r.items()), flush=True)
