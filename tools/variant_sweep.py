"""Times alternative builds of libtcb200.so (tools/_bin/variants/lib_*.so, made by tools/build_variants.sh) kernel-only
on the bench pool (BC1 -q 9 mix) and on each synthetic class, and checks every output against the `base` build's
(sha of the records).  Development aid.   usage: python tools/variant_sweep.py [name-filter]"""
import glob, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, json
sys.path.insert(0, %r)
import numpy as np, torch, hashlib
import bench
from tileconv_b200 import Encoder, capi, synth
B = 16384
pool = bench.make_pool()
mix = np.zeros((2 * B, 5120), np.uint8); bench.fill_batch(mix, pool, 0)
sets = {"mix": mix, "S-U": np.concatenate([synth.uniform_tiles(512, 1)] * (B // 512)), "S-G": np.concatenate([synth.smooth_tiles(512, 2)] * (B // 512)),
        "S-K": np.concatenate([synth.keyed_tiles(512, 3)] * (B // 512))}
out = {}
for (t, q) in ((1, 9), (3, 9), (1, 2)):
    enc = Encoder(t, q)
    for name, tiles in sets.items():
        if (t == 3 and name == "mix") or (q == 2 and name not in ("mix", "S-K")): continue
        n = len(tiles)
        d_in = torch.from_numpy(tiles.reshape(-1)).cuda(); d_out = torch.zeros(n * enc.stride, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        for _ in range(2): enc.encode_device(d_in.data_ptr(), None, n, d_out.data_ptr())
        torch.cuda.synchronize()
        ms = min(enc.time_kernel_ms(d_in.data_ptr(), None, n, d_out.data_ptr(), 3) for _ in range(2))
        h = hashlib.sha1(d_out.cpu().numpy().tobytes()).hexdigest()[:10]
        out[f"bc{t}{'q2' if q == 2 else ''}_{name}"] = [n / (ms * 1e-3), h]
        del d_in, d_out
    enc.close()
print(json.dumps(out))
''' % ROOT
flt = sys.argv[1] if len(sys.argv) > 1 else ""
libs = sorted(glob.glob(os.path.join(ROOT, "tools", "_bin", "variants", "lib_*.so")))
libs = [l for l in libs if "base" in l] + [l for l in libs if "base" not in l and flt in l]
ref = None
for lib in libs:
    env = dict(os.environ, TCB200_LIB=lib)
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    try:
        r = json.loads(res.stdout.strip().splitlines()[-1])
    except Exception:
        print(os.path.basename(lib), "FAILED", res.stderr[-400:]); continue
    if ref is None: ref = r
    cells = []
    for k, (rate, h) in r.items():
        mark = "" if h == ref[k][1] else " MISMATCH"
        cells.append(f"{k} {rate/1e3:8.1f}k ({100*(rate/ref[k][0]-1):+5.1f}%){mark}")
    print(f"{os.path.basename(lib)[4:-3]:16s} " + " | ".join(cells), flush=True)
