"""Times the kernel-only BC1 -q9 rate of alternative builds of libtcb200.so (tools/_bin/variants/lib_*.so)
on the bench pool and checks each against the default build's output.  Development aid."""
import glob, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, json
sys.path.insert(0, %r)
import numpy as np, torch, hashlib
import bench
from tileconv_b200 import Encoder, capi
B = 32768
pool = bench.make_pool()
tiles = np.zeros((B, 5120), np.uint8); bench.fill_batch(tiles, pool, 0)
enc = Encoder(int(os.environ.get("SWEEP_TYPE", "1")), int(os.environ.get("SWEEP_Q", "9")))
d_in = torch.from_numpy(tiles.reshape(-1)).cuda(); d_out = torch.zeros(B * enc.stride, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for _ in range(2): enc.encode_device(d_in.data_ptr(), None, B, d_out.data_ptr())
torch.cuda.synchronize()
ms = min(enc.time_kernel_ms(d_in.data_ptr(), None, B, d_out.data_ptr(), 3) for _ in range(2))
h = hashlib.sha1(d_out.cpu().numpy().tobytes()).hexdigest()[:12]
print(json.dumps({"tiles_per_s": B / (ms * 1e-3), "ms": ms, "sha": h}))
''' % ROOT
libs = sorted(glob.glob(os.path.join(ROOT, "tools", "_bin", "variants", "lib_*.so")))
ref = None
for lib in libs:
    env = dict(os.environ, TCB200_LIB=lib)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    try:
        r = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception:
        print(os.path.basename(lib), "FAILED", out.stderr[-300:]); continue
    if "base" in lib: ref = r["sha"]
    print(f"{os.path.basename(lib):24s} {r['tiles_per_s']/1e3:9.1f} k tiles/s  {r['ms']:8.2f} ms  sha {r['sha']}", flush=True)
print("base sha", ref)
