"""Wall-clock of the real tileconv CLI (reference sources) as shipped vs with this repo's plugin dropped in,
on a synthetic headerless TIS.  Needs oracle/_ref/tileconv_ref and tileconv_cuda (built where /root/reference exists)."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
pool = bench.make_pool()
with tempfile.TemporaryDirectory() as tmp:
    for exe, n, args in (("tileconv_ref", 2048, ["-u", "-t", "1", "-q", "-9"]), ("tileconv_cuda", 65536, ["-u", "-t", "1", "-q", "-9"]),
                         ("tileconv_cuda", 65536, ["-t", "1", "-q", "-9"]), ("tileconv_cuda", 262144, ["-u", "-t", "1", "-q", "-2"]),
                         ("tileconv_ref", 16384, ["-u", "-t", "1", "-q", "-2"])):
        tiles = np.zeros((n, 5120), np.uint8); bench.fill_batch(tiles, pool, 0)
        src, dst = os.path.join(tmp, "in.tis"), os.path.join(tmp, "out.tbc")
        tiles.tofile(src)
        path = os.path.join(ROOT, "oracle", "_ref", exe)
        subprocess.run([path, "-s", "-T", "-j", "0", "-o", dst] + args + [src], capture_output=True)      # warm-up (page cache, CUDA init)
        t0 = time.perf_counter()
        res = subprocess.run([path, "-s", "-T", "-j", "0", "-o", dst] + args + [src], capture_output=True, text=True)
        dt = time.perf_counter() - t0
        print(f"{exe:14s} {' '.join(args):22s} n={n:7d}: rc={res.returncode} {dt:7.2f} s  {n / dt / 1e3:8.1f} k tiles/s  out {os.path.getsize(dst) / 1e6:.1f} MB", flush=True)
