"""Throughput of the reference-facing batch queue (TileBatchQueue behind createThreadPool's interface): tiles go
in as TileData objects (three heap buffers per tile, like graphics.cpp:101-103) and come back in index order."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
host = ctypes.CDLL(os.path.join(ROOT, "tileconv_b200", "libtcb200_host.so"))
def p(a): return ctypes.c_void_p(a.ctypes.data)
pool = bench.make_pool()
for (t, q, n) in ((1, 9, 65536), (1, 2, 262144)):
    tiles = np.zeros((n, 5120), np.uint8); bench.fill_batch(tiles, pool, 0)
    stride = 2052
    out = np.zeros((n, stride), np.uint8); sizes = np.zeros(n, np.int32)
    for slab in (4096, 16384):
        host.tcbh_queue_encode(t, q, 0, p(tiles), None, 4096, p(out), stride, p(sizes), slab, 1, 4)   # warm-up
        t0 = time.perf_counter()
        rc = host.tcbh_queue_encode(t, q, 0, p(tiles), None, n, p(out), stride, p(sizes), slab, 1, 4)
        dt = time.perf_counter() - t0
        print(f"BC{t} q{q} n={n} slab={slab}: rc={rc} {n / dt / 1e3:.1f} k tiles/s through TileBatchQueue", flush=True)
