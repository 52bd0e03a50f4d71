"""Host-to-host rate of tcb200_encode() against the chunk size (max_batch_tiles) for BC1/BC3 -q 2 and BC1 -q 9: how the
H2D / kernel / D2H pipeline of one context overlaps.  Development aid (GPU).  usage: python tools/chunk_sweep.py"""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from tileconv_b200 import Encoder, capi
B = 65536
pool = bench.make_pool()
h_in = capi.PinnedBuffer(B * 5120); tiles = h_in.array.reshape(B, 5120); bench.fill_batch(tiles, pool, 0)
for t, q in ((1, 2), (3, 2), (1, 9)):
    for mb in (1024, 2048, 4096, 8192, 16384):
        with Encoder(t, q, max_batch_tiles=mb) as enc:
            h_out = capi.PinnedBuffer(B * enc.stride)
            for _ in range(2): enc.encode_ptr(h_in.ptr, None, B, h_out.ptr)
            t0 = time.perf_counter()
            reps = 5 if q == 2 else 2
            for _ in range(reps): enc.encode_ptr(h_in.ptr, None, B, h_out.ptr)
            dt = (time.perf_counter() - t0) / reps
            print(f"BC{t} q{q} chunk {mb:6d}: {B / dt / 1e6:8.3f} M tiles/s", flush=True)
            h_out.free()
