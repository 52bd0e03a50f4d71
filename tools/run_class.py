"""Runs the encode kernel a few times on one synthetic tile class (for ncu): python tools/run_class.py S-G 1 9 [n]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tileconv_b200 import Encoder, synth
name, t, q = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 8192
gen = {"S-U": synth.uniform_tiles, "S-G": synth.smooth_tiles, "S-K": synth.keyed_tiles}[name]
pool = gen(512, 5)
tiles = np.concatenate([pool] * (B // len(pool)))
with Encoder(t, q) as enc:
    d_in = torch.from_numpy(tiles.reshape(-1)).cuda(); d_out = torch.zeros(B * enc.stride, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for _ in range(3):
        enc.encode_device(d_in.data_ptr(), None, B, d_out.data_ptr())
    ms = enc.time_kernel_ms(d_in.data_ptr(), None, B, d_out.data_ptr(), 2)
    print(name, t, q, f"{B / (ms * 1e-3) / 1e3:.1f} k tiles/s")
