"""Kernel-only rate per synthetic tile class (S-U, S-G, S-K) and mode; development aid."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tileconv_b200 import Encoder, synth
from oracle import pyoracle
B = 16384
sets = {"S-U": synth.uniform_tiles(512, 1), "S-G": synth.smooth_tiles(512, 2), "S-K": synth.keyed_tiles(256, 3)}
for (t, q) in ((1, 9), (1, 4), (3, 9), (1, 2)):
    with Encoder(t, q) as enc:
        for name, pool in sets.items():
            tiles = np.concatenate([pool] * (B // len(pool)))
            d_in = torch.from_numpy(tiles.reshape(-1)).cuda(); d_out = torch.zeros(B * enc.stride, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            enc.encode_device(d_in.data_ptr(), None, B, d_out.data_ptr()); torch.cuda.synchronize()
            ms = enc.time_kernel_ms(d_in.data_ptr(), None, B, d_out.data_ptr(), 3)
            pyoracle.totals_reset(); pyoracle.encode_tiles(pool[:64], t, q, threads=16); c = pyoracle.totals()
            cand = (c["cand3"] + c["cand4"]) / 64
            rate = B / (ms * 1e-3)
            # SMSP-cycles per 32 candidates at 1.965 GHz, 592 SMSPs
            cyc = (592 * 1.965e9 / rate) / (cand / 32) if cand else 0
            print(f"BC{t} q{q} {name}: {rate/1e3:9.1f} k tiles/s  cand/tile {cand:9.0f}  SMSP-cycles per 32 candidates {cyc:6.1f}  points/blk {c['points']/max(1,c['cluster_blocks']):.1f} sweeps/blk {(c['sweeps3']+c['sweeps4'])/max(1,c['cluster_blocks']):.2f}", flush=True)
