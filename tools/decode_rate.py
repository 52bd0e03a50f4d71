"""Kernel-only throughput of the BCn decode kernel against the HBM roofline (records in, B,G,R,A pixels out)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tileconv_b200 import Encoder, synth
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
B = 32768
pool = synth.mixed_tiles(1024, 9)
for t in (1, 3):
    with Encoder(t, 2) as enc:
        rec = enc.encode(pool)
        recs = np.concatenate([rec] * (B // len(rec)))
        d_rec = torch.from_numpy(recs.reshape(-1)).cuda(); d_pix = torch.zeros(B * 4096 * 4, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        enc.time_decode_ms(d_rec.data_ptr(), B, d_pix.data_ptr(), 2)
        ms = enc.time_decode_ms(d_rec.data_ptr(), B, d_pix.data_ptr(), 10)
        got = d_pix.cpu().numpy().reshape(B, 4096, 4)[:8]
        assert np.array_equal(got, enc.decode(rec[:8]))
        bytes_tile = enc.stride + 16384
        gbs = B * bytes_tile / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": f"decode_bc{t}", "tiles_per_s": B / (ms * 1e-3), "ms": ms, "bytes_per_tile": bytes_tile,
                          "achieved_gbs": gbs, "hbm_peak_gbs": peak, "frac": gbs / peak}))
