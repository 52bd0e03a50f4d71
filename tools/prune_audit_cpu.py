"""CPU counterpart of tools/prune_audit.py: the oracle restates the CUDA path's bound test in binary32 and checks it against
the error of EVERY 4-colour candidate, with the sequential best-so-far as incumbent (tighter than anything the kernels use).
usage: python tools/prune_audit_cpu.py [tiles per class] [seed] [default|perceptual|sse|perceptual+sse]   -> JSON lines"""
import json, os, sys
from concurrent.futures import ProcessPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def work(job):
    name, t, q, n, seed, variant = job
    from oracle import pyoracle
    from tileconv_b200 import synth
    gen = {"S-U": synth.uniform_tiles, "S-G": synth.smooth_tiles, "S-K": synth.keyed_tiles, "S-C": synth.corner_case_tiles}[name]
    pyoracle.set_variant(pyoracle.PERCEPTUAL if "perceptual" in variant else None, "sse" in variant)
    a = pyoracle.bound_audit(gen(n, seed), t, q)
    a["variant"] = variant
    a.update({"type": t, "quality": q, "class": name, "tiles": n, "skipped_frac": round(a["skipped"] / max(1, a["tested"]), 4)})
    return a


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    variant = sys.argv[3] if len(sys.argv) > 3 else "default"
    jobs = [(c, t, q, n, seed + i, variant) for i, c in enumerate(("S-U", "S-G", "S-K", "S-C")) for (t, q) in ((1, 4), (1, 9), (2, 6), (3, 4), (3, 9))]
    bad = 0
    with ProcessPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for row in ex.map(work, jobs):
            bad += row["wrongly_skipped"]
            print(json.dumps(row), flush=True)
    print(json.dumps({"wrongly_skipped_total": bad}))
    sys.exit(1 if bad else 0)
