#!/usr/bin/env python
"""bench.py -- throughput of the TIS -> TBC encode hot path on B200, one JSON line on stdout.

Workload (config.workload): BASELINE.json configs[4] -- the BC1 `-q 9` throughput sweep on synthetic
64x64 tiles (half S-U, half S-G, tileconv_b200/synth.py), every tile byte-distinct (palette
rotation).  One step = one pass of the hot path over one batch of `--batch` tiles per GPU; with N
GPUs the tile-index range is sharded N ways with no collective (weak scaling: each rank encodes its
own `--batch` tiles per step).

  value        tiles/s, kernel only, tiles resident in HBM, CUDA events on the launching stream,
               max over ranks.  Inputs (batch x 5120 B) are larger than the 126 MB L2.
  e2e          tiles/s through the C ABI tcb200_encode(): pinned HOST buffers in, pinned HOST buffers
               out, H2D + kernels + D2H inside the timed region.
  roofline     FP32-ALU roofline of the dominant kernel (the path is ALU-bound, not a contraction):
               algorithmic flops = candidates the ORACLE evaluated on the same tiles x the per-
               candidate op counts of SURVEY.md s8d (160 / 138) + the fixed per-block term;
               peak = FFMA-loop rate measured live on this GPU.  `hbm` sub-object: algorithmic
               bytes (7172 B/tile BC1) over the same kernel time against MEASURED_PEAKS.json.
  cpu_baseline the reference's own per-tile CPU path (oracle/_ref: reference glue + restated squish)
               on all host cores, bounded sample, rank 0 only.

  modes        the metric's other modes (BC1 -q 2, BC3 -q 9, BC3 -q 2) on the same tiles: kernel-only AND
               host-to-host through tcb200_encode(), each with its own roofline (ALU for the cluster levels,
               measured pinned-copy bandwidth for the RangeFit levels, which are PCIe-bound end to end).
  classes      kernel-only rate and ALU-roofline fraction per synthetic tile class (S-U, S-G, S-K).
  e2e_in_process (N > 1) rank 0 re-runs the whole N-GPU job in ONE process through tcb200_encode_multi()
               after the timed region and checks every shard against the ranks' own results.

`--impl reference` times that CPU path alone (the reference arm).
`--scaling strong --total T`: one T-tile job (BASELINE configs[4] literally: 1 M tiles) split N ways.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

F4_OPS = 160       # FP32 operations per 4-colour candidate
F3_OPS = 138       # per 3-colour candidate
POOL_TILES = 4096  # distinct synthetic tiles; a batch is palette-rotated copies of the pool


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--type", type=int, default=1)
    ap.add_argument("--quality", type=int, default=9)
    ap.add_argument("--batch", type=int, default=65536, help="tiles per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=2048, help="tiles in the CPU baseline sample")
    ap.add_argument("--no-modes", action="store_true", help="skip the BC3/-q2 side measurements")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch tiles per GPU per step; strong: --total tiles per step split over the GPUs")
    ap.add_argument("--total", type=int, default=1 << 20, help="tiles of the whole job per step with --scaling strong")
    return ap.parse_args()


def make_pool():
    from tileconv_b200 import synth
    return synth.mixed_tiles(POOL_TILES, seed=777)


def fill_batch(dst: np.ndarray, pool: np.ndarray, first_global_tile: int) -> None:
    """dst[(k*POOL + i)] = pool[i] with its palette rotated by a shift derived from the global copy
    number, so no two tiles of the whole job are byte-identical while every copy encodes to the
    same blocks as its pool tile."""
    from tileconv_b200 import synth
    n = dst.shape[0]
    p = pool.shape[0]
    for lo in range(0, n, p):
        m = min(p, n - lo)
        copy = (first_global_tile + lo) // p
        shift = copy % 256
        dst[lo:lo + m] = synth.rotate_palettes(pool[:m], shift) if shift else pool[:m]


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for p in self.samples:
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # samples under load: drop the idle ones (clock well below the rest) if any load samples exist
        load = [c for c in sm if c > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_reference_rate(pool: np.ndarray, sample: int, type_: int, quality: int, repeats: int = 3, simd: bool = False):
    """Times the reference's CPU path on all host cores over `sample` tiles, best of `repeats` passes (the first
    one warms the caches and the thread pool up).  Returns (tiles_per_s, cores, kind,
    counters_per_tile).  simd = False is the configuration the parity definition names (libsquish's scalar simd_float.h
    path, SQUISH_USE_SSE = 0, the default of its Makefile build); simd = True evaluates the cluster solve on 4-wide SSE2
    vectors the way a SQUISH_USE_SSE build lays out its Vec4 (same results here, ~2.3x faster): reported beside it."""
    from oracle import pyoracle
    pyoracle.set_simd(simd)
    try:
        return _cpu_reference_rate(pyoracle, pool, sample, type_, quality, repeats)
    finally:
        pyoracle.set_simd(False)


def _cpu_reference_rate(pyoracle, pool: np.ndarray, sample: int, type_: int, quality: int, repeats: int):
    cores = os.cpu_count() or 1
    tiles = pool[:sample]
    best = None
    counters = None
    use_ref = pyoracle.reference() is not None
    for _ in range(repeats):
        if use_ref:
            pyoracle.ref_totals_reset()
            t0 = time.perf_counter()
            pyoracle.ref_encode_tiles(tiles, type_, quality, cores)
            dt = time.perf_counter() - t0
            c = pyoracle.ref_totals()
        else:
            pyoracle.totals_reset()
            t0 = time.perf_counter()
            pyoracle.encode_tiles(tiles, type_, quality, cores)
            dt = time.perf_counter() - t0
            c = pyoracle.totals()
        if best is None or dt < best:
            best, counters = dt, c
    per_tile = {k: v / len(tiles) for k, v in counters.items()}
    return len(tiles) / best, cores, ("reference" if use_ref else "port"), per_tile


def cpu_tuned_port_rate(pool: np.ndarray, sample: int, type_: int, quality: int, repeats: int = 3):
    """The restated squish + tile loop built with -O3 -mavx2 (oracle/libtco_avx2.so; identical results), all cores, best of
    `repeats`: what a deployment that tunes its CPU build would get.  None where that build or AVX2 is missing."""
    from oracle import pyoracle
    if pyoracle.tuned() is None:
        return None
    cores = os.cpu_count() or 1
    tiles = pool[:sample]
    best = None
    pyoracle.set_simd(True)
    try:
        for _ in range(repeats):
            t0 = time.perf_counter()
            pyoracle.tuned_encode_tiles(tiles, type_, quality, cores)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
    finally:
        pyoracle.set_simd(False)
    return {"value": len(tiles) / best, "unit": "tiles/s", "cores": cores, "kind": "port",
            "flags": "-O3 -mavx2 -ffp-contract=off (the reference's Makefile builds with -O2 -msse -mfpmath=sse), SSE2-vector cluster solve",
            "note": "same restated squish and tile loop, bit-identical output: the fastest CPU build of this algorithm made here"}


def as_shipped_cli_rate(pool: np.ndarray, sample: int, type_: int, quality: int):
    """BASELINE.md baseline A: the reference CLI as shipped (its own polling thread pool), wall clock of
    `tileconv -s -T -u -t T -q -Q -j 0` on a headerless TIS of `sample` tiles.  None when the CLI was not
    built (oracle/_ref/tileconv_ref needs /root/reference at build time)."""
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "tileconv_ref")
    if not os.path.exists(exe):
        return None
    with tempfile.TemporaryDirectory() as tmp:
        src, dst = os.path.join(tmp, "in.tis"), os.path.join(tmp, "out.tbc")
        pool[:sample].tofile(src)
        t0 = time.perf_counter()
        res = subprocess.run([exe, "-s", "-T", "-u", "-t", str(type_), "-q", f"-{quality}", "-j", "0", "-o", dst, src],
                             capture_output=True, text=True)
        dt = time.perf_counter() - t0
        if res.returncode != 0 or not os.path.exists(dst):
            return None
    return {"value": sample / dt, "unit": "tiles/s", "sample": f"{sample} tiles through the reference CLI (-j 0, -u), wall clock incl. file I/O",
            "note": "capped by the reference pool's 64-tile window and 50 ms polling sleeps (SURVEY.md F4)"}


def pool_seam_rate(tiles: np.ndarray, type_: int, quality: int):
    """Side measurement: the same tiles through the reference-facing pool seam -- TileBatchQueue behind
    createThreadPool()'s interface, fed TileData objects the way Graphics::tisToTBC does (three heap buffers
    per tile, results drained in index order).  The cost of that caller loop is part of the number."""
    import ctypes
    path = os.path.join(ROOT, "tileconv_b200", "libtcb200_host.so")
    if not os.path.exists(path):
        return None
    host = ctypes.CDLL(path)
    n = tiles.shape[0]
    stride = 2052 if type_ == 1 else 4100
    out = np.zeros((n, stride), np.uint8)
    sizes = np.zeros(n, np.int32)
    args = lambda m: (type_, quality, 0, ctypes.c_void_p(tiles.ctypes.data), None, m, ctypes.c_void_p(out.ctypes.data), stride,
                      ctypes.c_void_p(sizes.ctypes.data), 4096, 1, 4)
    host.tcbh_queue_encode(*args(min(n, 8192)))          # warm-up (contexts, pinned slabs, allocator)
    t0 = time.perf_counter()
    rc = host.tcbh_queue_encode(*args(n))
    dt = time.perf_counter() - t0
    if rc != 0:
        return None
    return {"value": n / dt, "unit": "tiles/s", "tiles": n,
            "path": "TileData objects -> TileBatchQueue (createThreadPool seam) -> ordered results, 1 GPU, -u",
            "note": "bounded by the single-threaded reference caller loop (per-tile heap buffers + ordered drain), not by the GPU"}


def pruning_facts(type_: int, quality: int):
    """Fraction of the 4-colour candidates the kernels skipped by their exact lower bound, per tile class, from the committed
    audit run (profiles/r2_prune_audit.jsonl: tools/prune_audit.py on the audit build) -- not measured in this run."""
    if quality <= 2:
        return None                                                # RangeFit levels: no cluster sweeps
    group = 9 if quality >= 7 else (6 if quality >= 5 else 4)      # the audit runs one level per quality class
    out = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r2_prune_audit.jsonl")) as f:
            for line in f:
                r = json.loads(line)
                if r.get("variant", "default") == "default" and r.get("type") == type_ and r.get("quality") == group and r.get("class") in ("S-U", "S-G", "S-K"):
                    out[r["class"]] = {"skipped_frac": r["skipped_frac"], "wrongly_skipped": r["wrongly_skipped"]}
    except Exception:
        return None
    if not out:
        return None
    out["source"] = "profiles/r2_prune_audit.jsonl (committed audit run, 256 tiles per class)"
    return out


def flops_per_tile(per_tile: dict) -> float:
    """W = C3*138 + C4*160 + F_fixed, F_fixed(n) ~ 30 n + 8*27 per cluster block."""
    return (per_tile["cand3"] * F3_OPS + per_tile["cand4"] * F4_OPS
            + 30.0 * per_tile["points"] + 8 * 27 * per_tile["cluster_blocks"])


def workload_name(type_: int, quality: int, batch: int) -> str:
    return (f"BASELINE configs[4]: synthetic 64x64 tiles (50% S-U uniform, 50% S-G smooth, palette-rotated copies of a "
            f"{POOL_TILES}-tile pool) -> TBC records, BC{type_} -q {quality}, {batch} tiles per GPU per step, tile-range sharding")


def run_reference(args, rank: int, world: int) -> int:
    if rank != 0:
        return 0
    pool = make_pool()
    sample = min(args.cpu_sample, POOL_TILES)
    rates = []
    kind, cores = "port", os.cpu_count() or 1
    for step in range(args.warmup + args.steps):
        rate, cores, kind, _ = cpu_reference_rate(pool, sample, args.type, args.quality, repeats=1)
        if step >= args.warmup:
            rates.append(rate)
    value = len(rates) * sample / sum(sample / r for r in rates)
    line = {
        "impl": "reference",
        "metric": f"tiles_per_s_bc{args.type}_q{args.quality}", "value": value, "unit": "tiles/s",
        "mpix_per_s": value * 4096 / 1e6,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sample / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.type, args.quality, args.batch),
                   "step": f"one pass over a bounded sample of {sample} tiles of that workload on all host cores"},
        "cpu_baseline": {"value": value, "unit": "tiles/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} tiles (first of the {POOL_TILES}-tile pool), {cores} std::threads, static partition, "
                                   "each looping the reference's TileData::encode with -u semantics (BASELINE.md baseline B); "
                                   "squish = restated oracle (libsquish is not vendored by the reference), scalar simd_float path = the parity definition"},
        "e2e": {"value": value, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def oracle_counts(tiles: np.ndarray, type_: int, quality: int, threads: int) -> dict:
    """Work the ORACLE does on `tiles` (instrumented counters of the restated squish), per tile."""
    from oracle import pyoracle
    pyoracle.totals_reset()
    pyoracle.encode_tiles(tiles, type_, quality, threads)
    c = pyoracle.totals()
    return {k: v / len(tiles) for k, v in c.items()}


def copy_peaks(torch, nbytes: int = 256 << 20) -> dict:
    """Pinned-memory copy bandwidth of this GPU's host link, measured live with CUDA events: H2D alone, D2H alone and
    both at once on two streams (GB/s each way).  The denominator of the PCIe roofline of the RangeFit levels."""
    h_a = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=4):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        s1.synchronize(); s2.synchronize()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    def both():
        h2d(); d2h()

    # events recorded on the default stream bracket work on s1/s2 only through the explicit stream syncs above
    t0 = time.perf_counter(); timed(h2d); t_h2d = timed(h2d)
    t_d2h = timed(d2h); t_both = timed(both)
    return {"h2d_gbs": nbytes / t_h2d / 1e9, "d2h_gbs": nbytes / t_d2h / 1e9,
            "bidir_each_gbs": nbytes / t_both / 1e9, "bytes": nbytes,
            "how": "torch pinned<->device copies of 256 MiB, wall time around 4 back-to-back copies bracketed by stream syncs"}


def main() -> int:
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from tileconv_b200 import Encoder, capi, encode_multi, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # a host-side group for the one long wait (rank 0's in-process multi-GPU job): an NCCL barrier would park a
    # spinning kernel on every waiting rank's GPU and time-slice it against rank 0's kernels on that GPU
    cpu_group = dist.new_group(backend="gloo") if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x: float) -> float:
        return reduce_ranks(x, dist.ReduceOp.MAX) if world > 1 else x

    strong = args.scaling == "strong"
    total = args.total if strong else args.batch * world
    first = total * rank // world                       # this rank's tile-index range [first, first + B)
    B = total * (rank + 1) // world - first
    pool = make_pool()
    enc = Encoder(args.type, args.quality, device=local_rank, max_batch_tiles=8192)
    stride = enc.stride
    cores = os.cpu_count() or 1

    # ---- this rank's shard in pinned host memory ------------------------------------------------
    h_in = capi.PinnedBuffer(B * capi.TILE_RECORD)
    h_out = capi.PinnedBuffer(B * stride)
    tiles = h_in.array.reshape(B, capi.TILE_RECORD)
    fill_batch(tiles, pool, first)

    d_in = torch.empty(B * capi.TILE_RECORD, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(B * stride, dtype=torch.uint8, device="cuda")
    d_in.copy_(torch.from_numpy(tiles.reshape(-1)), non_blocking=False)
    torch.cuda.synchronize()

    # ---- kernel-only: W warm-up steps, then exactly K timed steps -----------------------------
    for _ in range(args.warmup):
        enc.encode_device(d_in.data_ptr(), None, B, d_out.data_ptr())
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = enc.launch_count()
    t_wall0 = time.perf_counter()
    ms_kernel = enc.time_kernel_ms(d_in.data_ptr(), None, B, d_out.data_ptr(), args.steps)   # CUDA events on the launch stream
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = enc.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = max_over_ranks(ms_kernel)
    value = total / (ms_step * 1e-3)

    # the device result of the timed run must be the real thing: spot-check it on rank 0 below
    dev_result = d_out.cpu().numpy().reshape(B, stride)

    # ---- end to end through the C ABI: pinned host in -> pinned host out -----------------------
    def time_e2e(e, out_ptr, steps, warm):
        for _ in range(warm):
            e.encode_ptr(h_in.ptr, None, B, out_ptr)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e.encode_ptr(h_in.ptr, None, B, out_ptr)
        torch.cuda.synchronize()
        return max_over_ranks((time.perf_counter() - t0) / steps)

    e2e_s = time_e2e(enc, h_out.ptr, args.steps, min(args.warmup, 2))
    e2e_value = total / e2e_s
    host_result = h_out.array.reshape(B, stride)
    same = bool(np.array_equal(host_result, dev_result))
    shard_sum = float(np.frombuffer(host_result.tobytes(), dtype=np.uint32).sum(dtype=np.uint64) % (1 << 52))

    # ---- live peaks (every rank measures its own GPU; rank 0's go into the line) ----------------
    peak_fused = enc.fp32_peak(True) * 2.0      # FLOP/s, an FFMA counts 2
    peak_unfused = enc.fp32_peak(False)         # FLOP/s of separate FMUL / FADD
    barrier()
    link = copy_peaks(torch) if not args.no_modes else None
    barrier()

    # ---- the metric's other modes on the same tiles: kernel only and host to host ---------------
    modes = {}
    if not args.no_modes:
        for (t_, q_) in ((1, 9), (1, 2), (3, 9), (3, 2)):
            if (t_, q_) == (args.type, args.quality):
                continue
            with Encoder(t_, q_, device=local_rank, max_batch_tiles=8192) as e2:
                d_o2 = torch.zeros(B * e2.stride, dtype=torch.uint8, device="cuda")
                e2.encode_device(d_in.data_ptr(), None, B, d_o2.data_ptr())
                torch.cuda.synchronize()
                ms2 = max_over_ranks(e2.time_kernel_ms(d_in.data_ptr(), None, B, d_o2.data_ptr(), 2))
                del d_o2
                h_o2 = capi.PinnedBuffer(B * e2.stride)
                s2 = time_e2e(e2, h_o2.ptr, 3 if q_ <= 2 else 2, 1)
                h_o2.free()
                modes[f"bc{t_}_q{q_}"] = {"tiles_per_s": total / (ms2 * 1e-3), "ms_per_step": ms2,
                                          "e2e": {"value": total / s2, "unit": "tiles/s", "ms_per_step": s2 * 1e3,
                                                  "h2d_bytes_per_step": B * capi.TILE_RECORD, "d2h_bytes_per_step": B * e2.stride},
                                          "_stride": e2.stride}

        # "next" row N3: the BCn decode kernel is the HBM-bound one of the set (2 KB in, 16 KB out per tile)
        for t_ in (1, 3):
            with Encoder(t_, 2, device=local_rank, max_batch_tiles=4096) as e3:
                rec = torch.from_numpy(e3.encode(pool[:2048])).cuda()
                reps_n = max(1, min(B, 65536) // 2048)
                d_rec = rec.repeat(reps_n, 1).contiguous()
                d_pix = torch.empty(reps_n * 2048 * 4096 * 4, dtype=torch.uint8, device="cuda")
                torch.cuda.synchronize()
                e3.time_decode_ms(d_rec.data_ptr(), reps_n * 2048, d_pix.data_ptr(), 2)
                ms3 = max_over_ranks(e3.time_decode_ms(d_rec.data_ptr(), reps_n * 2048, d_pix.data_ptr(), 5))
                nt = reps_n * 2048
                gbs = nt * (e3.stride + 16384) / (ms3 * 1e-3) / 1e9
                modes[f"decode_bc{t_}"] = {"tiles_per_s": world * nt / (ms3 * 1e-3), "ms_per_step": ms3,
                                           "roofline": {"bound": "hbm", "achieved": gbs, "unit": "GB/s", "bytes_per_tile": e3.stride + 16384}}
                del d_rec, d_pix, rec

    # ---- kernel-only rate per synthetic tile class (rank 0's GPU; the bench mix is S-U + S-G) ----
    classes = {}
    class_pools = {}
    if not args.no_modes and rank == 0:
        nb = 16384
        class_pools = {"S-U": synth.uniform_tiles(512, 1), "S-G": synth.smooth_tiles(512, 2), "S-K": synth.keyed_tiles(512, 3)}
        d_o = torch.zeros(nb * stride, dtype=torch.uint8, device="cuda")
        for name, cp in class_pools.items():
            d_c = torch.from_numpy(np.concatenate([cp] * (nb // len(cp))).reshape(-1)).cuda()
            torch.cuda.synchronize()
            enc.encode_device(d_c.data_ptr(), None, nb, d_o.data_ptr())
            msc = enc.time_kernel_ms(d_c.data_ptr(), None, nb, d_o.data_ptr(), 3)
            classes[name] = {"tiles_per_s": nb / (msc * 1e-3), "tiles": nb}
            del d_c
        del d_o

    if world > 1:
        sums = [None] * world
        dist.all_gather_object(sums, shard_sum)
    else:
        sums = [shard_sum]

    if rank != 0:
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)      # rank 0 runs the in-process multi-GPU job meanwhile (GPU left idle)
        enc.close()
        dist.destroy_process_group()
        return 0

    # ---- rank 0: CPU baseline, parity spot check, rooflines --------------------------------------
    from oracle import pyoracle
    sample = min(args.cpu_sample, POOL_TILES)
    cpu_rate, cores, kind, per_tile = cpu_reference_rate(pool, sample, args.type, args.quality)
    shipped = as_shipped_cli_rate(pool, min(sample, 1024), args.type, args.quality)
    tuned_port = cpu_tuned_port_rate(pool, sample, args.type, args.quality)
    simd_rate = cpu_reference_rate(pool, sample, args.type, args.quality, repeats=2, simd=True)[0]
    seam = pool_seam_rate(tiles[:min(B, 65536)], args.type, args.quality) if world == 1 else None
    # parity of the measured outputs: rank 0's first tiles are pool tiles (copy 0)
    want = pyoracle.encode_tiles(pool[:256], args.type, args.quality, threads=cores)
    parity_ok = bool(np.array_equal(dev_result[:256], want))

    # ---- N > 1: the same N-GPU job in ONE process through tcb200_encode_multi --------------------
    in_process = None
    if world > 1:
        try:
            encs = [enc] + [Encoder(args.type, args.quality, device=g, max_batch_tiles=8192) for g in range(1, world)]
            m_in = capi.PinnedBuffer(total * capi.TILE_RECORD)
            m_out = capi.PinnedBuffer(total * stride)
            mt = m_in.array.reshape(total, capi.TILE_RECORD)
            for g in range(world):
                lo, hi = total * g // world, total * (g + 1) // world
                fill_batch(mt[lo:hi], pool, lo)
            encode_multi(encs, m_in.ptr, None, total, m_out.ptr)          # warm-up (contexts, first touch)
            t0 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                encode_multi(encs, m_in.ptr, None, total, m_out.ptr)
            dt = (time.perf_counter() - t0) / reps
            mo = m_out.array.reshape(total, stride)
            ok = []
            for g in range(world):
                lo, hi = total * g // world, total * (g + 1) // world
                sg = float(np.frombuffer(mo[lo:hi].tobytes(), dtype=np.uint32).sum(dtype=np.uint64) % (1 << 52))
                ok.append(sg == sums[g])
            in_process = {"value": total / dt, "unit": "tiles/s", "n_gpus": world, "ms_per_step": dt * 1e3,
                          "path": "tcb200_encode_multi(): one process, one host thread + 3 streams per device, D2H into disjoint slices of one pinned buffer",
                          "shards_match_rank_results": bool(all(ok)), "vs_torchrun_e2e": (total / dt) / e2e_value}
            for e in encs[1:]:
                e.close()
            m_in.free(); m_out.free()
        except Exception as exc:   # reported, never hidden
            in_process = {"error": str(exc)}
        dist.barrier(group=cpu_group)

    flops_tile = flops_per_tile(per_tile)
    achieved_tflops = flops_tile * (B / (ms_kernel * 1e-3)) / 1e12          # this rank's kernel
    bytes_tile = capi.TILE_RECORD + stride
    achieved_gbs = bytes_tile * (B / (ms_kernel * 1e-3)) / 1e9
    hbm_peak, hbm_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass

    def alu_roofline(rate_per_gpu: float, counts: dict) -> dict:
        ft = flops_per_tile(counts)
        tf = ft * rate_per_gpu / 1e12
        return {"bound": "fp32_alu", "achieved": tf, "peak": peak_fused / 1e12, "unit": "TFLOP/s", "frac": tf / (peak_fused / 1e12),
                "peak_unfused": peak_unfused / 1e12, "frac_of_unfused": tf / (peak_unfused / 1e12), "flops_per_tile": ft,
                "candidates_per_tile": {"c3": counts["cand3"], "c4": counts["cand4"]}}

    for name, m in modes.items():
        if "roofline" in m:          # decode kernels: HBM
            m["roofline"].update({"peak": hbm_peak, "frac": m["roofline"]["achieved"] / hbm_peak, "peak_source": f"MEASURED_PEAKS.json ({hbm_src})",
                                  "note": "the peak is a COPY measurement (half reads, half writes); this stream is 89 % writes and may exceed it"})
            continue
        t_, q_ = int(name[2]), int(name[-1])
        st = m.pop("_stride")
        per_gpu_e2e = m["e2e"]["value"] / world
        if q_ <= 2:
            # RangeFit levels: the kernel needs ~3 % of HBM bandwidth and host to host the link is the limit
            h2d = per_gpu_e2e * capi.TILE_RECORD / 1e9
            d2h = per_gpu_e2e * st / 1e9
            m["roofline"] = {"bound": "pcie_h2d", "achieved": h2d, "peak": link["h2d_gbs"], "unit": "GB/s", "frac": h2d / link["h2d_gbs"],
                             "d2h_achieved": d2h, "d2h_peak_while_h2d_runs": link["bidir_each_gbs"],
                             "limiter": "host->device copy of the 5120-byte tile records over the GPU's host link (kernel-only rate is "
                                        f"{m['tiles_per_s'] / world / 1e6:.1f} M tiles/s per GPU)",
                             "peak_source": "pinned-copy bandwidth measured live (link)"}
            m["kernel_hbm"] = {"bound": "hbm", "achieved": (capi.TILE_RECORD + st) * m["tiles_per_s"] / world / 1e9, "peak": hbm_peak, "unit": "GB/s"}
            m["kernel_hbm"]["frac"] = m["kernel_hbm"]["achieved"] / hbm_peak
        else:
            m["roofline"] = alu_roofline(m["tiles_per_s"] / world, oracle_counts(pool[:256], t_, q_, cores))
    for name, c in classes.items():
        c["roofline"] = alu_roofline(c["tiles_per_s"], oracle_counts(class_pools[name][:64], args.type, args.quality, cores))

    traffic, ncu_facts = None, None
    for prof in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", prof)) as f:
                t = json.load(f).get(f"bc{args.type}_q{args.quality}_batch{B}")
            if t:
                traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
                ncu_facts = {k: t[k] for k in ("smsp_issue_active_pct", "sm_pipe_fma_cycles_active_pct", "active_lanes_per_instruction",
                                               "dominant_stall") if k in t}
                ncu_facts["source"] = "profiles/" + prof
                break
        except Exception:
            pass

    main_roof = alu_roofline(B / (ms_kernel * 1e-3), per_tile)
    main_roof.update({"traffic": traffic,
                      "traffic_note": "DRAM bytes of one launch from the committed ncu capture (profiles/); algorithmic bytes per launch = " + str(bytes_tile * B),
                      "peak_source": "FFMA loop measured live by tcb200_fp32_peak (an FFMA = 2 flops)",
                      "note": "bit-exactness forbids mul+add fusion, so the ceiling of EXECUTING the reference's arithmetic is peak_unfused (separate "
                              "FMUL/FADD).  `achieved` counts the reference algorithm's work (the candidates the oracle evaluated on the same tiles); "
                              "since round 2 the kernel proves by an exact lower bound that most 4-colour candidates cannot win and skips them "
                              "(bc_device.cuh run_sweep; skipped fractions and the audit of the bound: profiles/r2_prune_audit.jsonl), so "
                              "frac_of_unfused can exceed 1",
                      "ncu": ncu_facts,      # from the committed capture of this configuration, not measured in this run
                      "pruning": pruning_facts(args.type, args.quality),
                      "classes": {k: {"tiles_per_s": v["tiles_per_s"], "frac": v["roofline"]["frac"], "frac_of_unfused": v["roofline"]["frac_of_unfused"],
                                      "achieved": v["roofline"]["achieved"]} for k, v in classes.items()},
                      "hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                              "bytes_per_tile": bytes_tile, "peak_source": f"MEASURED_PEAKS.json ({hbm_src})"}})

    line = {
        "metric": f"tiles_per_s_bc{args.type}_q{args.quality}", "value": value, "unit": "tiles/s",
        "mpix_per_s": value * 4096 / 1e6,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.type, args.quality, B) + (f" (strong scaling: one {total}-tile job)" if strong else ""),
                   "l2": f"inputs larger than L2: {B * capi.TILE_RECORD / 1e6:.0f} MB read per step per GPU",
                   "timing": "CUDA events on the launching stream around K back-to-back launches, max over ranks"},
        "e2e": {"value": e2e_value, "unit": "tiles/s", "h2d_bytes_per_step": B * capi.TILE_RECORD, "d2h_bytes_per_step": B * stride,
                "ms_per_step": e2e_s * 1e3, "path": "tcb200_encode(): pinned host -> H2D -> kernel -> D2H -> pinned host, 8192-tile chunks on 3 streams",
                "matches_device_result": same},
        "e2e_in_process": in_process,
        "gpu_launches": int(launches),
        "wall_ms_per_step": 1e3 * t_wall / args.steps,
        "clocks": clocks,
        "roofline": main_roof,
        "cpu_baseline": {"value": cpu_rate, "unit": "tiles/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} pool tiles, {cores} threads over the reference's TileData::encode (-u), restated squish (scalar simd_float path = the "
                                   "parity definition), best of 3 passes",
                         "simd_squish": {"value": simd_rate, "unit": "tiles/s",
                                         "note": "the same run with the cluster solve on 4-wide SSE2 vectors (identical records): the speed of a libsquish "
                                                 "built with SQUISH_USE_SSE, whose own arithmetic (rcpps) is the `sse` knob of tcb200_create_ex"},
                         "as_shipped_cli": shipped, "tuned_port": tuned_port},
        "e2e_pool_seam": seam,
        "parity_spot_check": parity_ok,
        "link": link,
        "modes": modes,
        "classes": classes,
    }
    print(json.dumps(line), flush=True)
    enc.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
