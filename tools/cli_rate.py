"""Whole-program rates on a synthetic headerless TIS (1 Mi tiles by default, in /dev/shm when there is room):
  * this repo's driver `tbc_encode` -- wall clock of the process AND the steady-state figures of its slab pipeline
    (TCB200_TIMING=1: CUDA start-up excluded, per-stage busy times);
  * the same driver on the reference-style TileData route (TCB200_TILEDATA_ROUTE=1);
  * the reference CLI as shipped and with this repo's plugin dropped in (oracle/_ref, built where /root/reference exists).
usage: python tools/cli_rate.py [tiles]      -> JSON lines on stdout"""
import json, os, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
pool = bench.make_pool()
TBC = os.path.join(ROOT, "tileconv_b200", "tbc_encode")
REF = os.path.join(ROOT, "oracle", "_ref", "tileconv_ref")
PLUG = os.path.join(ROOT, "oracle", "_ref", "tileconv_cuda")
base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > N * 9000 else None


def write_tis(path, n):
    with open(path, "wb") as f:
        step = 65536
        buf = np.zeros((step, 5120), np.uint8)
        for lo in range(0, n, step):
            m = min(step, n - lo)
            bench.fill_batch(buf[:m], pool, lo)
            f.write(buf[:m].tobytes())


def run(cmd, env=None, n=0, label="", args=""):
    t0 = time.perf_counter()
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    dt = time.perf_counter() - t0
    row = {"program": label, "args": args, "tiles": n, "rc": res.returncode, "wall_s": round(dt, 3), "wall_tiles_per_s": round(n / dt, 1)}
    for line in res.stderr.splitlines():
        if line.startswith("{"):
            row["pipeline"] = json.loads(line)
    print(json.dumps(row), flush=True)
    return row


with tempfile.TemporaryDirectory(dir=base) as tmp:
    src, dst = os.path.join(tmp, "in.tis"), os.path.join(tmp, "out.tbc")
    write_tis(src, N)
    env_t = dict(os.environ, TCB200_TIMING="1")
    for args in (["-u", "-t", "1", "-q", "-9"], ["-u", "-t", "1", "-q", "-2"], ["-u", "-t", "3", "-q", "-2"], ["-t", "1", "-q", "-9"], ["-t", "1", "-q", "-2"]):
        cmd = [TBC, "-s", "-T", "-j", "0", "-o", dst] + args + [src]
        subprocess.run(cmd, capture_output=True)                       # warm-up: page cache, driver
        run(cmd, env_t, N, "tbc_encode (slab pipeline)", " ".join(args))
    small = min(N, 262144)
    src2 = os.path.join(tmp, "small.tis")
    write_tis(src2, small)
    for args in (["-u", "-t", "1", "-q", "-9"], ["-u", "-t", "1", "-q", "-2"]):
        run([TBC, "-s", "-T", "-j", "0", "-o", dst] + args + [src2], dict(os.environ, TCB200_TILEDATA_ROUTE="1"), small, "tbc_encode (TileData route)", " ".join(args))
        if os.path.exists(PLUG):
            run([PLUG, "-s", "-T", "-j", "0", "-o", dst] + args + [src2], None, small, "reference CLI + plugin", " ".join(args))
    if os.path.exists(REF):
        src3 = os.path.join(tmp, "tiny.tis")
        write_tis(src3, 2048)
        run([REF, "-s", "-T", "-j", "0", "-o", dst, "-u", "-t", "1", "-q", "-9", src3], None, 2048, "reference CLI as shipped", "-u -t 1 -q -9")
