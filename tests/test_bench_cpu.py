"""bench.py's reference arm and host-side helpers run without a GPU: the JSON contract of `--impl reference`, the
batch filler (byte-distinct copies with identical pixels) and the flop model."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "48"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "tiles/s" and line["higher_is_better"] is True
    assert line["metric"] == "tiles_per_s_bc1_q9" and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    assert "configs[4]" in line["config"]["workload"]


def test_non_zero_ranks_of_the_reference_arm_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_fill_batch_makes_byte_distinct_copies_with_identical_pixels(oracle):
    import bench
    from tileconv_b200 import synth
    pool = synth.mixed_tiles(8, 5)
    batch = np.zeros((24, 5120), np.uint8)
    bench.fill_batch(batch, pool, first_global_tile=8)       # copies 1, 2, 3 of the pool
    assert len({bytes(t) for t in batch}) == 24 and not any(np.array_equal(t, p) for t in batch[:8] for p in pool)
    want = oracle.encode_tiles(pool, 1, 4, threads=2)
    for lo in range(0, 24, 8):
        assert np.array_equal(oracle.encode_tiles(batch[lo:lo + 8], 1, 4, threads=2), want)


def test_flop_model_matches_survey_8d():
    import bench
    per_tile = {"cand3": 10.0, "cand4": 100.0, "points": 16.0, "cluster_blocks": 1.0}
    assert bench.flops_per_tile(per_tile) == 10 * 138 + 100 * 160 + 30 * 16 + 216


def test_pruning_facts_come_from_the_committed_audit():
    import bench
    f = bench.pruning_facts(1, 9)
    assert f is not None and set(f) >= {"S-U", "S-G", "S-K", "source"}
    assert all(f[c]["wrongly_skipped"] == 0 and 0.3 < f[c]["skipped_frac"] < 1.0 for c in ("S-U", "S-G", "S-K"))
    assert bench.pruning_facts(1, 2) is None          # RangeFit levels have no cluster sweeps
