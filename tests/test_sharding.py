"""Host logic of the multi-GPU path on CPU: tile-range sharding under a world_size-2 gloo group.  The
data path has no collective; the group is only used the way bench.py uses it --
barrier, MAX over ranks of the step time, SUM of tiles -- and to prove that the shards' records,
gathered in rank order, equal the single-rank result.  The per-shard "encoder" here is the CPU oracle
standing in for the device (this is a test of the plumbing, not of the kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tileconv_b200.sharding import tile_range


def test_ranges_partition_the_index_space():
    for n in (0, 1, 7, 1024, 1000003):
        for world in (1, 2, 3, 8):
            edges = [tile_range(r, world, n) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        tile_range(2, 2, 10)


def _worker(rank, world, port, n, tmp):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import pyoracle
    from tileconv_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tiles = synth.mixed_tiles(n, 81)
    lo, hi = tile_range(rank, world, n)
    out = np.zeros((n, pyoracle.record_size(1)), np.uint8)
    out[lo:hi] = pyoracle.encode_tiles(tiles[lo:hi], 1, 2)           # this rank's disjoint slice
    dist.barrier()
    step = torch.tensor([1.0 + rank], dtype=torch.float64)           # pretend step times
    dist.all_reduce(step, op=dist.ReduceOp.MAX)
    count = torch.tensor([float(hi - lo)], dtype=torch.float64)
    dist.all_reduce(count, op=dist.ReduceOp.SUM)
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, out[lo:hi].tobytes()))
    if rank == 0:
        whole = np.zeros_like(out)
        for plo, phi, blob in parts:
            whole[plo:phi] = np.frombuffer(blob, np.uint8).reshape(phi - plo, -1)
        ok = np.array_equal(whole, pyoracle.encode_tiles(tiles, 1, 2)) and step.item() == float(world) and count.item() == float(n)
        open(os.path.join(tmp, "ok"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_two_rank_gloo_shards_reassemble(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 13, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"
