"""N>1 host logic on CPU: two gloo ranks shard a batch and all-gather metric partials."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dsin_b200.dist import gather_metrics, shard_range


def test_shard_range_partitions_exactly():
    for n in (1, 7, 8, 32, 255, 256):
        for w in (1, 2, 4, 8):
            blocks = [shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bits = [100.0, 250.0, 40.0, 10.0, 75.0]  # per-image bit counts of a 5-image batch
        lo, hi = shard_range(len(bits), rank, world)
        mine = bits[lo:hi]
        res = gather_metrics(sum(mine), 1000.0 * len(mine), 0.9 * len(mine), len(mine))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    out = dict(q.get(timeout=10) for _ in range(2))
    single = gather_metrics(475.0, 5000.0, 4.5, 5)
    for r in (0, 1):
        assert out[r]["bpp"] == pytest.approx(single["bpp"], rel=1e-12)
        assert out[r]["msssim"] == pytest.approx(0.9, rel=1e-12)
        assert out[r]["n_images"] == 5
    assert out[0]["per_rank"] == out[1]["per_rank"]
    assert out[0]["per_rank"][0][3] == 3 and out[0]["per_rank"][1][3] == 2
