"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: block ownership and the position
exchange that runs between drift and force (rebound_b200/distributed.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rebound_b200 import distributed as D


def test_shard_ranges_partition():
    for n in (0, 1, 7, 16384, (1 << 20) + 10):
        for world in (1, 2, 3, 8):
            r = [D.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank starts from the same state, then "drifts" only its own block
        rng = np.random.default_rng(5)
        base = [torch.from_numpy(rng.normal(size=n)) for _ in range(3)]
        fields = [b.clone() for b in base]
        b0, e0 = D.shard_range(n, rank, world)
        for f in fields:
            f[b0:e0] += 1000.0 * (rank + 1)
        ex = D.BlockExchange(fields)
        ex()
        ok = True
        for f, b in zip(fields, base):
            want = b.clone()
            for r in range(world):
                rb, re = D.shard_range(n, r, world)
                want[rb:re] += 1000.0 * (r + 1)
            ok = ok and torch.equal(f, want)
        out[rank] = 1 if ok and ex.calls == 1 else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 65, 3])
def test_block_exchange_gloo_world2(n):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1] * world
