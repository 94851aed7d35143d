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


def test_exchange_fields_follow_the_request_mask():
    from rebound_b200 import abi

    assert D.exchange_fields(abi.EXCHANGE_POSITIONS) == [0, 1, 2]
    assert D.exchange_fields(abi.EXCHANGE_POSITIONS | abi.EXCHANGE_VELOCITIES) == [0, 1, 2, 3, 4, 5]
    assert D.exchange_fields(abi.EXCHANGE_ALL) == list(range(14))


def _serial_list(n, n_ghost, seed):
    """A made-up 'serial' collision list in DIRECT order (ghost box, projectile, target)."""
    from rebound_b200 import abi

    rng = np.random.default_rng(seed)
    rows = []
    for g in range(n_ghost):
        for i in range(n):
            for j in sorted(rng.choice(n, size=min(n, int(rng.integers(0, 3))), replace=False)):
                rows.append((i, j, float(g)))
    out = np.zeros(len(rows), dtype=abi.COLLISION_DTYPE)
    for k, (i, j, g) in enumerate(rows):
        out[k]["p1"], out[k]["p2"], out[k]["gb_x"] = i, j, g
    return out


class _FakeEngine:
    """What rebcu_collision_search leaves on one rank: the serial list restricted to its projectile block."""

    def __init__(self, serial, n, n_ghost, rank, world):
        b, e = D.shard_range(n, rank, world)
        keep = (serial["p1"] >= b) & (serial["p1"] < e)
        self.local = serial[keep]
        self.seg = [int(np.sum(self.local["gb_x"] == float(g))) for g in range(n_ghost)]

    def collisions_fetch(self):
        return self.local

    def collisions_segments(self):
        return self.seg


def test_merge_collision_segments_restores_serial_order():
    for n, n_ghost, world in ((17, 1, 2), (17, 9, 2), (40, 9, 3), (5, 27, 8)):
        serial = _serial_list(n, n_ghost, seed=n + n_ghost)
        fakes = [_FakeEngine(serial, n, n_ghost, r, world) for r in range(world)]
        merged = D.merge_collision_segments([f.local for f in fakes], [f.seg for f in fakes])
        assert merged.tobytes() == serial.tobytes()
    from rebound_b200 import abi

    assert len(D.merge_collision_segments([np.zeros(0, abi.COLLISION_DTYPE)] * 2, [[0, 0], [0, 0]])) == 0


def _gather_worker(rank, world, port, n, n_ghost, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        serial = _serial_list(n, n_ghost, seed=11)
        merged = D.gather_collisions(_FakeEngine(serial, n, n_ghost, rank, world), "cpu")
        out[rank] = 1 if merged.tobytes() == serial.tobytes() else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,n_ghost", [(33, 9), (2, 1), (1, 9)])
def test_gather_collisions_gloo_world2(n, n_ghost):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, n, n_ghost, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1] * world
