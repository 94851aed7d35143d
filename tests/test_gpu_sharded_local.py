"""Sharded runs on ONE GPU: W engines in one process (a thread each), all on device 0, joined by the engine's LOCAL
exchange transport (csrc/comm.cu: peer copies fenced by events and host barriers; NCCL refuses two ranks on one
device).  Everything else is the multi-GPU code path -- rebcu_set_shard ranges, the exchange between drift and
force, the full-range boundary check with re-cut blocks, tree_shard_list walks, sharded collision searches merged
segment by segment -- so a box with a single GPU gives it correctness evidence: the assembled state / list must equal
the single-device oracle result bit for bit (STRICT mode does not depend on the sharding).
tests/test_gpu_multi.py runs the same cases over NCCL with one process per GPU when >= 2 devices are present."""
import numpy as np
import pytest

import checkers
from rebound_b200 import abi, distributed as D, ics
from rebound_b200.simulation import Engine
from test_gpu_multi import _extras_setup, make_case, make_collision_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]      # a rank that dies leaves its peers in a barrier


def sharded(world, p, fn, transport=abi.TRANSPORT_LOCAL):
    """Uploads p on `world` engines of device 0, shards them, runs fn(rank, engine) on every engine in its own thread."""
    engines = [Engine(0) for _ in range(world)]
    try:
        for e in engines:
            e.upload(np.ascontiguousarray(p))
        grp = D.LocalGroup(engines, transport)
        out = grp.run(fn)
        return grp, out
    finally:
        pass


def close(grp):
    for e in grp.engines:
        e.close()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["plummer_basic", "plummer_comp", "testp_type1", "testp_type1_rows", "disc_tree",
                                  "open_basic", "open_tree"])
def test_sharded_steps_bitwise_on_one_gpu(case, world):
    p, cfg, steps = make_case(case)
    want, _, _ = checkers.oracle().steps(cfg, p, steps)

    def run(rank, eng):
        c = cfg.copy()
        eng.steps(c, steps)
        eng.exchange(abi.EXCHANGE_ALL)
        eng.synchronize()
        st = eng.comm_stats()
        return (eng.download() if rank == 0 else None), st, c

    grp, out = sharded(world, p, run)
    try:
        got, st, c = out[0]
        assert st["transport"] == "local" and st["exchanges"] >= steps
        assert len(got) == len(want)
        assert checkers.bits_equal(got, want)
        assert np.array_equal(got["name"], want["name"])
        if case.startswith("open"):
            assert 0 < len(want) < len(p) - 20          # the case does remove particles, across block borders
    finally:
        close(grp)


@pytest.mark.parametrize("case", ["sheet_tree", "sheet_direct", "sheet_line", "sheet_linetree"])
def test_sharded_collision_search_bitwise_on_one_gpu(case):
    p, cfg, steps, mode = make_collision_case(case)
    orc = checkers.oracle()
    c0 = cfg.copy()
    c0.collision = abi.COLLISION_NONE
    q, c1, _ = orc.steps(c0, p, steps)
    c1.collision = mode
    want = orc.collision_search(c1, q)
    assert len(want) > 0

    def run(rank, eng):
        c = cfg.copy()
        c.collision = abi.COLLISION_NONE
        eng.steps(c, steps)
        c.collision = mode
        before = eng.comm_stats()["bytes_received"]
        local = eng.collision_search(c)
        b, e = eng.shard_range()
        assert np.all((local["p1"] >= b) & (local["p1"] < e))
        # x y z vx vy vz of the other blocks were gathered for the search
        assert eng.comm_stats()["bytes_received"] - before == 6 * 8 * (eng.N - (e - b))
        return len(local)

    grp, out = sharded(2, p, run)
    try:
        got = grp.collisions()
        assert sum(out) == len(got)
        assert checkers.collisions_equal(got, want, with_ri=(mode in (abi.COLLISION_TREE, abi.COLLISION_LINETREE)))
    finally:
        close(grp)


def test_sharded_jerk_exit_checks_and_subset_search_bitwise_on_one_gpu():
    p, cfg, steps, sub, nt, v = _extras_setup()
    orc = checkers.oracle()
    q, c1, _ = orc.steps(cfg, p, steps)
    q = orc.apply_jerk(c1, q, v)
    status = orc.exit_check(c1, q, 30.0, 1.5)
    c1.collision = abi.COLLISION_DIRECT
    want_col = orc.collision_search_subset(c1, q, sub, nt)
    assert len(want_col) > 0 and status != 0

    def run(rank, eng):
        c = cfg.copy()
        eng.steps(c, steps)
        eng.apply_jerk(c, v)
        flags = eng.exit_check(30.0, 1.5)
        c.collision = abi.COLLISION_DIRECT
        eng.set_collision_subset(sub, nt)
        eng.collision_search(c)
        return flags

    grp, out = sharded(2, p, run)
    try:
        got_col = grp.collisions()
        for e in grp.engines:
            e.set_collision_subset()
        grp.run(lambda r, e: e.exchange(abi.EXCHANGE_ALL))
        got = grp.engines[0].download()
        assert checkers.bits_equal(got, q)
        flags = out[0]
        assert (3 if flags[1] else (4 if flags[0] else 0)) == status
        assert checkers.collisions_equal(got_col, want_col, with_ri=False)
    finally:
        close(grp)


def test_sharded_host_blocks_upload_and_download():
    """rebcu_upload_shard / rebcu_download_shard: every rank's host memory holds only its block; the other blocks
    arrive over the exchange.  The stitched blocks equal the oracle's full state."""
    p, cfg, steps = make_case("plummer_basic")
    want, _, _ = checkers.oracle().steps(cfg, p, steps)
    world = 3
    engines = [Engine(0) for _ in range(world)]
    try:
        grp = D.LocalGroup(engines, abi.TRANSPORT_LOCAL)
        n = len(p)
        blocks = [np.ascontiguousarray(p[D.shard_range(n, r, world)[0]:D.shard_range(n, r, world)[1]]) for r in range(world)]

        def run(rank, eng):
            eng.upload_shard(blocks[rank], n)
            c = cfg.copy()
            eng.steps(c, steps)
            out = np.zeros(len(blocks[rank]), dtype=abi.PARTICLE_DTYPE)
            return eng.download_shard(out)

        out = grp.run(run)
        got = np.concatenate(out)
        assert checkers.bits_equal(got, want)
    finally:
        for e in engines:
            e.close()


def test_sharded_fast_tree_matches_single_engine():
    """FAST group walk on a shard list (groups of 32 key-adjacent particles of the rank's own block): same accuracy
    class as the single-engine FAST walk -- compared against the strict tree with the reference-level tolerance."""
    p = ics.selfgravity_disc(6000, seed=6)
    cfg = ics.selfgravity_disc_config(boundary=abi.BOUNDARY_NONE)
    cfg.mode = abi.MODE_FAST
    strict = cfg.copy(); strict.mode = abi.MODE_STRICT
    want, _ = checkers.oracle().gravity(strict, p)

    def run(rank, eng):
        eng.update_acceleration(cfg.copy())
        eng.exchange(abi.EXCHANGE_ALL)
        return eng.download() if rank == 0 else None

    grp, out = sharded(2, p, run)
    try:
        got = out[0]
        a = np.stack([got["ax"], got["ay"], got["az"]], 1)
        b = np.stack([want["ax"], want["ay"], want["az"]], 1)
        rel = np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)
        assert np.sqrt(np.mean(rel**2)) < 5e-3 and np.all(np.isfinite(a))
    finally:
        close(grp)
