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


def sharded(world, p, fn, transport=abi.TRANSPORT_LOCAL, shard_build=2):
    """Uploads p on `world` engines of device 0, shards them, runs fn(rank, engine) on every engine in its own thread."""
    engines = [Engine(0) for _ in range(world)]
    try:
        for e in engines:
            e.upload(np.ascontiguousarray(p))
            e.set_sharded_build(shard_build)
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
        nrm = np.linalg.norm(b, axis=1)
        rel = np.linalg.norm(a - b, axis=1) / np.maximum(nrm, 0.1 * np.median(nrm))     # the star's net pull nearly cancels
        assert np.sqrt(np.mean(rel**2)) < 5e-3 and np.all(np.isfinite(a))
    finally:
        close(grp)


@pytest.mark.parametrize("world", [2, 3, 5])
@pytest.mark.parametrize("case", ["disc_tree", "open_tree", "sheet_rootboxes", "disc_20k", "deep_pairs", "plummer_3d"])
def test_sharded_tree_build_bitwise_on_one_gpu(case, world):
    """Key-range ownership of the tree BUILD (rebcu_set_sharded_build(1): every rank sorts and builds only the subtrees of
    its buckets, the traversal records are all-gathered, the top cells filled in by everyone): accelerations and
    trajectories must be the single-device bits -- discs, an open box that loses particles, a shearing sheet with 2x2
    root boxes and 25 ghost boxes, pairs deeper than the 63-bit key, a 3-d cluster."""
    if case in ("disc_tree", "open_tree"):
        p, cfg, steps = make_case(case)
    elif case == "sheet_rootboxes":
        p = ics.shearing_sheet(root_size=40.0, seed=5)
        cfg, steps = ics.shearing_sheet_config(root_size=40.0, t=55.5, collision=abi.COLLISION_NONE), 3
    elif case == "disc_20k":
        p, cfg, steps = ics.selfgravity_disc(20000, seed=12), ics.selfgravity_disc_config(), 2
    elif case == "deep_pairs":
        p = ics.selfgravity_disc(400, seed=6)
        p["x"][201:] = p["x"][1:201] + 1e-9 * np.arange(1, 201)
        p["y"][201:] = p["y"][1:201] - 3e-10
        p["z"][201:] = p["z"][1:201]
        cfg, steps = ics.selfgravity_disc_config(), 2
    else:
        p = ics.plummer(5000, seed=4)
        cfg, steps = ics.plummer_config(5000, gravity=abi.GRAVITY_TREE, root_size=200.0, opening_angle2=0.25, boundary=abi.BOUNDARY_OPEN), 2
    want, _, _ = checkers.oracle().steps(cfg, p, steps)

    def run(rank, eng):
        c = cfg.copy()
        eng.steps(c, steps)
        eng.exchange(abi.EXCHANGE_ALL)
        return eng.download() if rank == 0 else None

    grp, out = sharded(world, p, run, shard_build=1)
    try:
        got = out[0]
        assert len(got) == len(want)
        assert checkers.bits_equal(got, want)
    finally:
        close(grp)


def test_sharded_tree_build_errors_reach_every_rank():
    """Two particles with the same coordinates sit in ONE rank's key range; every rank must report the reference's
    error (the flags travel with the gathered table), none may hang in a collective."""
    from rebound_b200.simulation import ReboundCudaError

    p = ics.selfgravity_disc(3000, seed=2)
    p["x"][700] = p["x"][300]; p["y"][700] = p["y"][300]; p["z"][700] = p["z"][300]
    cfg = ics.selfgravity_disc_config()

    def run(rank, eng):
        try:
            eng.update_acceleration(cfg.copy())
        except ReboundCudaError as e:
            return e.msg
        return None

    grp, out = sharded(3, p, run, shard_build=1)
    try:
        assert out == ["Cannot add two particles with the same coordinates to the tree."] * 3
    finally:
        close(grp)


def test_sharded_tree_build_fast_walk():
    """FAST group walk on records produced by the sharded build."""
    p = ics.selfgravity_disc(20000, seed=6)
    cfg = ics.selfgravity_disc_config(boundary=abi.BOUNDARY_NONE)
    strict = cfg.copy()
    cfg.mode = abi.MODE_FAST
    want, _ = checkers.oracle().gravity(strict, p)

    def run(rank, eng):
        eng.update_acceleration(cfg.copy())
        eng.exchange(abi.EXCHANGE_ALL)
        return eng.download() if rank == 0 else None

    grp, out = sharded(4, p, run, shard_build=1)
    try:
        got = out[0]
        a = np.stack([got["ax"], got["ay"], got["az"]], 1)
        b = np.stack([want["ax"], want["ay"], want["az"]], 1)
        nrm = np.linalg.norm(b, axis=1)
        rel = np.linalg.norm(a - b, axis=1) / np.maximum(nrm, 0.1 * np.median(nrm))     # the star's net pull nearly cancels
        assert np.sqrt(np.mean(rel**2)) < 5e-3 and np.all(np.isfinite(a))
    finally:
        close(grp)
