"""Two-GPU parity (needs >= 2 devices, skipped otherwise): targets sharded in contiguous blocks, positions
all-gathered over NCCL between drift and force; the assembled state must equal the single-device /
oracle result bit for bit (STRICT mode does not depend on the sharding)."""
import os
import socket

import numpy as np
import pytest
import torch

import checkers
from rebound_b200 import abi, ics

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, path):
    import torch.distributed as dist

    from rebound_b200 import distributed as D
    from rebound_b200.simulation import Engine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        p, cfg, steps = make_case(case)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        eng = Engine(rank, stream.cuda_stream)
        eng.upload(np.ascontiguousarray(p))
        state = D.attach(eng, torch.device("cuda", rank))
        eng.set_sharded_build(1)            # tree cases: per-rank subtree builds, records gathered over NCCL (ragged broadcasts)
        c = cfg.copy()
        eng.steps(c, steps)
        D.gather_owned(eng, torch.device("cuda", rank))
        torch.cuda.synchronize()
        if rank == 0:
            out = eng.download()
            np.save(path, np.frombuffer(out.tobytes(), dtype=np.uint8))
            assert state["calls"] >= steps
        eng.close()
    finally:
        dist.destroy_process_group()


def _collision_worker(rank, world, port, case, path):
    """Sharded steps without collisions (so velocities diverge between ranks), then one sharded collision
    search; the merged list goes to `path`."""
    import torch.distributed as dist

    from rebound_b200 import distributed as D
    from rebound_b200.simulation import Engine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        p, cfg, steps, mode = make_collision_case(case)
        dev = torch.device("cuda", rank)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        eng = Engine(rank, stream.cuda_stream)
        eng.upload(np.ascontiguousarray(p))
        state = D.attach(eng, dev)
        c = cfg.copy()
        c.collision = abi.COLLISION_NONE
        eng.steps(c, steps)
        c.collision = mode
        fields_before = state.get("fields", 0)
        local = eng.collision_search(c)
        assert state["fields"] - fields_before == 6          # x y z vx vy vz were gathered for the search
        assert state["transport"] == "nccl"
        b, e = eng.shard_range()
        assert np.all((local["p1"] >= b) & (local["p1"] < e))
        merged = D.gather_collisions(eng, dev)
        torch.cuda.synchronize()
        if rank == 0:
            np.save(path, np.frombuffer(merged.tobytes(), dtype=np.uint8))
        eng.close()
    finally:
        dist.destroy_process_group()


def make_collision_case(case):
    p = ics.shearing_sheet(root_size=40.0, seed=5)
    base = dict(root_size=40.0, t=55.5)
    if case == "sheet_tree":
        return p, ics.shearing_sheet_config(**base), 2, abi.COLLISION_TREE
    if case == "sheet_direct":
        return p, ics.shearing_sheet_config(**base), 2, abi.COLLISION_DIRECT
    if case == "sheet_line":
        return p, ics.shearing_sheet_config(**base), 2, abi.COLLISION_LINE
    if case == "sheet_linetree":
        return p, ics.shearing_sheet_config(**base), 2, abi.COLLISION_LINETREE
    raise ValueError(case)


def make_case(case):
    if case == "plummer_basic":
        return ics.plummer(3001, seed=3), ics.plummer_config(3001), 3
    if case == "plummer_comp":
        return ics.plummer(2048, seed=4), ics.plummer_config(2048, gravity=abi.GRAVITY_COMPENSATED), 2
    if case == "testp_type1":
        q = ics.planetesimal_disk(2000, seed=5)
        q["m"][10:] = 1e-9
        return q, ics.planetesimal_config(testparticle_type=1), 2
    if case == "testp_type1_rows":
        # N >= 4096: the massive rows take the term-buffer path on the rank that owns them
        q = ics.planetesimal_disk(6000, seed=8)
        q["m"][10:] = 1e-9
        return q, ics.planetesimal_config(testparticle_type=1), 2
    if case == "disc_tree":
        return ics.selfgravity_disc(3000, seed=6), ics.selfgravity_disc_config(boundary=abi.BOUNDARY_NONE), 2
    if case in ("open_basic", "open_tree"):
        # a hot cluster in a small open box: particles leave during the run, on both ranks' blocks
        q = ics.plummer(1500, seed=7)
        q["vx"] *= 6.0
        q["vy"] *= 6.0
        q["vz"] *= 6.0
        q["x"] *= 0.3
        q["y"] *= 0.3
        q["z"] *= 0.3
        q["name"] = np.arange(1, len(q) + 1, dtype=np.uint64)      # the tag fields must travel with the particle
        c = ics.plummer_config(1500, boundary=abi.BOUNDARY_OPEN, root_size=4.0, dt=0.05,
                               gravity=abi.GRAVITY_BASIC if case == "open_basic" else abi.GRAVITY_TREE)
        return q, c, 12
    raise ValueError(case)


def _run(target, case, path, world=2):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    for attempt in range(3):            # the free port can be taken between the probe and the rendezvous: try another one
        port = _free_port()
        procs = [ctx.Process(target=target, args=(r, world, port, case, path)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
        if all(p.exitcode == 0 for p in procs):
            return
        for p in procs:
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case", ["open_basic", "open_tree"])
def test_two_gpu_open_boundary_removal_bitwise(case, tmp_path):
    """Particles leaving an open box while sharded: every field is gathered from its owner before the
    compaction (rebcu_exchange_request == ALL), the blocks are re-cut, and the run continues bit-identically."""
    path = str(tmp_path / "out.npy")
    _run(_worker, case, path)
    got = np.frombuffer(np.load(path).tobytes(), dtype=abi.PARTICLE_DTYPE)
    p, cfg, steps = make_case(case)
    want, _, _ = checkers.oracle().steps(cfg, p, steps)
    assert 0 < len(want) < len(p) - 20          # the case does remove particles
    assert len(got) == len(want)
    assert checkers.bits_equal(got, want)
    assert np.array_equal(got["name"], want["name"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case", ["sheet_tree", "sheet_direct", "sheet_line", "sheet_linetree"])
def test_two_gpu_sharded_collision_search_bitwise(case, tmp_path):
    path = str(tmp_path / "col.npy")
    _run(_collision_worker, case, path)
    got = np.frombuffer(np.load(path).tobytes(), dtype=abi.COLLISION_DTYPE)
    p, cfg, steps, mode = make_collision_case(case)
    orc = checkers.oracle()
    c0 = cfg.copy()
    c0.collision = abi.COLLISION_NONE
    q, c1, _ = orc.steps(c0, p, steps)
    c1.collision = mode
    want = orc.collision_search(c1, q)
    assert len(want) > 0
    assert checkers.collisions_equal(got, want, with_ri=(mode in (abi.COLLISION_TREE, abi.COLLISION_LINETREE)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case", ["plummer_basic", "plummer_comp", "testp_type1", "testp_type1_rows", "disc_tree"])
def test_two_gpu_sharded_steps_bitwise(case, tmp_path):
    path = str(tmp_path / "out.npy")
    _run(_worker, case, path)
    got = np.frombuffer(np.load(path).tobytes(), dtype=abi.PARTICLE_DTYPE)
    p, cfg, steps = make_case(case)
    want, _, _ = checkers.oracle().steps(cfg, p, steps)
    assert checkers.bits_equal(got, want)


def _extras_setup():
    p = ics.shearing_sheet(root_size=40.0, seed=5)
    cfg = ics.shearing_sheet_config(root_size=40.0, t=55.5, collision=abi.COLLISION_NONE)
    rng = np.random.default_rng(11)
    sub = rng.permutation(len(p))[: (2 * len(p)) // 3].astype(np.uint64)
    return p, cfg, 2, sub, len(sub) // 2, 0.37


def _extras_worker(rank, world, port, case, path):
    """Sharded steps, then the jerk kick (needs every block's accelerations: exchange ALL), the exit checks (every
    rank scans all particles) and a DIRECT collision search on an r->map / N_targets subset (slots are sharded)."""
    import torch.distributed as dist

    from rebound_b200 import distributed as D
    from rebound_b200.simulation import Engine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        p, cfg, steps, sub, nt, v = _extras_setup()
        dev = torch.device("cuda", rank)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        eng = Engine(rank, stream.cuda_stream)
        eng.upload(np.ascontiguousarray(p))
        D.attach(eng, dev)
        c = cfg.copy()
        eng.steps(c, steps)
        eng.apply_jerk(c, v)
        flags = eng.exit_check(30.0, 1.5)
        c.collision = abi.COLLISION_DIRECT
        eng.set_collision_subset(sub, nt)
        eng.collision_search(c)
        merged = D.gather_collisions(eng, dev)
        eng.set_collision_subset()
        D.gather_owned(eng, dev)
        torch.cuda.synchronize()
        if rank == 0:
            np.save(path, np.frombuffer(eng.download().tobytes(), dtype=np.uint8))
            np.save(path + ".col.npy", np.frombuffer(merged.tobytes(), dtype=np.uint8))
            np.save(path + ".flags.npy", np.array(flags, dtype=np.int64))
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_jerk_exit_checks_and_subset_search_bitwise(tmp_path):
    path = str(tmp_path / "out.npy")
    _run(_extras_worker, "extras", path)
    got = np.frombuffer(np.load(path).tobytes(), dtype=abi.PARTICLE_DTYPE)
    got_col = np.frombuffer(np.load(path + ".col.npy").tobytes(), dtype=abi.COLLISION_DTYPE)
    flags = np.load(path + ".flags.npy")
    p, cfg, steps, sub, nt, v = _extras_setup()
    orc = checkers.oracle()
    q, c1, _ = orc.steps(cfg, p, steps)
    q = orc.apply_jerk(c1, q, v)
    assert checkers.bits_equal(got, q)
    status = orc.exit_check(c1, q, 30.0, 1.5)
    assert (3 if flags[1] else (4 if flags[0] else 0)) == status and status != 0
    c1.collision = abi.COLLISION_DIRECT
    want_col = orc.collision_search_subset(c1, q, sub, nt)
    assert len(want_col) > 0
    assert checkers.collisions_equal(got_col, want_col, with_ri=False)
