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


def make_case(case):
    if case == "plummer_basic":
        return ics.plummer(3001, seed=3), ics.plummer_config(3001), 3
    if case == "plummer_comp":
        return ics.plummer(2048, seed=4), ics.plummer_config(2048, gravity=abi.GRAVITY_COMPENSATED), 2
    if case == "testp_type1":
        q = ics.planetesimal_disk(2000, seed=5)
        q["m"][10:] = 1e-9
        return q, ics.planetesimal_config(testparticle_type=1), 2
    if case == "disc_tree":
        return ics.selfgravity_disc(3000, seed=6), ics.selfgravity_disc_config(boundary=abi.BOUNDARY_NONE), 2
    raise ValueError(case)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case", ["plummer_basic", "plummer_comp", "testp_type1", "disc_tree"])
def test_two_gpu_sharded_steps_bitwise(case, tmp_path):
    import torch.multiprocessing as mp

    world = 2
    path = str(tmp_path / "out.npy")
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, path)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    got = np.frombuffer(np.load(path).tobytes(), dtype=abi.PARTICLE_DTYPE)
    p, cfg, steps = make_case(case)
    want, _, _ = checkers.oracle().steps(cfg, p, steps)
    assert checkers.bits_equal(got, want)
