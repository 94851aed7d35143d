"""Several GPUs behind ONE handle (rebcu_create_group, csrc/group.cu) and, through it, behind the reference's C API:
a single-threaded program calling reb_simulation_integrate() shards over the devices named by REBOUND_B200_DEVICES.

On a one-GPU box the group lists device 0 several times (the ranks then exchange through the LOCAL transport); with
>= 2 GPUs the same tests also run over NCCL on distinct devices.  Everything is compared with the oracle / the
unmodified reference bit for bit: STRICT results do not depend on the sharding."""
import os
import subprocess

import numpy as np
import pytest
import torch

import checkers
from rebound_b200 import abi, ics
from rebound_b200.simulation import Engine, ReboundCudaError
from test_gpu_multi import make_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "rebound_b200", "_dropin")


def device_lists():
    out = [[0, 0], [0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        out.append(list(range(min(torch.cuda.device_count(), 8))))
    return out


@pytest.mark.parametrize("devices", device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
@pytest.mark.parametrize("case", ["plummer_basic", "plummer_comp", "testp_type1", "disc_tree", "open_basic", "open_tree"])
def test_group_handle_steps_host_bitwise(case, devices):
    """rebcu_steps_host on the leader: every rank uploads its block of the host array, steps, downloads its block."""
    p, cfg, steps = make_case(case)
    want, cw, _ = checkers.oracle().steps(cfg, p, steps)
    eng = Engine(devices=devices)
    try:
        assert eng.f["group_size"](eng.h) == len(devices)
        eng.set_sharded_build(1)
        q, c = p.copy(), cfg.copy()
        n = eng.steps_host(c, q, steps)
        assert n == len(want) and c.t == cw.t and c.N_active == cw.N_active
        assert checkers.bits_equal(q[:n], want)
        assert np.array_equal(q["name"][:n], want["name"])
    finally:
        eng.close()


@pytest.mark.parametrize("devices", device_lists()[:1] + device_lists()[2:], ids=lambda d: "dev" + "".join(map(str, d)))
def test_group_handle_call_by_call(devices):
    """The calls the drop-in makes one by one: upload, integrator step, boundary check, collision search (merged list),
    download -- on a shearing sheet (SEI, tree gravity with ghost boxes, tree collision search)."""
    p = ics.shearing_sheet(root_size=40.0, seed=5)
    cfg = ics.shearing_sheet_config(root_size=40.0, t=55.5)
    orc = checkers.oracle()
    eng = Engine(devices=devices)
    try:
        eng.upload(np.ascontiguousarray(p))
        c = cfg.copy()
        q, cw = p.copy(), cfg.copy()
        for _ in range(3):
            eng.integrator_step(c)
            eng.boundary_check(c)
            q, cw = orc.integrator_step(cw, q)
            q, cw = orc.boundary_check(cw, q)
        got_col = eng.collision_search(c)
        want_col = orc.collision_search(cw, q)
        assert len(want_col) > 0
        assert checkers.collisions_equal(got_col, want_col)
        got = eng.download()
        assert checkers.bits_equal(got, q)
        with pytest.raises(ReboundCudaError):
            eng.energy(c)                      # diagnostics are not available on a group handle
    finally:
        eng.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(DROPIN, "driver_dropin")), reason="rebound_b200/_dropin not built")
@pytest.mark.parametrize("devices", ["0,0", "0,0,0"] + (["0-%d" % (min(torch.cuda.device_count(), 8) - 1)] if torch.cuda.device_count() >= 2 else []))
@pytest.mark.parametrize("scen,n,steps", [("plummer", 3000, 5), ("plummer_comp", 1500, 3), ("testparticles", 2000, 4), ("disc", 5000, 4),
                                          ("sheet", 40, 8), ("merge", 400, 30), ("open_direct", 1200, 12), ("integ_tree", 3000, 80),
                                          ("escape", 500, 400), ("mercurius", 40, 200)])
@pytest.mark.parametrize("resident", ["0", "1", ""], ids=["host_authoritative", "resident", "auto"])
def test_dropin_shards_over_devices_bitwise(scen, n, steps, devices, resident, tmp_path):
    """tests/c/dropin_driver.c (reference public API only, one thread) with REBOUND_B200_DEVICES: the drop-in creates a
    multi-GPU group; final states equal the unmodified reference's bit for bit."""
    def run(binary, env):
        out = tmp_path / f"{binary}.bin"
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([os.path.join(DROPIN, binary), scen, str(out), str(n), str(steps)], capture_output=True, text=True, env=e, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "Error!" not in r.stderr, r.stderr[-2000:]
        return np.fromfile(out, dtype=np.float64).view(np.uint64)

    ref = run("driver_ref", {})
    got = run("driver_dropin", {"REBOUND_B200_RESIDENT": resident, "REBOUND_B200_DEVICES": devices, "REBOUND_B200_SHARD_BUILD": "1"})
    assert len(ref) == len(got)
    assert np.array_equal(ref, got)
