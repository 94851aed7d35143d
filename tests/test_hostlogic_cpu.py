"""Host logic of the drop-in shim on the CPU: the product's shim objects linked against a MOCK of the engine that
delegates to the oracle (tests/hostlogic/mock_engine.c -- test infrastructure, never shipped).  The mock drop-in must
reproduce the reference driver bit for bit in every scenario and residency mode, and the mock's counters show that
the residency modes and the device batches behind reb_simulation_steps / reb_simulation_integrate do what
INTEGRATION.md says (one upload per call, not one per step)."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "rebound_b200", "_dropin")
HOSTLOGIC = os.path.join(ROOT, "tests", "hostlogic")
BUILD = os.path.join(HOSTLOGIC, "_build")

pytestmark = [pytest.mark.needs_ref,
              pytest.mark.skipif(not os.path.isdir(os.path.join(DROPIN, "obj")) or not os.path.isdir("/root/reference/src"),
                                 reason="needs rebound_b200/_dropin/obj and the reference headers (authoring container)")]

# (scenario, N, steps): tests/c/dropin_driver.c, sized for the CPU oracle
SCENARIOS = [("plummer", 300, 4), ("plummer_comp", 200, 3), ("testparticles", 300, 4), ("disc", 600, 4), ("sheet", 25, 25),
             ("lf4", 150, 3), ("lf8", 150, 2), ("tp0", 400, 6), ("merge", 200, 20), ("line", 200, 20), ("periodic", 300, 4),
             ("open_direct", 300, 10), ("ias15", 60, 3), ("ias15_comp", 60, 3), ("whfast", 60, 8), ("mercurius", 30, 120),
             ("trace", 30, 120), ("escape", 300, 300), ("encounter", 300, 300), ("eos", 60, 5), ("edit", 200, 3),
             ("integ_exact", 150, 40), ("integ_over", 150, 40), ("integ_back", 150, 40), ("integ_tree", 500, 60),
             ("archive", 200, 7)]
MODES = [("0", "host_authoritative"), ("1", "resident"), ("", "auto")]


@pytest.fixture(scope="module")
def mock_driver():
    r = subprocess.run(["make", "-C", HOSTLOGIC], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return os.path.join(BUILD, "driver_mock")


def run(binary, scen, n, steps, out, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([binary, scen, str(out), str(n), str(steps)], capture_output=True, text=True, env=e, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Error!" not in r.stderr, r.stderr
    return np.fromfile(out, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("scen,n,steps", SCENARIOS, ids=[f"{s[0]}-{s[1]}-{s[2]}" for s in SCENARIOS])
def test_mock_dropin_matches_reference_bitwise(mock_driver, scen, n, steps, tmp_path):
    ref = run(os.path.join(DROPIN, "driver_ref"), scen, n, steps, tmp_path / "ref.bin")
    for value, name in MODES:
        if scen == "edit" and name == "resident":
            continue        # explicit residency: host edits must be flagged with r->did_modify_particles
        got = run(mock_driver, scen, n, steps, tmp_path / f"mock_{name}.bin", env={"REBOUND_B200_RESIDENT": value})
        assert len(ref) == len(got), (scen, name)
        assert np.array_equal(ref, got), (scen, name)


def stats(mock_driver, scen, n, steps, tmp_path, **env):
    path = tmp_path / "stats.json"
    e = {"MOCK_ENGINE_STATS": str(path)}
    e.update(env)
    run(mock_driver, scen, n, steps, tmp_path / "out.bin", env=e)
    return json.loads(path.read_text())


def test_residency_modes_move_the_particles_as_documented(mock_driver, tmp_path):
    """tp0 = leapfrog, nothing observes the particles between steps."""
    n, steps = 300, 12
    host = stats(mock_driver, "tp0", n, steps, tmp_path, REBOUND_B200_RESIDENT="0")
    assert host["uploads"] >= steps and host["downloads"] >= steps          # every replaced call goes both ways
    res = stats(mock_driver, "tp0", n, steps, tmp_path, REBOUND_B200_RESIDENT="1")
    assert res["uploads"] == 1 and res["downloads"] == 1 and res["steps"] == steps and res["step_calls"] == steps
    auto = stats(mock_driver, "tp0", n, steps, tmp_path, REBOUND_B200_RESIDENT="")
    assert auto == {"uploads": 1, "downloads": 1, "step_calls": 1, "steps": steps}      # one device batch
    nobatch = stats(mock_driver, "tp0", n, steps, tmp_path, REBOUND_B200_RESIDENT="", REBOUND_B200_BATCH="0")
    assert nobatch["uploads"] == 1 and nobatch["downloads"] == 1 and nobatch["step_calls"] == steps


def test_integrate_hands_whole_steps_to_one_batch_and_keeps_the_exit_logic(mock_driver, tmp_path):
    """integ_exact integrates twice to times that are not a whole number of steps away (exact_finish_time = 1): each
    call is one batch (all whole steps but the last two) plus the reference's own loop for the end of the run."""
    s = stats(mock_driver, "integ_exact", 150, 40, tmp_path, REBOUND_B200_RESIDENT="")
    # 2 calls x (1 batch + a few single steps of the reference loop, incl. the shortened last one)
    assert 2 <= s["step_calls"] - 2 <= 8
    assert s["steps"] >= 40
    # per call: the batch; the two whole steps of the tail; the shortened last step (reb_check_exit synchronises before
    # it changes dt, simulation.c:325-327, which ends the residency of the tail)
    assert s["uploads"] == 6 and s["downloads"] == 6
    off = stats(mock_driver, "integ_exact", 150, 40, tmp_path, REBOUND_B200_RESIDENT="", REBOUND_B200_BATCH="0")
    assert off["step_calls"] == off["steps"] >= 40


def test_observers_keep_the_simulation_host_current(mock_driver, tmp_path):
    """plummer installs a heartbeat: in automatic mode every step ends with the particles on the host."""
    steps = 5
    s = stats(mock_driver, "plummer", 200, steps, tmp_path, REBOUND_B200_RESIDENT="")
    assert s["downloads"] >= steps and s["step_calls"] == steps
    # exit distances without boundary / collisions: checked on the device, the simulation stays resident
    e = stats(mock_driver, "escape", 300, 300, tmp_path, REBOUND_B200_RESIDENT="")
    assert e["uploads"] == 1 and e["downloads"] == 1 and e["steps"] > 10
