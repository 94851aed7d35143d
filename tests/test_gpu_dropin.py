"""Drop-in boundary, end to end: the same C program (tests/c/dropin_driver.c, reference public API only)
linked against the unmodified reference and against the drop-in librebound (reference sources + CUDA hot
path behind the reference's own symbol names) must produce bit-identical final states."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "rebound_b200", "_dropin")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(DROPIN, "driver_dropin")),
                                 reason="rebound_b200/_dropin not built (needs the reference sources at build time)")]

SCENARIOS = [("plummer", 3000, 5), ("plummer_comp", 1500, 3), ("testparticles", 2000, 4), ("disc", 5000, 4), ("sheet", 40, 8), ("sheet", 25, 400),
             # a sheet large enough (N ~ 2300, array > 256 KB) for the lazy host copy under the example's heartbeat, with the
             # exact device-side resolve (shim_lazy.c, rebcu_collision_resolve_pairs)
             ("sheet_hb", 125, 12), ("sheet", 125, 12),
             ("lf4", 700, 3), ("lf6", 700, 3), ("lf8", 700, 2), ("tp0", 3000, 6), ("merge", 400, 30), ("line", 400, 30),
             ("periodic", 1500, 6), ("open_direct", 1200, 12), ("ias15", 300, 3), ("ias15_comp", 300, 3), ("whfast", 300, 10),
             # r->map / r->N_targets collision subsets of the hybrid integrators; exit conditions of run_heartbeat
             ("mercurius", 40, 200), ("trace", 40, 200), ("escape", 500, 400), ("encounter", 500, 400),
             # EOS with a modified-kick scheme: force evaluation + jerk kick per interaction step
             ("eos", 200, 6),
             # reb_simulation_integrate with its exit logic around device batches; long reb_simulation_steps runs in pieces
             ("integ_exact", 800, 40), ("integ_over", 800, 40), ("integ_back", 800, 40), ("integ_tree", 3000, 80),
             ("testparticles", 1500, 4200)]


def run(binary, scen, n, steps, tmp_path, env=None):
    out = tmp_path / f"{os.path.basename(binary)}_{scen}.bin"
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([os.path.join(DROPIN, binary), scen, str(out), str(n), str(steps)], capture_output=True, text=True, env=e, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Error!" not in r.stderr, r.stderr
    return np.fromfile(out, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("scen,n,steps", SCENARIOS, ids=[f"{s[0]}-{s[1]}-{s[2]}" for s in SCENARIOS])
@pytest.mark.parametrize("resident", ["0", "1", ""], ids=["host_authoritative", "resident", "auto"])
def test_dropin_matches_reference_bitwise(scen, n, steps, resident, tmp_path):
    if steps > 1000 and resident != "":
        pytest.skip("long runs exercise the device batches of the automatic mode only")
    ref = run("driver_ref", scen, n, steps, tmp_path)
    got = run("driver_dropin", scen, n, steps, tmp_path, env={"REBOUND_B200_RESIDENT": resident})
    assert len(ref) == len(got)
    assert np.array_equal(ref, got)
    hdr = ref[:6].view(np.float64)
    if scen in ("mercurius", "trace"):
        assert hdr[0] < n + 1                      # particles merged during close encounters
    if scen == "integ_tree":
        assert hdr[0] < n + 1                      # particles left the open box during the batched steps
    if scen == "escape":
        assert hdr[4] == 4 and hdr[1] < steps * 2e-2      # REB_STATUS_ESCAPE before tmax
    if scen == "encounter":
        assert hdr[4] == 3 and hdr[1] < steps * 2e-2      # REB_STATUS_ENCOUNTER before tmax


@pytest.mark.parametrize("resident", ["0", "1", ""], ids=["host_authoritative", "resident", "auto"])
def test_dropin_archive_restart_and_midrun_diagnostics(resident, tmp_path):
    """Simulationarchive snapshots written from inside reb_simulation_steps, and energy / COM /
    angular momentum read from a heartbeat, see the current particles also while a resident simulation is
    unsynchronised; restarting from a snapshot lands on the same bits as the uninterrupted run."""
    n, steps = 800, 7
    ref = run("driver_ref", "archive", n, steps, tmp_path)
    got = run("driver_dropin", "archive", n, steps, tmp_path, env={"REBOUND_B200_RESIDENT": resident})
    assert len(ref) == len(got) == 2 * (6 + 11 * n)
    assert np.array_equal(ref, got)
    direct, restart = got[6:6 + 11 * n], got[12 + 11 * n:]
    assert np.array_equal(direct, restart)
    tail = got[6 + 11 * n:12 + 11 * n].view(np.float64)
    assert tail[0] == n and tail[2] == 2.0 and tail[3] != 0.0       # restarted from the snapshot after 2 steps


@pytest.mark.parametrize("resident", ["0", ""], ids=["host_authoritative", "auto"])
def test_dropin_host_edits_between_calls_need_no_flag(resident, tmp_path):
    """Particles edited in r->particles between two reb_simulation_steps calls without setting
    r->did_modify_particles (legal with the reference's leapfrog, which keeps no state) are picked up: automatic
    residency drops the device copy at the synchronize that ends every call."""
    ref = run("driver_ref", "edit", 900, 4, tmp_path)
    got = run("driver_dropin", "edit", 900, 4, tmp_path, env={"REBOUND_B200_RESIDENT": resident})
    assert np.array_equal(ref, got)
