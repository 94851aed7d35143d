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
SCENARIOS = [("plummer", 300, 4), ("plummer_comp", 200, 3), ("testparticles", 300, 4), ("disc", 600, 4), ("sheet", 25, 25), ("sheet_hb", 25, 10),
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


def test_long_runs_are_cut_into_pieces(mock_driver, tmp_path):
    """More than 4096 steps in one call: reb_simulation_steps and reb_simulation_integrate hand them over in pieces
    (shim_steps.c); the results stay bit-identical and every piece is one engine call."""
    ref = run(os.path.join(DROPIN, "driver_ref"), "testparticles", 40, 9000, tmp_path / "ref.bin")
    got = run(mock_driver, "testparticles", 40, 9000, tmp_path / "mock.bin", env={"REBOUND_B200_RESIDENT": ""})
    assert np.array_equal(ref, got)
    s = stats(mock_driver, "testparticles", 40, 9000, tmp_path, REBOUND_B200_RESIDENT="")
    assert s == {"uploads": 3, "downloads": 3, "step_calls": 3, "steps": 9000}
    ref = run(os.path.join(DROPIN, "driver_ref"), "integ_over", 30, 12000, tmp_path / "ref2.bin")
    got = run(mock_driver, "integ_over", 30, 12000, tmp_path / "mock2.bin", env={"REBOUND_B200_RESIDENT": ""})
    assert np.array_equal(ref, got)
    s = stats(mock_driver, "integ_over", 30, 12000, tmp_path, REBOUND_B200_RESIDENT="")
    assert s["steps"] >= 12000 and s["step_calls"] <= 12


def test_observers_keep_the_simulation_host_current(mock_driver, tmp_path):
    """plummer installs a heartbeat: in automatic mode every step ends with the particles on the host."""
    steps = 5
    s = stats(mock_driver, "plummer", 200, steps, tmp_path, REBOUND_B200_RESIDENT="")
    assert s["downloads"] >= steps and s["step_calls"] == steps
    # exit distances without boundary / collisions: checked on the device, the simulation stays resident
    e = stats(mock_driver, "escape", 300, 300, tmp_path, REBOUND_B200_RESIDENT="")
    assert e["uploads"] == 1 and e["downloads"] == 1 and e["steps"] > 10


PY_SCRIPT = r"""
import math, random, sys
import rebound
print("LIB", rebound.__libpath__)

def cloud(sim, n, seed, vel=0.3, radius=0.0):
    rng = random.Random(seed)
    for _ in range(n):
        sim.add(m=1.0 / n, x=rng.uniform(-1, 1), y=rng.uniform(-1, 1), z=rng.uniform(-1, 1),
                vx=rng.gauss(0, vel), vy=rng.gauss(0, vel), vz=rng.gauss(0, vel), r=radius)

def dump(tag, sim):
    import hashlib
    h = hashlib.sha256()
    for p in sim.particles:
        for v in (p.x, p.y, p.z, p.vx, p.vy, p.vz, p.ax, p.ay, p.az):
            h.update(v.hex().encode())
    print(tag, sim.N, sim.t.hex(), sim.dt.hex(), sim.steps_done, h.hexdigest()[:24])

# 1. integrate() to a time that is not a whole number of steps away, forwards then backwards
sim = rebound.Simulation(); sim.integrator = "leapfrog"; sim.softening = 0.05; sim.dt = 0.01
cloud(sim, 120, 1)
sim.integrate(0.4567); dump("A1", sim)
sim.integrate(0.9); dump("A2", sim)
sim.integrate(0.33); dump("A3", sim)
sim.exact_finish_time = 0
sim.integrate(0.71); dump("A4", sim)

# 2. exit_max_distance raises Escape from integrate()
sim = rebound.Simulation(); sim.integrator = "leapfrog"; sim.softening = 0.05; sim.dt = 0.02
cloud(sim, 150, 2, vel=1.0)
sim.exit_max_distance = 3.0
try:
    sim.integrate(50.0)
    print("B no escape")
except rebound.Escape as e:
    dump("B escape", sim)

# 4. tree gravity in an open box, steps()
sim = rebound.Simulation(); sim.integrator = "leapfrog"; sim.gravity = "tree"; sim.boundary = "open"
sim.root_size = 6.0; sim.opening_angle2 = 0.3; sim.softening = 0.05; sim.dt = 0.05
cloud(sim, 300, 4, vel=1.5)
sim.steps(40); dump("D", sim)

# 5. direct collisions with merging
sim = rebound.Simulation(); sim.integrator = "leapfrog"; sim.softening = 0.01; sim.dt = 0.02
sim.collision = "direct"; sim.collision_resolve = "merge"; sim.rand_seed = 42     # the seed of the collision shuffle defaults to time + pid
cloud(sim, 150, 5, vel=0.05, radius=0.06)
sim.integrate(1.0); dump("E", sim)

# 6. Simulationarchive snapshot and restart
sim = rebound.Simulation(); sim.integrator = "leapfrog"; sim.softening = 0.05; sim.dt = 0.01
cloud(sim, 80, 6)
sim.integrate(0.2)
sim.save_to_file(sys.argv[1], delete_file=True)
sim.integrate(0.5); dump("F1", sim)
sim2 = rebound.Simulation(sys.argv[1])
sim2.integrate(0.5); dump("F2", sim2)

# 7. shearing sheet: SEI, tree gravity, tree collisions, hard spheres, ghost boxes
sim = rebound.Simulation(); sim.integrator = "sei"; sim.gravity = "tree"; sim.collision = "tree"; sim.boundary = "shear"
sim.collision_resolve = "hardsphere"; sim.opening_angle2 = 0.5; sim.rand_seed = 42
sim.OMEGA = 0.00013143527; sim.G = 6.67428e-11; sim.softening = 0.1; sim.dt = 1e-3 * 2 * math.pi / sim.OMEGA
sim.root_size = 30.0; sim.N_root_x = 2; sim.N_root_y = 2; sim.N_root_z = 1; sim.N_ghost_x = 2; sim.N_ghost_y = 2; sim.N_ghost_z = 0
rng = random.Random(7)
for _ in range(250):
    x = rng.uniform(-30, 30); rad = rng.uniform(1.0, 2.0)
    sim.add(m=400.0 * 4.0 / 3.0 * math.pi * rad**3, r=rad, x=x, y=rng.uniform(-30, 30), z=rng.gauss(0, 1.0), vy=-1.5 * x * sim.OMEGA)
sim.steps(15); dump("G", sim)
"""


def _python_env(tmp_path, libfile, extra=()):
    lib_dir = tmp_path / ("lib_" + os.path.basename(os.path.dirname(libfile)))
    lib_dir.mkdir(parents=True, exist_ok=True)
    import shutil
    shutil.copy(libfile, lib_dir / "librebound.so")
    for f in extra:
        shutil.copy(f, lib_dir / os.path.basename(f))
    return dict(os.environ, PYTHONPATH=f"{lib_dir}:/root/reference"), lib_dir


@pytest.mark.skipif(not os.path.isdir("/root/reference/rebound"), reason="needs the reference's Python package")
def test_reference_python_package_on_the_mock_dropin_matches_the_reference_library(mock_driver, tmp_path):
    """The reference's own Python package (ctypes) driving (a) the unmodified reference library and (b) the drop-in with
    the mock engine: integrate() with its exit logic, Escape, tree gravity with an open boundary,
    merging collisions, a Simulationarchive restart, a shearing sheet.  Every printed state must be identical."""
    import sys
    outs = {}
    ref_lib = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")
    mock = (os.path.join(BUILD, "librebound.so"), (os.path.join(BUILD, "librebound_b200.so"),))
    for name, lib, extra in (("ref", ref_lib, ()), ("mock", *mock), ("mock0", *mock), ("mock1", *mock)):
        env, lib_dir = _python_env(tmp_path, lib, extra)
        if name.startswith("mock"):
            env["REBOUND_B200_RESIDENT"] = name[4:]
            env["LD_LIBRARY_PATH"] = f"{lib_dir}:{os.path.join(ROOT, 'oracle')}:" + env.get("LD_LIBRARY_PATH", "")
        r = subprocess.run([sys.executable, "-c", PY_SCRIPT, str(tmp_path / f"{name}.sa")], capture_output=True, text=True,
                           env=env, timeout=600, cwd=str(tmp_path))
        assert r.returncode == 0, r.stderr[-3000:]
        lines = [l for l in r.stdout.splitlines() if not l.startswith("LIB")]
        assert f"LIB {lib_dir}" in r.stdout
        outs[name] = lines
    assert len(outs["ref"]) >= 10
    assert any(l.startswith("B escape") for l in outs["ref"])
    for name in ("mock", "mock0", "mock1"):          # automatic, host-authoritative, resident
        assert outs["ref"] == outs[name], name


HL_SCENARIOS = ["addremove", "switch", "copy", "error", "short", "threads", "many", "hooks", "seicache", "odes"]


@pytest.mark.parametrize("scen", HL_SCENARIOS)
def test_host_side_call_sequences_match_the_reference_bitwise(mock_driver, scen, tmp_path):
    """tests/c/hostlogic_driver.c: particle arrays that grow (and move), shrink and are edited between calls, integrator
    and gravity switches, copies and diffs, an error that ends an integration, integrations shorter than a step or to the
    current time or backwards, several simulations interleaved and in threads, a sweep of 700 short-lived simulations and
    600 simulations alive at once (the shim's side table grows and is released with reb_simulation_free), host callbacks
    (additional_forces, pre/post_timestep_modifications, a collision_resolve that removes particles), an open boundary
    with track_energy_offset, MEGNO with variational particles -- on the
    unmodified reference and on the
    drop-in with the mock engine in the three residency modes."""
    ref_out = tmp_path / "ref.bin"
    r = subprocess.run([os.path.join(BUILD, "hl_ref"), scen, str(ref_out), "60"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    ref = ref_out.read_bytes()
    assert len(ref) > 1000
    for value, name in MODES:
        out = tmp_path / f"mock_{name}.bin"
        e = dict(os.environ, REBOUND_B200_RESIDENT=value)
        r = subprocess.run([os.path.join(BUILD, "hl_mock"), scen, str(out), "60"], capture_output=True, text=True, env=e, timeout=300)
        assert r.returncode == 0, r.stderr
        assert out.read_bytes() == ref, (scen, name)


def test_quadrupole_dropin_on_the_mock_engine(mock_driver, tmp_path):
    """A -DQUADRUPOLE build of the reference sources and the shim (rebound_b200/shim/Makefile QUADRUPOLE=1): the shim
    tells the engine to carry the quadrupole tensors (rebcu_config.quadrupole); tree-gravity scenarios must match the
    reference compiled the same way (oracle/_ref/libref_harness_quad.so) and differ from the monopole build."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_harness_quad.so")):
        pytest.skip("oracle/_ref/libref_harness_quad.so not built")
    r = subprocess.run(["make", "-C", HOSTLOGIC, "quad"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    quad = os.path.join(BUILD, "quad")
    for scen, n, steps in (("disc", 600, 4), ("sheet", 25, 20), ("periodic", 300, 4)):
        ref = run(os.path.join(quad, "driver_ref"), scen, n, steps, tmp_path / "refq.bin")
        mono = run(os.path.join(DROPIN, "driver_ref"), scen, n, steps, tmp_path / "ref0.bin")
        assert not np.array_equal(ref, mono), scen
        for value, name in MODES:
            got = run(os.path.join(quad, "driver_mock"), scen, n, steps, tmp_path / f"mockq_{name}.bin", env={"REBOUND_B200_RESIDENT": value})
            assert np.array_equal(ref, got), (scen, name)


# the reference's own Python tests for the hot path and its callers (SURVEY.md section 8c lists them as behavioural pins)
REFERENCE_TESTS = ["test_gravity.py", "test_collisions.py", "test_shearingsheet.py", "test_boundary.py", "test_leapfrog.py",
                   "test_simulation.py", "test_eos.py", "test_mercurius.py", "test_trace.py", "test_additional_forces.py",
                   "test_post_timestep_modifications.py", "test_copy.py", "test_simulationarchive.py",
                   "test_fpcontract.py", "test_size_of_simulation.py", "test_whfast.py"]
# explicit resident mode: the files that drive leapfrog / SEI, where residency changes what the shim does
REFERENCE_TESTS_RESIDENT = ["test_gravity.py", "test_collisions.py", "test_shearingsheet.py", "test_boundary.py", "test_leapfrog.py",
                            "test_simulation.py", "test_post_timestep_modifications.py", "test_additional_forces.py"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/rebound/tests"), reason="needs the reference's Python tests")
@pytest.mark.parametrize("resident", ["", "1"], ids=["auto", "resident"])
def test_the_references_own_python_tests_pass_on_the_mock_dropin(mock_driver, resident, tmp_path):
    """The reference's unit tests, unmodified, with `import rebound` resolving to the drop-in on the mock engine: what
    passes on the reference's own library must pass here (same counts), in automatic and in explicit resident mode."""
    import sys
    files = [os.path.join("/root/reference/rebound/tests", f) for f in (REFERENCE_TESTS_RESIDENT if resident else REFERENCE_TESTS)]
    files = [f for f in files if os.path.exists(f)]
    assert len(files) >= 8
    counts = {}
    ref_lib = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")
    mock = (os.path.join(BUILD, "librebound.so"), (os.path.join(BUILD, "librebound_b200.so"),))
    for name, lib, extra in (("ref", ref_lib, ()), ("mock", *mock)):
        env, lib_dir = _python_env(tmp_path, lib, extra)
        env["OMP_NUM_THREADS"] = "1"          # tiny systems, thousands of calls: the oracle's OpenMP teams only cost time
        if name == "mock":
            env["REBOUND_B200_RESIDENT"] = resident
            env["LD_LIBRARY_PATH"] = f"{lib_dir}:{os.path.join(ROOT, 'oracle')}:" + env.get("LD_LIBRARY_PATH", "")
        r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--no-header", "-x", *files],
                           capture_output=True, text=True, env=env, timeout=1500, cwd=str(tmp_path))
        tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:]
        assert r.returncode == 0, (name, r.stdout[-3000:])
        counts[name] = tail.split(" in ")[0]
    assert "passed" in counts["ref"] and counts["ref"] == counts["mock"]


def test_fuzzed_api_sequences_match_the_reference_bitwise(mock_driver, tmp_path):
    """Pseudo-random sequences of 60 public-API calls each (tests/c/hostlogic_driver.c, scenario fuzz<seed>): steps and
    integrations in both directions, particles added / removed / edited, gravity, integrator, collision, boundary,
    heartbeat, exit-distance and test-particle switches, copies, diagnostics.  The drop-in on the mock engine must
    leave the same trace as the unmodified reference in every residency mode.  (Seeds on which the reference itself
    does not finish within the time limit are skipped: 4 of the first 120.)"""
    checked = 0
    cases = [(f"fuzz{seed}", "30") for seed in range(1, 41)]
    # the shearing-sheet family: SEI (cache refreshed by dt changes, kept across OMEGA changes), shear boundary, ghost
    # rings, tree / direct collision searches with the reference's hard-sphere resolver
    cases += [(f"fuzs{seed}", "120") for seed in range(1, 21)]
    for scen, n in cases:
        ref_out = tmp_path / "ref.bin"
        try:
            r = subprocess.run([os.path.join(BUILD, "hl_ref"), scen, str(ref_out), n], capture_output=True, timeout=30)
        except subprocess.TimeoutExpired:
            continue
        assert r.returncode == 0
        ref = ref_out.read_bytes()
        for value, name in MODES:
            out = tmp_path / "mock.bin"
            e = dict(os.environ, REBOUND_B200_RESIDENT=value)
            r = subprocess.run([os.path.join(BUILD, "hl_mock"), scen, str(out), n], capture_output=True, env=e, timeout=120)
            assert r.returncode == 0, (scen, name, r.stderr[-500:])
            assert out.read_bytes() == ref, (scen, name)
        checked += 1
    assert checked >= 55


@pytest.mark.parametrize("times", [1, 2])
def test_ctrl_c_during_a_device_batch_ends_the_integration(mock_driver, times, tmp_path):
    """A SIGINT that arrives while a device batch of reb_simulation_integrate runs (once: the batch piece finishes;
    twice: the engine stops between two steps and returns REBCU_INTERRUPTED) ends the run with REB_STATUS_SIGINT,
    synchronised and short of tmax -- it must not fall through to the reference's loop, which clears reb_sigint."""
    hl = os.path.join(BUILD, "hl_mock")
    out = tmp_path / "sig.bin"
    e = dict(os.environ)
    e.update({"REBOUND_B200_RESIDENT": "", "MOCK_SIGINT_AT_STEP": f"300,{times}", "MOCK_ENGINE_STATS": str(tmp_path / "st.json")})
    r = subprocess.run([hl, "sigint", str(out), "40"], capture_output=True, text=True, env=e, timeout=300)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(out, dtype=np.uint8)
    info = raw[:32].view(np.float64)
    REB_STATUS_SIGINT = 6          # src/rebound.h:217-233
    assert info[0] == REB_STATUS_SIGINT and info[1] == REB_STATUS_SIGINT
    st = json.loads((tmp_path / "st.json").read_text())
    if times == 1:
        assert st["steps"] == 4096 and info[3] == 4096           # the piece in flight is completed, nothing after it
    else:
        assert st["steps"] == 300 and info[3] == 300              # stopped between two steps
    assert abs(info[2] - info[3] * 0.01) < 1e-9 and info[2] < 500.0
    hdr = raw[32 + 4:32 + 4 + 40].view(np.float64)
    assert hdr[0] == 40 and hdr[1] == info[2]


@pytest.mark.parametrize("scen", ["lazy_blind", "lazy_read", "lazy_write", "lazy_grow"])
def test_heartbeat_keeps_the_particles_on_the_device_until_somebody_looks(mock_driver, scen, tmp_path):
    """shim_lazy.c: with a heartbeat installed the automatic mode stays resident and page-protects r->particles; the host
    copy is fetched when (and only when) the heartbeat -- or anything else -- touches it.  Results: the reference's bits
    in every mode; transfers: a heartbeat that never looks costs one upload and one download per call, a reader one
    download per step it reads in, a writer additionally one upload per edit."""
    ref_out = tmp_path / "ref.bin"
    r = subprocess.run([os.path.join(BUILD, "hl_ref"), scen, str(ref_out), "3000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    ref = ref_out.read_bytes()
    counts = {}
    for value, name, lazy in (("0", "host", "1"), ("1", "resident", "1"), ("", "auto", "1"), ("", "auto_nolazy", "0")):
        if name == "resident" and scen != "lazy_blind":
            continue        # explicit residency is WHFast's protocol: a heartbeat must call reb_simulation_synchronize before it looks
        out = tmp_path / f"mock_{name}.bin"
        st = tmp_path / f"st_{name}.json"
        e = dict(os.environ, REBOUND_B200_RESIDENT=value, REBOUND_B200_LAZY=lazy, MOCK_ENGINE_STATS=str(st))
        r = subprocess.run([os.path.join(BUILD, "hl_mock"), scen, str(out), "3000"], capture_output=True, text=True, env=e, timeout=600)
        assert r.returncode == 0, (name, r.stderr[-2000:])
        assert out.read_bytes() == ref, (scen, name)
        counts[name] = json.loads(st.read_text())
    a, b = counts["auto"], counts["auto_nolazy"]
    assert b["downloads"] >= 20 and b["uploads"] >= 20           # without the mechanism: both ways every step
    if scen == "lazy_blind":
        # three calls (steps, integrate incl. its shortened last step, steps): a handful of transfers in total
        assert a["uploads"] <= 5 and a["downloads"] <= 5
    elif scen == "lazy_read":
        assert a["uploads"] <= 5 and 4 <= a["downloads"] <= 10     # one download per read step
    elif scen == "lazy_write":
        assert a["uploads"] <= 9 and a["downloads"] <= 10           # + one upload per edit
    else:
        assert a["uploads"] <= 9 and a["downloads"] <= 10
