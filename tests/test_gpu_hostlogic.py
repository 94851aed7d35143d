"""Host logic of the drop-in on the REAL engine (B200): what tests/test_hostlogic_cpu.py runs on the CPU mock of the
engine -- the stress driver tests/c/hostlogic_driver.c (particle arrays that grow, move, shrink and are edited
between calls, integrator / gravity switches, copies, hooks, 600 live simulations, threads, the SEI cache, user ODEs,
fuzzed API sequences) and the reference's own Python package -- linked against the drop-in librebound with the CUDA
engine underneath.  Every dump must equal the unmodified reference's, bit for bit, in the three residency modes."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "rebound_b200", "_dropin")
PYREF = os.path.join(DROPIN, "pyref")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")
MODES = [("0", "host_authoritative"), ("1", "resident"), ("", "auto")]

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1200),
              pytest.mark.skipif(not os.path.exists(os.path.join(DROPIN, "hl_dropin")),
                                 reason="rebound_b200/_dropin/hl_dropin not built (needs the reference sources at build time)")]

HL_SCENARIOS = ["addremove", "switch", "copy", "error", "short", "threads", "many", "hooks", "seicache", "odes",
                # heartbeats on a particle array large enough to stay on the device and be fetched on demand (shim_lazy.c)
                "lazy_blind", "lazy_read", "lazy_write", "lazy_grow"]


def _hl(binary, scen, n, out, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([os.path.join(DROPIN, binary), scen, str(out), str(n)], capture_output=True, text=True, env=e, timeout=timeout)
    assert r.returncode == 0, (scen, r.stderr[-2000:])
    return open(out, "rb").read()


@pytest.mark.parametrize("scen", HL_SCENARIOS)
def test_host_side_call_sequences_on_the_real_engine(scen, tmp_path):
    ref = _hl("hl_ref", scen, 60, tmp_path / "ref.bin")
    assert len(ref) > 1000
    for value, name in MODES:
        if scen.startswith("lazy") and scen != "lazy_blind" and name == "resident":
            continue        # explicit residency: a heartbeat has to synchronise before it looks (WHFast's protocol)
        got = _hl("hl_dropin", scen, 60, tmp_path / f"got_{name}.bin", env={"REBOUND_B200_RESIDENT": value})
        assert got == ref, (scen, name)


def test_fuzzed_api_sequences_on_the_real_engine(tmp_path):
    """A sample of the fuzz seeds of the CPU suite (leapfrog cloud and shearing-sheet families)."""
    checked = 0
    cases = [(f"fuzz{seed}", "30") for seed in range(1, 13)] + [(f"fuzs{seed}", "120") for seed in range(1, 9)]
    for scen, n in cases:
        try:
            ref = _hl("hl_ref", scen, n, tmp_path / "ref.bin", timeout=30)
        except subprocess.TimeoutExpired:
            continue            # the reference itself does not finish on a few seeds
        for value, name in MODES:
            got = _hl("hl_dropin", scen, n, tmp_path / "got.bin", env={"REBOUND_B200_RESIDENT": value}, timeout=300)
            assert got == ref, (scen, name)
        checked += 1
    assert checked >= 15


def _python_env(tmp_path, libfile, name):
    lib_dir = tmp_path / ("lib_" + name)
    lib_dir.mkdir(parents=True, exist_ok=True)
    shutil.copy(libfile, lib_dir / "librebound.so")
    return dict(os.environ, PYTHONPATH=f"{lib_dir}:{PYREF}"), lib_dir


@pytest.mark.skipif(not os.path.isdir(os.path.join(PYREF, "rebound")), reason="reference Python package fixture not built")
def test_reference_python_package_on_the_real_engine(tmp_path):
    """`import rebound` (the reference's ctypes package, unchanged) on the drop-in with the CUDA engine: integrate()
    with its exit logic, Escape, tree gravity with an open boundary, merging collisions, a Simulationarchive restart,
    a shearing sheet -- every printed state equal to the run on the unmodified reference library."""
    from test_hostlogic_cpu import PY_SCRIPT

    outs = {}
    for name, lib, resident in (("ref", REF_LIB, None), ("auto", os.path.join(DROPIN, "librebound.so"), ""),
                                ("host", os.path.join(DROPIN, "librebound.so"), "0"), ("resident", os.path.join(DROPIN, "librebound.so"), "1")):
        env, lib_dir = _python_env(tmp_path, lib, name)
        if resident is not None:
            env["REBOUND_B200_RESIDENT"] = resident
            env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "rebound_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
        r = subprocess.run([sys.executable, "-c", PY_SCRIPT, str(tmp_path / f"{name}.sa")], capture_output=True, text=True,
                           env=env, timeout=900, cwd=str(tmp_path))
        assert r.returncode == 0, (name, r.stderr[-3000:])
        assert f"LIB {lib_dir}" in r.stdout
        outs[name] = [l for l in r.stdout.splitlines() if not l.startswith("LIB")]
    assert len(outs["ref"]) >= 10 and any(l.startswith("B escape") for l in outs["ref"])
    for name in ("auto", "host", "resident"):
        assert outs[name] == outs["ref"], name


REFERENCE_TESTS = ["test_gravity.py", "test_collisions.py", "test_shearingsheet.py", "test_boundary.py", "test_leapfrog.py",
                   "test_simulation.py", "test_eos.py", "test_mercurius.py", "test_trace.py", "test_additional_forces.py",
                   "test_post_timestep_modifications.py", "test_copy.py", "test_simulationarchive.py", "test_fpcontract.py",
                   "test_size_of_simulation.py"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(PYREF, "rebound", "tests")), reason="reference Python tests fixture not built")
def test_the_references_own_python_tests_pass_on_the_real_engine(tmp_path):
    """The reference's unit tests for the hot path and its callers (SURVEY.md 8c lists them as behavioural pins),
    unmodified, with `import rebound` resolving to the drop-in on the CUDA engine: the same tests pass as on the
    reference's own library."""
    files = [os.path.join(PYREF, "rebound", "tests", f) for f in REFERENCE_TESTS]
    files = [f for f in files if os.path.exists(f)]
    assert len(files) >= 8
    counts = {}
    for name, lib in (("ref", REF_LIB), ("dropin", os.path.join(DROPIN, "librebound.so"))):
        env, lib_dir = _python_env(tmp_path, lib, name)
        env["OMP_NUM_THREADS"] = "1"
        if name == "dropin":
            env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "rebound_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
        r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--no-header", *files],
                           capture_output=True, text=True, env=env, timeout=1100, cwd=str(tmp_path))
        tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:]
        assert r.returncode == 0, (name, r.stdout[-3000:])
        counts[name] = tail.split(" in ")[0]
    assert "passed" in counts["ref"] and counts["ref"] == counts["dropin"]
