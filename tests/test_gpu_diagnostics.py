"""Device-side diagnostics (rebcu_energy / rebcu_com / rebcu_angular_momentum) against
the oracle's restatement of reb_simulation_energy / _com / _angular_momentum (src/tools.c:108-174, 376-408).

The reference sums each quantity into ONE scalar in index order; a parallel reduction cannot reproduce that
rounding sequence, so these are tolerance tests: 1e-12 relative to the magnitude of the terms summed
(the reference's own sequential sum carries a rounding error of that order at these N)."""
import numpy as np
import pytest

import checkers
from rebound_b200 import abi, ics
from rebound_b200.simulation import Engine, Simulation

pytestmark = pytest.mark.gpu

RTOL = 1e-12


@pytest.fixture(scope="module")
def eng():
    e = Engine(0)
    yield e
    e.close()


def cases():
    yield "plummer", ics.plummer_config(4096), ics.plummer(4096, seed=2)
    yield "plummer_odd", ics.plummer_config(1001), ics.plummer(1001, seed=3)
    q = ics.planetesimal_disk(5000, seed=4)
    yield "testp_type0", ics.planetesimal_config(), q
    q1 = q.copy()
    q1["m"][10:] = 1e-9
    yield "testp_type1", ics.planetesimal_config(testparticle_type=1), q1
    yield "disc", ics.selfgravity_disc_config(), ics.selfgravity_disc(3000, seed=5)
    yield "two", abi.default_config(), ics.plummer(2, seed=6)
    yield "one", abi.default_config(), ics.plummer(1, seed=7)


CASES = list(cases())


@pytest.mark.parametrize("name,cfg,p", CASES, ids=[c[0] for c in CASES])
def test_energy_com_angular_momentum(eng, name, cfg, p):
    orc = checkers.oracle()
    eng.upload(np.ascontiguousarray(p))
    ek, ep, et = eng.energy(cfg)
    want = orc.energy(cfg, p)
    scale = abs(ek) + abs(ep) + 1e-300
    assert et == ek + ep
    assert abs(et - want) <= RTOL * scale
    assert eng.energy(cfg) == (ek, ep, et)                     # deterministic run to run
    com, wcom = eng.com(), orc.com(cfg, p)
    assert com["m"] == pytest.approx(wcom["m"], rel=RTOL)
    for k, f in (("x", "x"), ("y", "y"), ("z", "z"), ("vx", "vx"), ("vy", "vy"), ("vz", "vz")):
        mag = float(np.sum(np.abs(p["m"] * p[f]))) / max(float(np.sum(p["m"])), 1e-300) + 1e-300
        assert abs(com[k] - wcom[k]) <= RTOL * mag, k
    L, wL = eng.angular_momentum(), orc.angular_momentum(cfg, p)
    lmag = float(np.sum(np.abs(p["m"]) * np.sqrt(p["x"] ** 2 + p["y"] ** 2 + p["z"] ** 2)
                        * np.sqrt(p["vx"] ** 2 + p["vy"] ** 2 + p["vz"] ** 2))) + 1e-300
    for a, b in zip(L, wL):
        assert abs(a - b) <= RTOL * lmag


def test_energy_large_n_against_blocked_host_sum(eng):
    """N = 2^16 (the largest N where the reference's O(N^2) energy loop is affordable)."""
    n = 1 << 16
    p = ics.plummer(n, seed=9)
    cfg = ics.plummer_config(n)
    eng.upload(np.ascontiguousarray(p))
    ek, ep, et = eng.energy(cfg)
    want = checkers.oracle().energy(cfg, p)
    assert abs(et - want) <= 1e-11 * (abs(ek) + abs(ep))


def test_simulation_energy_tracks_the_oracle_through_steps():
    p = ics.plummer(2048, seed=11)
    cfg = ics.plummer_config(2048)
    sim = Simulation()
    sim.G, sim.dt, sim.softening = cfg.G, cfg.dt, cfg.softening
    sim.add(p)
    e0 = sim.energy()
    sim.steps(20)
    e1 = sim.energy()                                          # no download in between
    orc = checkers.oracle()
    want, _, _ = orc.steps(cfg, p, 20)
    w0, w1 = orc.energy(cfg, p), orc.energy(cfg, want)
    assert abs(e0 - w0) <= RTOL * abs(w0) * 10
    assert abs((e1 - e0) - (w1 - w0)) <= 1e-11 * abs(w0)       # same energy error as the reference path
    sim.close()


def test_fp64_peak_probe_is_plausible(eng):
    tf = eng.measure_fp64_peak()
    assert 10.0 < tf < 80.0          # B200: ~37 TFLOP/s nominal at 1.97 GHz


def test_exit_checks_exact(eng):
    """run_heartbeat's exit conditions (simulation.c:242-272) on the device: exact predicates, incl. thresholds that sit
    on a particle / a pair, NaN coordinates, N = 0 and 1, and a size with several tiles per row."""
    from test_oracle_vs_reference import exit_cases
    for q, mx, mn in exit_cases():
        want = checkers.oracle().exit_check(abi.default_config(), q, mx, mn)
        if len(q) == 0:
            continue            # nothing to upload
        eng.upload(np.ascontiguousarray(q))
        escape, encounter = eng.exit_check(mx, mn)
        got = 3 if encounter else (4 if escape else 0)
        assert got == want, (mx, mn)
    rng = np.random.default_rng(5)
    n = 5000
    q = abi.particles(n)
    for f in ("x", "y", "z"):
        q[f] = rng.uniform(-1, 1, n)
    eng.upload(q)
    assert eng.exit_check(0.0, 1e-7) == (False, False)
    q["x"][4999] = q["x"][17] + 3e-8
    q["y"][4999] = q["y"][17]
    q["z"][4999] = q["z"][17]
    q["x"][1234], q["y"][1234], q["z"][1234] = 0.0, 0.0, 1.8
    eng.upload(q)
    assert eng.exit_check(1.75, 1e-7) == (True, True)
    assert eng.exit_check(1.81, 1e-8) == (False, False)
    assert checkers.oracle().exit_check(abi.default_config(), q, 1.75, 1e-7) == 3
