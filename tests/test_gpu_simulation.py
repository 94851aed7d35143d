"""The Python mirror of the reference's Simulation interface for the hot path (rebound_b200/simulation.py):
attribute names and string enums as in rebound/simulation.py, add / steps / integrate / synchronize, and
the reference's error texts."""
import numpy as np
import pytest

import checkers
from checkers import bits_equal
from rebound_b200 import abi, ics
from rebound_b200.simulation import ReboundCudaError, Simulation

pytestmark = pytest.mark.gpu


def test_simulation_mirror_plummer_leapfrog():
    p = ics.plummer(1500, seed=3)
    cfg = ics.plummer_config(1500)
    sim = Simulation()
    sim.integrator = "leapfrog"
    sim.gravity = "basic"
    sim.G = cfg.G
    sim.dt = cfg.dt
    sim.softening = cfg.softening
    sim.add(p)
    assert sim.N == 1500
    sim.steps(4)
    want, cw, _ = checkers.oracle().steps(cfg, p, 4)
    assert bits_equal(sim.particles, want)
    assert sim.t == cw.t and sim.steps_done == 4
    # particles modified on the host between steps (did_modify_particles protocol)
    q = sim.particles
    q["vx"] += 0.125
    sim.did_modify_particles()
    sim.steps(2)
    w2 = want.copy()
    w2["vx"] += 0.125
    want2, _, _ = checkers.oracle().steps(cw, w2, 2)
    assert bits_equal(sim.particles, want2)
    sim.close()


def test_simulation_mirror_integrate_and_enums():
    sim = Simulation()
    with pytest.raises(ValueError):
        sim.gravity = "nonsense"
    sim.gravity = "tree"
    sim.boundary = "open"
    sim.collision = "none"
    sim.integrator = "leapfrog"
    sim.root_size = 10.2
    sim.opening_angle2 = 0.25
    sim.softening = 0.02
    sim.dt = 3e-2
    p = ics.selfgravity_disc(2000, seed=5)
    sim.add(p)
    sim.integrate(0.1)            # 4 whole steps of 0.03
    cfg = ics.selfgravity_disc_config()
    want, cw, _ = checkers.oracle().steps(cfg, p, 4)
    assert sim.N == len(want)
    assert bits_equal(sim.particles, want)
    assert sim.t == cw.t
    cells = sim.tree()
    assert len(cells) == len(checkers.oracle().tree_dump(cw, checkers.oracle().boundary_check(cw, want)[0]))
    sim.close()


def test_simulation_mirror_errors_use_reference_text():
    sim = Simulation()
    sim.gravity = "tree"
    sim.integrator = "leapfrog"
    p = ics.selfgravity_disc(10, seed=5)
    sim.add(p)
    with pytest.raises(ReboundCudaError) as e:
        sim.steps(1)            # root_size not set
    assert e.value.msg == "Set root_size to a finite value to use a tree based gravity or collision solver."
    sim.root_size = 10.2
    q = sim.particles
    q["x"][3], q["y"][3], q["z"][3] = q["x"][2], q["y"][2], q["z"][2]
    sim.did_modify_particles()
    with pytest.raises(ReboundCudaError) as e:
        sim.update_acceleration()
    assert e.value.msg == "Cannot add two particles with the same coordinates to the tree."
    sim.close()


def test_simulation_mirror_collision_search_shearing_sheet():
    p = ics.shearing_sheet(root_size=30.0, seed=9)
    cfg = ics.shearing_sheet_config(root_size=30.0, t=55.5)
    sim = Simulation()
    for name, _ in abi.Config._fields_:
        setattr(sim, name, getattr(cfg, name))
    sim.add(p)
    got = sim.collision_search()
    want = checkers.oracle().collision_search(cfg, p)
    assert checkers.collisions_equal(got, want)
    sim.close()


def test_independent_simulations_in_parallel_threads():
    """SURVEY 8b threading: several simulations may be stepped from different threads (ctypes releases the GIL);
    every handle owns its stream and buffers, so concurrent runs give the bits of the serial runs."""
    import threading

    from rebound_b200.simulation import Engine

    cases = [(ics.plummer(3000, seed=1), ics.plummer_config(3000), 6),
             (ics.selfgravity_disc(4000, seed=2), ics.selfgravity_disc_config(collision=abi.COLLISION_NONE), 5),
             (ics.shearing_sheet(root_size=40.0, seed=3), ics.shearing_sheet_config(root_size=40.0), 8),
             (ics.planetesimal_disk(5000, seed=4), ics.planetesimal_config(), 12)]
    serial = []
    for p, cfg, steps in cases:
        e = Engine(0)
        q = p.copy()
        n = e.steps_host(cfg.copy(), q, steps)
        serial.append(q[:n].tobytes())
        e.close()
    results = [None] * len(cases)
    errors = []

    def work(k):
        try:
            p, cfg, steps = cases[k]
            e = Engine(0)
            out = None
            for _ in range(5):                      # repeat to give the threads time to overlap
                q = p.copy()
                n = e.steps_host(cfg.copy(), q, steps)
                out = q[:n].tobytes()
            e.close()
            results[k] = out
        except Exception as exc:                    # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(cases))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errors, errors
    assert results == serial


def test_interrupt_flag_stops_multi_step_calls():
    """rebcu_set_interrupt_flag: the reference leaves its long loops on a second Ctrl-C (reb_sigint > 1, src/gravity.c:196);
    rebcu_steps tests the caller's flag before every step, completes the step in progress and returns REBCU_INTERRUPTED."""
    import ctypes as C
    from rebound_b200.simulation import Engine
    eng = Engine(0)
    try:
        p = ics.plummer(600, seed=3)
        cfg = ics.plummer_config(600)
        flag = C.c_int(0)
        eng.set_interrupt_flag(flag)
        eng.upload(np.ascontiguousarray(p))
        c = cfg.copy()
        eng.steps(c, 3)                               # flag low: runs to the end
        t3 = c.t
        flag.value = 1                                # first Ctrl-C: the reference keeps going as well
        eng.steps(c, 1)
        assert c.t > t3
        flag.value = 2
        t4 = c.t
        with pytest.raises(KeyboardInterrupt):
            eng.steps(c, 5)
        assert c.t == t4                              # nothing was started
        want, cw, _ = checkers.oracle().steps(cfg, p, 4)
        assert checkers.bits_equal(eng.download(), want) and c.t == cw.t
        eng.set_interrupt_flag(None)
        eng.steps(c, 2)
        want, cw, _ = checkers.oracle().steps(cfg, p, 6)
        assert checkers.bits_equal(eng.download(), want)
    finally:
        eng.close()
