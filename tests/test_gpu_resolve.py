"""Device-side hard-sphere resolve (rebcu_set_device_resolve, csrc/resolve.cu) against the
oracle's restatement of the reference's shuffle + sequential resolve loop (src/collision.c:336-404, 573-665).

This is the one part of the library with a DOCUMENTED RELAXATION of the parity bar: the resolver's atan2/sin/cos/pow
come from the device's libm, so each resolved collision agrees with the reference to a few ulp, not bit for bit.
What must hold exactly: which collisions are resolved (collisions_log_n), and the order semantics (a collision sees
the outcome of every earlier collision of its particles).  Tolerances are stated per test."""
import numpy as np
import pytest

import checkers
from rebound_b200 import abi, ics
from rebound_b200.simulation import Engine

pytestmark = pytest.mark.gpu

BRIDGES = (0.32, 100.0, -0.234, 0.0, 1.0)        # examples/shearing_sheet/problem.c:96-103


@pytest.fixture(scope="module")
def eng():
    e = Engine(0)
    yield e
    e.close()


def rel_state_error(a, b):
    err = 0.0
    for f in ("x", "y", "z", "vx", "vy", "vz"):
        scale = np.max(np.abs(b[f])) + 1e-300
        err = max(err, float(np.max(np.abs(a[f] - b[f]))) / scale)
    return err


@pytest.mark.parametrize("law,steps,tol", [(1, 1, 1e-13), (1, 5, 1e-11), (2, 1, 1e-13), (2, 5, 1e-11), (2, 40, 1e-7)])
def test_sheet_steps_with_device_resolve(eng, law, steps, tol):
    """Shearing sheet (C5 recipe): SEI + tree gravity + tree collision search + hard-sphere resolve, everything on the
    device, against the oracle's host loop.  law 1: elastic, law 2: Bridges et al. velocity-dependent restitution."""
    p = ics.shearing_sheet(root_size=40.0, seed=5)
    cfg = ics.shearing_sheet_config(root_size=40.0)
    min_v = 1.0 * ics.SHEET_OMEGA * 0.001
    want, cw, aux = checkers.oracle().steps(cfg, p, steps, resolve=law, minimum_collision_velocity=min_v)
    eng.upload(np.ascontiguousarray(p))
    eng.set_device_resolve(True, restitution=None if law == 1 else BRIDGES, minimum_collision_velocity=min_v, rand_seed=42)
    c = cfg.copy()
    eng.steps(c, steps)
    got = eng.download()
    st = eng.collision_stats()
    eng.set_device_resolve(False)
    assert len(got) == len(want) and c.t == cw.t
    assert aux["collisions_log_n"] > 0
    assert st["collisions_log_n"] == aux["collisions_log_n"]            # the same collisions were resolved
    assert rel_state_error(got, want) <= tol
    assert abs(st["collisions_plog"] - aux["collisions_plog"]) <= 1e-9 * abs(aux["collisions_plog"]) + 1e-300
    assert st["rounds"] >= 1


def test_chained_collisions_keep_the_sequential_semantics(eng):
    """A crowded box: most particles sit in several collisions at once, so the outcome depends on the processing
    order; the conflict-free rounds must reproduce the reference's sequential loop (needs more than one round)."""
    rng = np.random.default_rng(3)
    n = 400
    p = abi.particles(n)
    for f in ("x", "y", "z"):
        p[f] = rng.uniform(-1.0, 1.0, n)
    for f in ("vx", "vy", "vz"):
        p[f] = rng.normal(0.0, 1.0, n)
    p["m"] = rng.uniform(0.5, 2.0, n)
    p["r"] = rng.uniform(0.10, 0.22, n)
    cfg = abi.default_config(gravity=abi.GRAVITY_NONE, collision=abi.COLLISION_DIRECT, dt=1e-3)
    want, _, aux = checkers.oracle().steps(cfg, p, 1, resolve=1)
    eng.upload(np.ascontiguousarray(p))
    eng.set_device_resolve(True, restitution=None, rand_seed=42)
    eng.steps(cfg.copy(), 1)
    got = eng.download()
    st = eng.collision_stats()
    eng.set_device_resolve(False)
    assert st["rounds"] > 3                                              # long dependency chains really occur
    assert st["collisions_log_n"] == aux["collisions_log_n"] > 100
    assert rel_state_error(got, want) <= 1e-12
    # an order-insensitive implementation (every collision from the pre-collision velocities) would be far off:
    # the total momentum is conserved by both, the individual velocities are not
    assert np.allclose((got["m"] * got["vx"]).sum(), (p["m"] * p["vx"]).sum(), rtol=0, atol=1e-10)


def test_device_resolve_is_reproducible(eng):
    p = ics.shearing_sheet(root_size=40.0, seed=6)
    cfg = ics.shearing_sheet_config(root_size=40.0)
    outs = []
    for _ in range(2):
        eng.upload(np.ascontiguousarray(p))
        eng.set_device_resolve(True, restitution=BRIDGES, minimum_collision_velocity=1e-7, rand_seed=7)
        eng.steps(cfg.copy(), 10)
        outs.append((eng.download().tobytes(), eng.collision_stats()))
        eng.set_device_resolve(False)
    assert outs[0] == outs[1]


def _hardsphere_python(eps_fn, min_v):
    """reb_collision_resolve_hardsphere (src/collision.c:573-665) on one abi.ResolvePair, expression by expression, with
    Python floats (IEEE doubles, no contraction) and the C library's atan2 / sin / cos / sqrt behind the math module."""
    import math

    def resolve(q):
        x1, y1, z1, vx1, vy1, vz1, m1, r1 = list(q.s1)
        x2, y2, z2, vx2, vy2, vz2, m2, r2 = list(q.s2)
        x21 = x1 + q.gb.x - x2; y21 = y1 + q.gb.y - y2; z21 = z1 + q.gb.z - z2
        rp = r1 + r2
        oldvyouter = vy1 if x21 > 0 else vy2
        if rp * rp < x21 * x21 + y21 * y21 + z21 * z21:
            return
        vx21 = vx1 + q.gb.vx - vx2; vy21 = vy1 + q.gb.vy - vy2; vz21 = vz1 + q.gb.vz - vz2
        if vx21 * x21 + vy21 * y21 + vz21 * z21 > 0:
            return
        theta = math.atan2(z21, y21); stheta = math.sin(theta); ctheta = math.cos(theta)
        vy21n = ctheta * vy21 + stheta * vz21
        y21n = ctheta * y21 + stheta * z21
        phi = math.atan2(y21n, x21); cphi = math.cos(phi); sphi = math.sin(phi)
        vx21nn = cphi * vx21 + sphi * vy21n
        eps = eps_fn(vx21nn)
        dvx2 = -(1.0 + eps) * vx21nn
        minr = r2 if r1 > r2 else r1
        maxr = r2 if r1 < r2 else r1
        mindv = minr * min_v
        rr = math.sqrt(x21 * x21 + y21 * y21 + z21 * z21)
        mindv *= 1. - (rr - maxr) / minr
        if mindv > maxr * min_v:
            mindv = maxr * min_v
        if dvx2 < mindv:
            dvx2 = mindv
        dvx2n = cphi * dvx2; dvy2n = sphi * dvx2; dvy2nn = ctheta * dvy2n; dvz2nn = stheta * dvy2n
        p2pf = m1 / (m1 + m2)
        q.v2[0] = vx2 - p2pf * dvx2n; q.v2[1] = vy2 - p2pf * dvy2nn; q.v2[2] = vz2 - p2pf * dvz2nn
        p1pf = m2 / (m1 + m2)
        q.v1[0] = vx1 + p1pf * dvx2n; q.v1[1] = vy1 + p1pf * dvy2nn; q.v1[2] = vz1 + p1pf * dvz2nn
        q.plog_term = -abs(x21) * (oldvyouter - q.v1[1]) * m1 if x21 > 0 else -abs(x21) * (oldvyouter - q.v2[1]) * m2
        q.logged = 1
    return resolve


@pytest.mark.parametrize("law", [1, 2])
def test_exact_resolve_with_the_callers_arithmetic_is_bitwise(eng, law):
    """rebcu_collision_resolve_pairs: the device keeps the shuffled order, the sequential semantics and the early exits;
    the caller supplies the arithmetic.  With a resolver that restates the reference's expressions on the C library's
    trigonometry the result must be the oracle's BITS (velocities, collisions_plog, collisions_log_n, rand_seed) -- the
    drop-in uses the reference's own function in that place (tests/test_gpu_dropin.py: sheet scenarios)."""
    import math
    p = ics.shearing_sheet(root_size=40.0, seed=5)
    cfg = ics.shearing_sheet_config(root_size=40.0)
    min_v = 1.0 * ics.SHEET_OMEGA * 0.001
    steps = 4
    want, cw, aux = checkers.oracle().steps(cfg, p, steps, resolve=law, minimum_collision_velocity=min_v)
    if law == 1:
        eps_fn = lambda v: 1.0
    else:
        eps_fn = lambda v: min(1.0, max(0.0, 0.32 * math.pow(abs(v) * 100., -0.234)))
    resolver = _hardsphere_python(eps_fn, min_v)
    eng.upload(np.ascontiguousarray(p))
    c = cfg.copy()
    seed, plog, logn = 42, 0.0, 0
    nocol = c.copy()
    rounds_seen = 0
    for _ in range(steps):
        # one step = integrator step, boundary check, search, resolve (src/simulation.c:527-584)
        eng.integrator_step(c)
        eng.boundary_check(c)
        eng.collision_search(c, cap=1)
        seed, plog, logn, rounds = eng.collision_resolve_pairs(resolver, seed, plog, logn)
        rounds_seen = max(rounds_seen, rounds)
    got = eng.download()
    assert aux["collisions_log_n"] > 0 and logn == aux["collisions_log_n"]
    assert checkers.bits_equal(got, want)
    assert plog == aux["collisions_plog"]
    assert rounds_seen >= 2
