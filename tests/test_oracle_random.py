"""Randomised (hypothesis, derandomised) comparison of the oracle restatement with the unmodified reference:
random particle counts, boxes, ghost rings, root-box grids, gravity / collision modes and test-particle
settings.  Everything must agree bit for bit."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import checkers
from checkers import bits_equal, collisions_equal
from rebound_b200 import abi

pytestmark = pytest.mark.needs_ref

SET = settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))


def random_state(seed, n, box, spread=0.49):
    rng = np.random.default_rng(seed)
    p = abi.particles(n)
    p["x"] = rng.uniform(-spread, spread, n) * box[0]
    p["y"] = rng.uniform(-spread, spread, n) * box[1]
    p["z"] = rng.uniform(-spread, spread, n) * box[2]
    for f in ("vx", "vy", "vz"):
        p[f] = rng.normal(0, 0.3, n)
    p["m"] = rng.uniform(0.1, 2.0, n) / n
    p["r"] = rng.uniform(0.01, 0.08, n) * min(box)
    return p


cfg_st = st.fixed_dictionaries({
    "seed": st.integers(0, 2**31 - 1),
    "n": st.integers(1, 160),
    "root_size": st.sampled_from([1.0, 10.2, 3.7, 100.0]),
    "nroot": st.tuples(st.integers(1, 3), st.integers(1, 3), st.integers(1, 2)),
    "nghost": st.tuples(st.integers(0, 2), st.integers(0, 2), st.integers(0, 1)),
    "boundary": st.sampled_from([abi.BOUNDARY_OPEN, abi.BOUNDARY_PERIODIC, abi.BOUNDARY_SHEAR]),
    "theta2": st.sampled_from([0.0, 0.25, 0.5, 1.5]),
    "softening": st.sampled_from([0.0, 0.01, 0.3]),
    "t": st.sampled_from([0.0, 1.7, 123.456]),
})


def make(d, **kw):
    box = [d["root_size"] * k for k in d["nroot"]]
    p = random_state(d["seed"], d["n"], box)
    c = abi.default_config(root_size=d["root_size"], N_root_x=d["nroot"][0], N_root_y=d["nroot"][1], N_root_z=d["nroot"][2],
                           N_ghost_x=d["nghost"][0], N_ghost_y=d["nghost"][1], N_ghost_z=d["nghost"][2],
                           boundary=d["boundary"], opening_angle2=d["theta2"], softening=d["softening"], t=d["t"],
                           OMEGA=0.31, G=1.0, dt=0.01, **kw)
    return c, p


@SET
@given(cfg_st)
def test_random_tree_cells_and_gravity(d):
    c, p = make(d, gravity=abi.GRAVITY_TREE)
    ref = checkers.reference()
    pb, cb = ref.boundary_check(c, p)
    assert ref.tree_dump(cb, pb).tobytes() == checkers.oracle().tree_dump(cb, pb).tobytes()
    a, ca = ref.gravity(c, p)
    b, cb2 = checkers.oracle().gravity(c, p)
    assert len(a) == len(b) and bits_equal(a, b) and ca.N_active == cb2.N_active


@SET
@given(cfg_st, st.sampled_from([abi.COLLISION_DIRECT, abi.COLLISION_TREE, abi.COLLISION_LINE, abi.COLLISION_LINETREE]),
       st.sampled_from([0.0, 0.05, -0.2]))
def test_random_collision_lists(d, col, dtl):
    c, p = make(d, collision=col, dt_last_done=dtl)
    a = checkers.reference().collision_search(c, p)
    b = checkers.oracle().collision_search(c, p)
    assert collisions_equal(a, b, with_ri=(col in (abi.COLLISION_TREE, abi.COLLISION_LINETREE)))


@SET
@given(cfg_st, st.sampled_from([abi.GRAVITY_BASIC, abi.GRAVITY_COMPENSATED]), st.integers(0, 1), st.integers(0, 2),
       st.sampled_from([None, 0, 1, 2, 5]))
def test_random_direct_gravity(d, grav, tptype, terms, nactive):
    c, p = make(d, gravity=grav, testparticle_type=tptype, gravity_ignore_terms=terms)
    if nactive is not None:
        c.N_active = min(nactive, len(p))
    # the gather form is the reference's OpenMP build; with ghost boxes its serial build differs in the last bits
    a, _ = checkers.reference(openmp=True).gravity(c, p)
    b, _ = checkers.oracle().gravity(c, p)
    assert bits_equal(a, b)


@SET
@given(cfg_st, st.sampled_from([(abi.INTEGRATOR_LEAPFROG, 2), (abi.INTEGRATOR_LEAPFROG, 4), (abi.INTEGRATOR_SEI, 0)]))
def test_random_full_steps(d, integ):
    c, p = make(d, gravity=abi.GRAVITY_TREE, collision=abi.COLLISION_TREE, integrator=integ[0], leapfrog_order=integ[1])
    a, ca, xa = checkers.reference().steps(c, p, 3, resolve=1)
    b, cb, xb = checkers.oracle().steps(c, p, 3, resolve=1)
    assert len(a) == len(b) and bits_equal(a, b)
    assert ca.t == cb.t and xa["collisions_log_n"] == xb["collisions_log_n"] and xa["collisions_plog"] == xb["collisions_plog"]


@SET
@given(cfg_st)
def test_random_quadrupole_gravity(d):
    """cfg.quadrupole = 1 against the reference compiled with -DQUADRUPOLE (src/tree.c:148-198, 293-303)."""
    ref = checkers.reference(quadrupole=True)
    if ref is None:
        pytest.skip("oracle/_ref/libref_harness_quad.so not built")
    c, p = make(d, gravity=abi.GRAVITY_TREE, quadrupole=1)
    a, _ = ref.gravity(c, p)
    b, _ = checkers.oracle().gravity(c, p)
    assert len(a) == len(b) and bits_equal(a, b)


subset_st = st.fixed_dictionaries({
    "map_frac": st.sampled_from([None, 0.3, 0.7, 1.0]),       # None: no r->map
    "targets_frac": st.sampled_from([None, 0.0, 0.4, 1.0]),   # None: N_targets = SIZE_MAX
    "mseed": st.integers(0, 2**31 - 1),
})


def make_subset(s, n):
    """(map or None, N_targets or None) for n particles: a random selection in random order, targets <= projectiles."""
    rng = np.random.default_rng(s["mseed"])
    sub = None
    n_proj = n
    if s["map_frac"] is not None:
        n_proj = int(round(s["map_frac"] * n))
        sub = rng.permutation(n)[:n_proj].astype(np.uint64)
    nt = None if s["targets_frac"] is None else int(s["targets_frac"] * n_proj)
    return sub, nt


@SET
@given(cfg_st, st.sampled_from([abi.COLLISION_DIRECT, abi.COLLISION_TREE, abi.COLLISION_LINE, abi.COLLISION_LINETREE]),
       st.sampled_from([0.05, -0.2]), subset_st)
def test_random_collision_subsets(d, col, dtl, s):
    c, p = make(d, collision=col, dt_last_done=dtl)
    sub, nt = make_subset(s, len(p))
    a = checkers.reference().collision_search_subset(c, p, sub, nt)
    b = checkers.oracle().collision_search_subset(c, p, sub, nt)
    assert collisions_equal(a, b, with_ri=(col in (abi.COLLISION_TREE, abi.COLLISION_LINETREE)))


@SET
@given(cfg_st, st.integers(0, 1), st.integers(0, 2), st.sampled_from([None, 0, 1, 2, 5]), st.sampled_from([0.37, -0.2, 1e-3]))
def test_random_jerk(d, tptype, terms, nactive, v):
    """reb_gravity_basic_calculate_and_apply_jerk on random states (the accelerations are whatever the state holds)."""
    c, p = make(d, gravity=abi.GRAVITY_BASIC, testparticle_type=tptype, gravity_ignore_terms=terms)
    if nactive is not None:
        c.N_active = min(nactive, len(p))
    rng = np.random.default_rng(d["seed"] ^ 0x5a5a)
    for f in ("ax", "ay", "az"):
        p[f] = rng.normal(0, 1, len(p))
    a = checkers.reference().apply_jerk(c, p, v)
    b = checkers.oracle().apply_jerk(c, p, v)
    assert bits_equal(a, b)


@SET
@given(cfg_st, st.sampled_from([0.0, 0.3, 0.45, 0.8]), st.sampled_from([0.0, 1e-3, 0.02, 0.2]))
def test_random_exit_checks(d, fmax, fmin):
    c, p = make(d)
    scale = d["root_size"] * max(d["nroot"])
    assert checkers.reference().exit_check(c, p, fmax * scale, fmin * scale) == checkers.oracle().exit_check(c, p, fmax * scale, fmin * scale)
