"""Randomised (hypothesis, derandomised) GPU parity through the C ABI against the oracle: the same strategy as
tests/test_oracle_random.py (random N, boxes, ghost rings, root-box grids, modes, test-particle settings)."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import checkers
from checkers import bits_equal, collisions_equal
from rebound_b200 import abi
from rebound_b200.simulation import Engine
from test_oracle_random import cfg_st, make, make_subset, subset_st

pytestmark = pytest.mark.gpu
# REBOUND_B200_FUZZ=<n>: a longer, non-derandomised campaign (default: 50 fixed examples per test)
_N = int(os.environ.get("REBOUND_B200_FUZZ", "50"))
SET = settings(max_examples=_N, deadline=None, derandomize=(_N == 50), suppress_health_check=list(HealthCheck))

_eng = None


def eng():
    global _eng
    if _eng is None:
        _eng = Engine(0)
    return _eng


@SET
@given(cfg_st, st.integers(0, 1))
def test_random_tree_cells_and_gravity(d, quad):
    c, p = make(d, gravity=abi.GRAVITY_TREE, quadrupole=quad)
    pb, cb = checkers.oracle().boundary_check(c, p)
    e = eng()
    e.upload(np.ascontiguousarray(pb))
    assert e.tree(cb.copy()).tobytes() == checkers.oracle().tree_dump(cb, pb).tobytes()
    want, cw = checkers.oracle().gravity(c, p)
    q, cc = p.copy(), c.copy()
    n = e.gravity_host(cc, q)
    assert n == len(want) and bits_equal(q[:n], want) and cc.N_active == cw.N_active


@SET
@given(cfg_st, st.sampled_from([abi.COLLISION_DIRECT, abi.COLLISION_TREE, abi.COLLISION_LINE, abi.COLLISION_LINETREE]),
       st.sampled_from([0.0, 0.05, -0.2]))
def test_random_collision_lists(d, col, dtl):
    c, p = make(d, collision=col, dt_last_done=dtl)
    got = eng().collision_search_host(c.copy(), np.ascontiguousarray(p))
    want = checkers.oracle().collision_search(c, p)
    assert collisions_equal(got, want, with_ri=(col in (abi.COLLISION_TREE, abi.COLLISION_LINETREE)))


@SET
@given(cfg_st, st.sampled_from([abi.GRAVITY_BASIC, abi.GRAVITY_COMPENSATED]), st.integers(0, 1), st.integers(0, 2),
       st.sampled_from([None, 0, 1, 2, 5]), st.sampled_from([abi.MODE_STRICT, abi.MODE_FAST]))
def test_random_direct_gravity(d, grav, tptype, terms, nactive, mode):
    c, p = make(d, gravity=grav, testparticle_type=tptype, gravity_ignore_terms=terms, mode=mode)
    if nactive is not None:
        c.N_active = min(nactive, len(p))
    want, _ = checkers.oracle().gravity(c, p)
    q = p.copy()
    eng().gravity_host(c.copy(), q)
    if mode == abi.MODE_STRICT:
        assert bits_equal(q, want)
    else:
        # 1e-12 relative of the largest acceleration in the system (single particles can cancel to ~0)
        scale = max(np.abs(want["ax"]).max(), np.abs(want["ay"]).max(), np.abs(want["az"]).max(), 1e-300)
        for f in ("ax", "ay", "az"):
            assert np.abs(q[f] - want[f]).max() <= 1e-12 * scale


@SET
@given(cfg_st, st.sampled_from([(abi.INTEGRATOR_LEAPFROG, 2), (abi.INTEGRATOR_LEAPFROG, 4), (abi.INTEGRATOR_SEI, 0)]),
       st.sampled_from([abi.GRAVITY_TREE, abi.GRAVITY_BASIC, abi.GRAVITY_NONE]))
def test_random_full_steps_without_resolve(d, integ, grav):
    c, p = make(d, gravity=grav, collision=abi.COLLISION_TREE, integrator=integ[0], leapfrog_order=integ[1])
    want, cw, _ = checkers.oracle().steps(c, p, 3, resolve=0)
    q, cc = p.copy(), c.copy()
    n = eng().steps_host(cc, q, 3)
    assert n == len(want) and bits_equal(q[:n], want) and cc.t == cw.t
    assert collisions_equal(eng().collisions_fetch(), checkers.oracle().collision_search(cw, want))


@SET
@given(cfg_st, st.sampled_from([abi.COLLISION_DIRECT, abi.COLLISION_TREE, abi.COLLISION_LINE, abi.COLLISION_LINETREE]),
       st.sampled_from([0.05, -0.2]), subset_st)
def test_random_collision_subsets(d, col, dtl, s):
    c, p = make(d, collision=col, dt_last_done=dtl)
    sub, nt = make_subset(s, len(p))
    e = eng()
    try:
        e.set_collision_subset(sub, nt)
        got = e.collision_search_host(c.copy(), np.ascontiguousarray(p))
    finally:
        e.set_collision_subset()
    want = checkers.oracle().collision_search_subset(c, p, sub, nt)
    assert collisions_equal(got, want, with_ri=(col in (abi.COLLISION_TREE, abi.COLLISION_LINETREE)))


@SET
@given(cfg_st, st.integers(0, 1), st.integers(0, 2), st.sampled_from([None, 0, 1, 2, 5]), st.sampled_from([0.37, -0.2, 1e-3]))
def test_random_jerk(d, tptype, terms, nactive, v):
    c, p = make(d, gravity=abi.GRAVITY_BASIC, testparticle_type=tptype, gravity_ignore_terms=terms)
    if nactive is not None:
        c.N_active = min(nactive, len(p))
    rng = np.random.default_rng(d["seed"] ^ 0x5a5a)
    for f in ("ax", "ay", "az"):
        p[f] = rng.normal(0, 1, len(p))
    want = checkers.oracle().apply_jerk(c, p, v)
    q = np.ascontiguousarray(p.copy())
    eng().jerk_host(c.copy(), q, v)
    assert bits_equal(q, want)


@SET
@given(cfg_st, st.sampled_from([0.0, 0.3, 0.45, 0.8]), st.sampled_from([0.0, 1e-3, 0.02, 0.2]))
def test_random_exit_checks(d, fmax, fmin):
    c, p = make(d)
    scale = d["root_size"] * max(d["nroot"])
    want = checkers.oracle().exit_check(c, p, fmax * scale, fmin * scale)
    e = eng()
    e.upload(np.ascontiguousarray(p))
    escape, encounter = e.exit_check(fmax * scale, fmin * scale)
    assert (3 if encounter else (4 if escape else 0)) == want
