"""The oracle restatement against the committed golden vectors (produced from the unmodified
reference by tests/golden/make_golden.py).  Runs without /root/reference and without oracle/_ref."""
import os

import numpy as np
import pytest

import checkers
from rebound_b200 import abi, ics

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def as_particles(a):
    return np.frombuffer(a.tobytes(), dtype=abi.PARTICLE_DTYPE).copy()


def acc(p):
    return np.stack([p["ax"], p["ay"], p["az"]], axis=1)


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def test_direct_golden():
    g = load("direct_plummer512.npz")
    p = as_particles(g["particles_in"])
    for key, grav in (("basic", abi.GRAVITY_BASIC), ("compensated", abi.GRAVITY_COMPENSATED)):
        cfg = ics.plummer_config(512, gravity=grav)
        assert cfg.softening == float(g["softening"])
        out, _ = checkers.oracle().gravity(cfg, p)
        assert same_bits(acc(out), g["acc_" + key])


@pytest.mark.parametrize("order", [2, 4, 6, 8])
def test_leapfrog_golden(order):
    g = load("leapfrog_plummer128.npz")
    p = as_particles(g["particles_in"])
    cfg = ics.plummer_config(128, leapfrog_order=order, dt=1e-3)
    out, c, _ = checkers.oracle().steps(cfg, p, 5)
    assert out.tobytes() == as_particles(g[f"out_order{order}"]).tobytes()
    assert c.t == float(g[f"t_order{order}"])


def test_tree_disc_golden():
    g = load("tree_disc400.npz")
    p = as_particles(g["particles_in"])
    cfg = ics.selfgravity_disc_config()
    pb, cb = checkers.oracle().boundary_check(cfg, p)
    assert len(pb) == int(g["n_after_boundary"])
    cells = checkers.oracle().tree_dump(cb, pb)
    assert cells.tobytes() == g["cells"].tobytes()
    out, _ = checkers.oracle().gravity(cfg, p)
    assert len(out) == int(g["n_out"])
    assert same_bits(acc(out), g["acc"])


def test_sheet_golden():
    g = load("sheet_root30.npz")
    p = as_particles(g["particles_in"])
    cfg = ics.shearing_sheet_config(root_size=30.0, t=55.5)
    assert checkers.oracle().tree_dump(cfg, p).tobytes() == g["cells"].tobytes()
    out, _ = checkers.oracle().gravity(cfg, p)
    assert same_bits(acc(out), g["acc"])
    col_t = checkers.oracle().collision_search(cfg, p)
    want_t = np.frombuffer(g["col_tree"].tobytes(), dtype=abi.COLLISION_DTYPE)
    assert len(want_t) > 0 and checkers.collisions_equal(col_t, want_t)
    cfg_d = ics.shearing_sheet_config(root_size=30.0, t=55.5, collision=abi.COLLISION_DIRECT)
    col_d = checkers.oracle().collision_search(cfg_d, p)
    want_d = np.frombuffer(g["col_direct"].tobytes(), dtype=abi.COLLISION_DTYPE)
    assert checkers.collisions_equal(col_d, want_d, with_ri=False)
    cfg0 = ics.shearing_sheet_config(root_size=30.0)
    fin, c, aux = checkers.oracle().steps(cfg0, p, 10, resolve=2, minimum_collision_velocity=float(g["mcv"]))
    assert fin.tobytes() == as_particles(g["steps_out"]).tobytes()
    assert aux["collisions_log_n"] == int(g["steps_log_n"]) > 0
    assert aux["collisions_plog"] == float(g["steps_plog"])
    assert c.t == float(g["steps_t"])


@pytest.mark.parametrize("b", [abi.BOUNDARY_OPEN, abi.BOUNDARY_PERIODIC, abi.BOUNDARY_SHEAR])
def test_boundary_golden(b):
    g = load("boundary300.npz")
    p = as_particles(g["particles_in"])
    c = abi.default_config(boundary=b, root_size=10.0, N_root_x=2, N_root_y=1, N_root_z=1, OMEGA=0.7, t=3.3, N_active=40)
    out, cc = checkers.oracle().boundary_check(c, p)
    assert out.tobytes() == as_particles(g[f"out_b{b}"]).tobytes()
    assert cc.N_active == int(g[f"n_active_b{b}"])


def test_extras_golden():
    """LINE / LINETREE lists, r->map / N_targets subsets in all four search modes, the jerk kick, the exit checks."""
    import sys
    sys.path.insert(0, GOLD)
    from make_golden import extras_inputs
    g = load("extras300.npz")
    q, sub, nt, base = extras_inputs()
    assert q.tobytes() == as_particles(g["particles_in"]).tobytes() and np.array_equal(sub, g["map"]) and nt == int(g["n_targets"])
    orc = checkers.oracle()
    n_hits = 0
    for mode in (abi.COLLISION_DIRECT, abi.COLLISION_TREE, abi.COLLISION_LINE, abi.COLLISION_LINETREE):
        c = abi.default_config(collision=mode, **base)
        tree = mode in (abi.COLLISION_TREE, abi.COLLISION_LINETREE)
        for key, got in ((f"col_m{mode}", orc.collision_search(c, q)),
                         (f"col_m{mode}_map", orc.collision_search_subset(c, q, sub, None)),
                         (f"col_m{mode}_map_targets", orc.collision_search_subset(c, q, sub, nt)),
                         (f"col_m{mode}_targets", orc.collision_search_subset(c, q, None, nt))):
            want = np.frombuffer(g[key].tobytes(), dtype=abi.COLLISION_DTYPE)
            assert checkers.collisions_equal(got, want, with_ri=tree), key
            n_hits += len(want)
    assert n_hits > 100
    cj = abi.default_config(softening=0.05, N_active=100, testparticle_type=1)
    qa = as_particles(g["jerk_in"])
    assert orc.apply_jerk(cj, qa, 0.37).tobytes() == as_particles(g["jerk_out"]).tobytes()
    cj2 = abi.default_config(softening=0.05, gravity_ignore_terms=abi.IGNORE_TERMS_INVOLVING_0)
    assert orc.apply_jerk(cj2, qa, -0.11).tobytes() == as_particles(g["jerk_out_ignore0"]).tobytes()
    got = np.array([[orc.exit_check(cj, q, mx, mn) for mn in g["exit_min"]] for mx in g["exit_max"]])
    assert np.array_equal(got, g["exit_status"]) and set(got.ravel()) == {0, 3, 4}
