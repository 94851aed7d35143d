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
