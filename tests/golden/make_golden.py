"""Generates the golden fixtures in this directory from the UNMODIFIED reference
(oracle/_ref/libref_harness.so, built from /root/reference/src by oracle/Makefile).

Run in the authoring container:  python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_*.py)
on machines where neither /root/reference nor oracle/_ref exists.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import checkers  # noqa: E402
from rebound_b200 import abi, ics  # noqa: E402


def raw(p):
    return np.frombuffer(p.tobytes(), dtype=np.uint8)


def acc(p):
    return np.stack([p["ax"], p["ay"], p["az"]], axis=1)


def cfg_bytes(c):
    return np.frombuffer(bytes(c), dtype=np.uint8)


def main():
    ref = checkers.reference()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"

    # 1. direct summation on the reference's own Plummer generator (tools.c:463-502), seed 42
    n = 512
    p = abi.particles(n)
    ref.lib.refh_make_plummer.argtypes = [C.c_uint64, C.c_double, C.c_double, C.c_uint, C.c_void_p]
    ref.lib.refh_make_plummer(n, 1.0, 1.0, 42, abi.as_ptr(p))
    cfg = ics.plummer_config(n)
    out_b, _ = ref.gravity(cfg, p)
    cfg_c = ics.plummer_config(n, gravity=abi.GRAVITY_COMPENSATED)
    out_c, _ = ref.gravity(cfg_c, p)
    np.savez_compressed(os.path.join(HERE, "direct_plummer512.npz"), particles_in=raw(p), softening=cfg.softening,
                        acc_basic=acc(out_b), acc_compensated=acc(out_c))

    # 2. leapfrog trajectories, orders 2..8, 5 steps
    d = {"particles_in": raw(p[:128].copy())}
    for order in (2, 4, 6, 8):
        c = ics.plummer_config(128, leapfrog_order=order, dt=1e-3)
        q, cc, _ = ref.steps(c, p[:128].copy(), 5)
        d[f"out_order{order}"] = raw(q)
        d[f"t_order{order}"] = cc.t
    np.savez_compressed(os.path.join(HERE, "leapfrog_plummer128.npz"), **d)

    # 3. octree + tree gravity on a disc (examples/selfgravity_disc recipe), theta^2 = 0.25
    pd = ics.selfgravity_disc(400, seed=2)
    cd = ics.selfgravity_disc_config()
    pdb, cdb = ref.boundary_check(cd, pd)
    cells = ref.tree_dump(cdb, pdb)
    gd, _ = ref.gravity(cd, pd)
    np.savez_compressed(os.path.join(HERE, "tree_disc400.npz"), particles_in=raw(pd), cells=raw(cells),
                        n_after_boundary=len(pdb), acc=acc(gd), n_out=len(gd))

    # 4. shearing sheet (examples/shearing_sheet recipe): tree cells with 2x2 root boxes, tree gravity with
    #    25 ghost boxes, collision lists (tree and direct), 10 full steps with hard-sphere resolve
    ps = ics.shearing_sheet(root_size=30.0, seed=9)
    cs = ics.shearing_sheet_config(root_size=30.0, t=55.5)
    cells_s = ref.tree_dump(cs, ps)
    gs, _ = ref.gravity(cs, ps)
    col_t = ref.collision_search(cs, ps)
    cs_d = ics.shearing_sheet_config(root_size=30.0, t=55.5, collision=abi.COLLISION_DIRECT)
    col_d = ref.collision_search(cs_d, ps)
    cs0 = ics.shearing_sheet_config(root_size=30.0)
    mcv = 1.0 * ics.SHEET_OMEGA * 0.001
    fin, cf, aux = ref.steps(cs0, ps, 10, resolve=2, minimum_collision_velocity=mcv)
    np.savez_compressed(os.path.join(HERE, "sheet_root30.npz"), particles_in=raw(ps), cells=raw(cells_s), acc=acc(gs),
                        col_tree=raw(col_t), col_direct=raw(col_d), steps_out=raw(fin), steps_t=cf.t,
                        steps_log_n=aux["collisions_log_n"], steps_plog=aux["collisions_plog"], mcv=mcv)

    # 5. boundary checks
    rng = np.random.default_rng(11)
    nb = 300
    pb = abi.particles(nb)
    for f in ("x", "y", "z"):
        pb[f] = rng.uniform(-14, 14, nb)
    for f in ("vx", "vy", "vz"):
        pb[f] = rng.normal(0, 1, nb)
    pb["m"] = 1.0
    d = {"particles_in": raw(pb)}
    for b in (abi.BOUNDARY_OPEN, abi.BOUNDARY_PERIODIC, abi.BOUNDARY_SHEAR):
        c = abi.default_config(boundary=b, root_size=10.0, N_root_x=2, N_root_y=1, N_root_z=1, OMEGA=0.7, t=3.3, N_active=40)
        q, cc = ref.boundary_check(c, pb)
        d[f"out_b{b}"] = raw(q)
        d[f"n_active_b{b}"] = cc.N_active
    np.savez_compressed(os.path.join(HERE, "boundary300.npz"), **d)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
