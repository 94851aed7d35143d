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

    # 6. the searches / kicks / checks around the five configurations: LINE and LINETREE lists, r->map / N_targets
    #    subsets in all four search modes, the jerk kick of the modified-kick schemes, run_heartbeat's exit checks
    extras(ref)
    print("golden fixtures written to", HERE)


def extras_inputs():
    """Inputs of fixture 6 (shared with tests/test_oracle_golden.py and the GPU tests)."""
    rng = np.random.default_rng(29)
    n = 300
    q = abi.particles(n)
    for f in ("x", "y", "z"):
        q[f] = rng.uniform(-4.9, 4.9, n)
    for f in ("vx", "vy", "vz"):
        q[f] = rng.normal(0, 3, n)
    q["r"] = rng.uniform(0.05, 0.35, n)
    q["m"] = rng.uniform(0.5, 1.5, n) / n
    sub = rng.permutation(n)[:200].astype(np.uint64)
    base = dict(root_size=10.0, boundary=abi.BOUNDARY_PERIODIC, N_ghost_x=1, N_ghost_y=1, N_ghost_z=0, dt_last_done=0.04)
    return q, sub, 60, base


def extras(ref):
    q, sub, nt, base = extras_inputs()
    d = {"particles_in": raw(q), "map": sub, "n_targets": nt}
    for mode in (abi.COLLISION_DIRECT, abi.COLLISION_TREE, abi.COLLISION_LINE, abi.COLLISION_LINETREE):
        c = abi.default_config(collision=mode, **base)
        d[f"col_m{mode}"] = raw(ref.collision_search(c, q))
        d[f"col_m{mode}_map"] = raw(ref.collision_search_subset(c, q, sub, None))
        d[f"col_m{mode}_map_targets"] = raw(ref.collision_search_subset(c, q, sub, nt))
        d[f"col_m{mode}_targets"] = raw(ref.collision_search_subset(c, q, None, nt))
    cj = abi.default_config(softening=0.05, N_active=100, testparticle_type=1)
    qa, _ = ref.gravity(cj, q)
    d["jerk_in"] = raw(qa)
    d["jerk_out"] = raw(ref.apply_jerk(cj, qa, 0.37))
    cj2 = abi.default_config(softening=0.05, gravity_ignore_terms=abi.IGNORE_TERMS_INVOLVING_0)
    d["jerk_out_ignore0"] = raw(ref.apply_jerk(cj2, qa, -0.11))
    r = np.sqrt(q["x"] ** 2 + q["y"] ** 2 + q["z"] ** 2)
    d["exit_max"] = np.array([float(r.max()), float(np.nextafter(r.max(), 0.0)), 2.0, 0.0])
    d["exit_min"] = np.array([0.0, 0.05, 0.5])
    d["exit_status"] = np.array([[ref.exit_check(cj, q, mx, mn) for mn in d["exit_min"]] for mx in d["exit_max"]])
    np.savez_compressed(os.path.join(HERE, "extras300.npz"), **d)


if __name__ == "__main__":
    main()
