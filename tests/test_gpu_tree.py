"""GPU parity for the tree path through the C ABI: octree cells (bit-exact leaf / cell assignment,
geometry and moments), Barnes-Hut accelerations, boundary check, SEI, collision lists, whole steps.
Checked against the CPU oracle (pinned to the reference) and the committed golden vectors."""
import os

import numpy as np
import pytest

import checkers
from checkers import bits_equal, collisions_equal, max_rel_acc_error
from rebound_b200 import abi, ics
from rebound_b200.simulation import Engine, ReboundCudaError

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def eng():
    e = Engine(0)
    yield e
    e.close()


def tree_cases():
    yield "disc", ics.selfgravity_disc_config(), ics.selfgravity_disc(3000, seed=2)
    yield "disc_theta1.5", ics.selfgravity_disc_config(opening_angle2=1.5), ics.selfgravity_disc(700, seed=3)
    p = ics.plummer(2000, seed=4)
    yield "plummer_box", ics.plummer_config(2000, gravity=abi.GRAVITY_TREE, root_size=200.0, opening_angle2=0.25), p
    yield "sheet", ics.shearing_sheet_config(root_size=40.0, t=123.4), ics.shearing_sheet(root_size=40.0, seed=5)
    yield "rootboxes_3d", ics.plummer_config(2000, gravity=abi.GRAVITY_TREE, root_size=50.0, N_root_x=2, N_root_y=3,
                                            N_root_z=2, boundary=abi.BOUNDARY_PERIODIC, N_ghost_x=1, N_ghost_y=1, N_ghost_z=1), p
    yield "single", ics.selfgravity_disc_config(), ics.selfgravity_disc(0, seed=1)
    yield "pair", ics.selfgravity_disc_config(), ics.selfgravity_disc(1, seed=1)
    # deep tree: pairs closer than 2^-21 of the box force the tie-break path beyond the 63-bit key
    q = ics.selfgravity_disc(400, seed=6)
    q["x"][201:] = q["x"][1:201] + 1e-9 * np.arange(1, 201)
    q["y"][201:] = q["y"][1:201] - 3e-10
    q["z"][201:] = q["z"][1:201]
    yield "deep_pairs", ics.selfgravity_disc_config(), q
    t = ics.selfgravity_disc(300, seed=7)
    t["x"][101:201] = t["x"][1:101] * (1 + 2.0**-50)
    t["y"][101:201] = t["y"][1:101]
    t["z"][101:201] = t["z"][1:101]
    t["x"][201:] = t["x"][1:101] * (1 + 2.0**-49)
    t["y"][201:] = t["y"][1:101]
    t["z"][201:] = t["z"][1:101]
    yield "ulp_triples", ics.selfgravity_disc_config(), t


TREE_CASES = list(tree_cases())
IDS = [c[0] for c in TREE_CASES]


@pytest.mark.parametrize("name,cfg,p", TREE_CASES, ids=IDS)
def test_tree_cells_bitwise(eng, name, cfg, p):
    pb, cb = checkers.oracle().boundary_check(cfg, p)
    want = checkers.oracle().tree_dump(cb, pb)
    eng.upload(np.ascontiguousarray(pb))
    got = eng.tree(cb.copy())
    assert len(got) == len(want)
    for f in abi.TREECELL_DTYPE.names:
        assert np.array_equal(got[f].view(np.uint64 if got[f].dtype == np.float64 else np.int32),
                              want[f].view(np.uint64 if want[f].dtype == np.float64 else np.int32)), f
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("name,cfg,p", TREE_CASES, ids=IDS)
def test_tree_gravity_bitwise(eng, name, cfg, p):
    want, cw = checkers.oracle().gravity(cfg, p)
    q, c = p.copy(), cfg.copy()
    n = eng.gravity_host(c, q)
    assert n == len(want)
    assert bits_equal(q[:n], want)
    assert c.N_active == cw.N_active


def _acc(q):
    return np.stack([q["ax"], q["ay"], q["az"]], 1)


def _rel_err(a, ref):
    """|a - ref| / |ref| per particle; a particle whose net acceleration nearly cancels (the star in the middle of a
    disc) is measured against a tenth of the median acceleration instead of its own tiny one."""
    nrm = np.linalg.norm(ref, axis=1)
    floor = 0.1 * np.median(nrm) if len(nrm) else 0.0
    den = np.maximum(nrm, floor)
    ok = den > 0
    return np.linalg.norm(a - ref, axis=1)[ok] / den[ok]


def _three_way(eng, cfg, p):
    """(direct, strict tree, FAST tree) accelerations of the particles the boundary check leaves in place."""
    pb, cb = checkers.oracle().boundary_check(cfg, p)
    pb = np.ascontiguousarray(pb)
    cd = cb.copy(); cd.gravity = abi.GRAVITY_BASIC
    qd, qs, qf = pb.copy(), pb.copy(), pb.copy()
    eng.gravity_host(cd, qd)
    eng.gravity_host(cb.copy(), qs)
    cf = cb.copy(); cf.mode = abi.MODE_FAST
    n = eng.gravity_host(cf, qf)
    assert n == len(pb)
    return _acc(qd), _acc(qs), _acc(qf)


@pytest.mark.parametrize("name,cfg,p", TREE_CASES[:5], ids=IDS[:5])
def test_tree_gravity_fast_mode(eng, name, cfg, p):
    """REBCU_MODE_FAST = the group walk (walk_group_kernel): its opening criterion is at least as strict as the
    reference's per-particle one, so its error against direct summation must not exceed the reference's own at the
    same opening angle -- BASELINE.json's tolerance for tree accelerations.  (The strict tree equals the reference bit
    for bit, test_tree_gravity_bitwise.)  rms error: no worse than the reference's; worst particle: within 2x."""
    a_d, a_s, a_f = _three_way(eng, cfg, p)
    e_s, e_f = _rel_err(a_s, a_d), _rel_err(a_f, a_d)
    assert np.all(np.isfinite(a_f))
    assert np.sqrt(np.mean(e_f**2)) <= 1.02 * np.sqrt(np.mean(e_s**2)) + 1e-13, (np.sqrt(np.mean(e_f**2)), np.sqrt(np.mean(e_s**2)))
    assert e_f.max() <= 2.0 * e_s.max() + 1e-12


@pytest.mark.parametrize("n", [0, 1, 2, 30, 31, 32, 33, 127, 129, 1000])
def test_tree_gravity_fast_mode_group_edges(eng, n):
    """Group sizes around the warp width (idle lanes shadow the last particle), a lone particle, a pair; theta = 0
    opens every cell, which turns the tree walk into a direct sum: the FAST walk must then agree with direct summation
    to rounding (1e-12 relative)."""
    p = ics.selfgravity_disc(n, seed=11)
    cfg = ics.selfgravity_disc_config(opening_angle2=0.0)
    a_d, a_s, a_f = _three_way(eng, cfg, p)
    if len(a_d) > 1:
        assert _rel_err(a_f, a_d).max() <= 1e-12
    else:
        assert np.all(a_f == 0.0)


def test_tree_gravity_fast_mode_without_softening_and_deep_pairs(eng):
    """softening = 0 (a particle's own leaf must be skipped by index, not by its zero distance) on a tree with
    pairs deeper than the 63-bit key."""
    name, cfg, p = TREE_CASES[IDS.index("deep_pairs")]
    cfg = cfg.copy(); cfg.softening = 0.0
    a_d, a_s, a_f = _three_way(eng, cfg, p)
    assert np.all(np.isfinite(a_f))
    e_s, e_f = _rel_err(a_s, a_d), _rel_err(a_f, a_d)
    assert np.sqrt(np.mean(e_f**2)) <= 1.02 * np.sqrt(np.mean(e_s**2)) + 1e-13


def test_tree_gravity_fast_mode_stack_overflow_falls_back():
    """A traversal stack too small for the tree (forced with REBOUND_B200_GW_STACK, read once per process) sends the
    warp to the per-particle FAST walk: same accuracy class."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import checkers
        from rebound_b200 import abi, ics
        from rebound_b200.simulation import Engine
        eng = Engine(0)
        cfg = ics.selfgravity_disc_config()
        p = ics.selfgravity_disc(5000, seed=3)
        pb, cb = checkers.oracle().boundary_check(cfg, p)
        pb = np.ascontiguousarray(pb)
        qd, qs, qf = pb.copy(), pb.copy(), pb.copy()
        cd = cb.copy(); cd.gravity = abi.GRAVITY_BASIC
        eng.gravity_host(cd, qd)
        eng.gravity_host(cb.copy(), qs)
        cf = cb.copy(); cf.mode = abi.MODE_FAST
        eng.gravity_host(cf, qf)
        st = eng.tree_walk_stats(cf)
        acc = lambda q: np.stack([q["ax"], q["ay"], q["az"]], 1)
        err = lambda a, b: np.sqrt(np.mean((np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1))**2))
        print("ERR_FAST", err(acc(qf), acc(qd)), "ERR_STRICT", err(acc(qs), acc(qd)), "GROUPS", st["groups"], "OF", (len(pb) + 31) // 32)
    """) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, REBOUND_B200_GW_STACK="12")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    tok = r.stdout.split()
    val = lambda k: float(tok[tok.index(k) + 1])
    assert val("GROUPS") < val("OF")                         # some warps did overflow and took the fallback
    assert val("ERR_FAST") <= 1.02 * val("ERR_STRICT") + 1e-13


def test_tree_walk_stats(eng):
    """rebcu_tree_walk_stats: the per-particle interaction count equals the oracle's tree-walk count, the group walk
    evaluates at least as many entries per particle (stricter criterion)."""
    cfg = ics.selfgravity_disc_config()
    p = ics.selfgravity_disc(4095, seed=2)
    pb, cb = checkers.oracle().boundary_check(cfg, p)
    pb = np.ascontiguousarray(pb)
    cf = cb.copy(); cf.mode = abi.MODE_FAST
    eng.upload(pb)
    eng.update_acceleration(cf)
    st = eng.tree_walk_stats(cf)
    assert 0.9 * ((len(pb) + 31) // 32) <= st["groups"] <= (len(pb) + 31) // 32
    assert st["cells"] == len(checkers.oracle().tree_dump(cb, pb))
    assert st["group_entries"] * 32 >= st["interactions"] > 100 * len(pb)
    assert st["visits"] > st["interactions"]


def test_tree_error_vs_direct_matches_reference_level(eng):
    """Tree accuracy against direct summation is the reference's own (identical interaction lists):
    rms relative error at theta^2=0.25 on a Plummer sphere stays at the few-1e-3 level."""
    n = 16384
    p = ics.plummer(n, seed=42)
    cd = ics.plummer_config(n)
    ct = ics.plummer_config(n, gravity=abi.GRAVITY_TREE, root_size=400.0, opening_angle2=0.25)
    inside = (np.abs(p["x"]) < 200) & (np.abs(p["y"]) < 200) & (np.abs(p["z"]) < 200)
    p = np.ascontiguousarray(p[inside])
    qd, qt = p.copy(), p.copy()
    eng.gravity_host(cd.copy(), qd)
    eng.gravity_host(ct.copy(), qt)
    a_d = np.stack([qd["ax"], qd["ay"], qd["az"]], 1)
    a_t = np.stack([qt["ax"], qt["ay"], qt["az"]], 1)
    rel = np.linalg.norm(a_t - a_d, axis=1) / np.linalg.norm(a_d, axis=1)
    assert 1e-4 < np.sqrt(np.mean(rel**2)) < 1e-2


def test_tree_errors(eng):
    cfg = ics.selfgravity_disc_config()
    p = ics.selfgravity_disc(50, seed=2)
    bad = p.copy(); bad["x"][7] = bad["x"][3]; bad["y"][7] = bad["y"][3]; bad["z"][7] = bad["z"][3]
    nan = p.copy(); nan["y"][5] = np.nan
    noroot = ics.selfgravity_disc_config(root_size=-1.0)
    outside = p.copy(); outside["x"][9] = 100.0
    cfg_nob = ics.selfgravity_disc_config(boundary=abi.BOUNDARY_NONE)
    both = outside.copy(); both["y"][5] = np.nan        # index 5 < 9: non-finite is reported first
    for c, q in ((cfg, bad), (cfg, nan), (noroot, p), (cfg_nob, outside), (cfg_nob, both)):
        with pytest.raises(checkers.CheckerError) as e1:
            checkers.oracle().tree_dump(c, q)
        eng.upload(np.ascontiguousarray(q))
        with pytest.raises(ReboundCudaError) as e2:
            eng.tree(c.copy())
        assert e1.value.msg == e2.value.msg
        assert e1.value.code == e2.value.code


def test_tree_golden(eng):
    g = np.load(os.path.join(GOLD, "tree_disc400.npz"))
    p = np.frombuffer(g["particles_in"].tobytes(), dtype=abi.PARTICLE_DTYPE).copy()
    cfg = ics.selfgravity_disc_config()
    q, c = p.copy(), cfg.copy()
    n = eng.gravity_host(c, q)
    assert n == int(g["n_out"])
    got = np.stack([q["ax"][:n], q["ay"][:n], q["az"][:n]], 1)
    assert np.array_equal(got.view(np.uint64), g["acc"].view(np.uint64))
    eng.boundary_check(c)
    cells = eng.tree(c)
    assert cells.tobytes() == g["cells"].tobytes()


@pytest.mark.parametrize("boundary", [abi.BOUNDARY_OPEN, abi.BOUNDARY_PERIODIC, abi.BOUNDARY_SHEAR])
def test_boundary_bitwise(eng, boundary):
    rng = np.random.default_rng(11)
    n = 5000
    p = abi.particles(n)
    for f in ("x", "y", "z"):
        p[f] = rng.uniform(-14, 14, n)
    for f in ("vx", "vy", "vz"):
        p[f] = rng.normal(0, 1, n)
    p["m"] = 1.0
    p["name"] = np.arange(n)          # pointer fields must travel with their particle
    cfg = abi.default_config(boundary=boundary, root_size=10.0, N_root_x=2, N_root_y=1, N_root_z=1, OMEGA=0.7, t=3.3, N_active=40)
    want, cw = checkers.oracle().boundary_check(cfg, p)
    eng.upload(p)
    c = cfg.copy()
    eng.boundary_check(c)
    got = eng.download()
    assert len(got) == len(want)
    assert bits_equal(got, want)
    assert np.array_equal(got["name"], want["name"])
    assert c.N_active == cw.N_active


def test_boundary_open_removes_everything(eng):
    p = abi.particles(5)
    p["x"] = 100.0
    cfg = abi.default_config(boundary=abi.BOUNDARY_OPEN, root_size=10.0, N_active=3)
    want, cw = checkers.oracle().boundary_check(cfg, p)
    eng.upload(p)
    c = cfg.copy()
    eng.boundary_check(c)
    assert eng.N == 0 == len(want)
    assert c.N_active == cw.N_active


def collision_cases():
    p = ics.shearing_sheet(root_size=40.0, seed=5)
    yield "sheet_tree", ics.shearing_sheet_config(root_size=40.0, t=55.5), p
    yield "sheet_direct", ics.shearing_sheet_config(root_size=40.0, t=55.5, collision=abi.COLLISION_DIRECT), p
    rng = np.random.default_rng(3)
    n = 1500
    q = abi.particles(n)
    for f in ("x", "y", "z"):
        q[f] = rng.uniform(-4.9, 4.9, n)
    for f in ("vx", "vy", "vz"):
        q[f] = rng.normal(0, 1, n)
    q["r"] = rng.uniform(0.02, 0.25, n)
    q["m"] = 1.0
    for col in (abi.COLLISION_DIRECT, abi.COLLISION_TREE):
        yield f"box_open_c{col}", abi.default_config(collision=col, root_size=10.0, boundary=abi.BOUNDARY_OPEN), q
        yield f"box_per_c{col}", abi.default_config(collision=col, root_size=10.0, boundary=abi.BOUNDARY_PERIODIC,
                                                    N_ghost_x=1, N_ghost_y=1, N_ghost_z=1), q
        yield f"box_2root_c{col}", abi.default_config(collision=col, root_size=5.0, N_root_x=2, N_root_y=2, N_root_z=2,
                                                      boundary=abi.BOUNDARY_PERIODIC, N_ghost_x=2, N_ghost_y=1), q
    e = abi.particles(3)
    e["x"] = [0.0, 3.0, -3.0]
    e["r"] = 0.1
    yield "no_collisions_tree", abi.default_config(collision=abi.COLLISION_TREE, root_size=10.0), e
    yield "no_collisions_direct", abi.default_config(collision=abi.COLLISION_DIRECT, root_size=10.0), e


COL_CASES = list(collision_cases())


@pytest.mark.parametrize("name,cfg,p", COL_CASES, ids=[c[0] for c in COL_CASES])
def test_collision_list_bitwise(eng, name, cfg, p):
    want = checkers.oracle().collision_search(cfg, p)
    got = eng.collision_search_host(cfg.copy(), np.ascontiguousarray(p))
    assert len(got) == len(want)
    assert collisions_equal(got, want, with_ri=(cfg.collision == abi.COLLISION_TREE))


def test_collision_golden(eng):
    g = np.load(os.path.join(GOLD, "sheet_root30.npz"))
    p = np.frombuffer(g["particles_in"].tobytes(), dtype=abi.PARTICLE_DTYPE).copy()
    cfg = ics.shearing_sheet_config(root_size=30.0, t=55.5)
    got = eng.collision_search_host(cfg.copy(), p)
    want = np.frombuffer(g["col_tree"].tobytes(), dtype=abi.COLLISION_DTYPE)
    assert collisions_equal(got, want)
    cfg_d = ics.shearing_sheet_config(root_size=30.0, t=55.5, collision=abi.COLLISION_DIRECT)
    got = eng.collision_search_host(cfg_d.copy(), p)
    want = np.frombuffer(g["col_direct"].tobytes(), dtype=abi.COLLISION_DTYPE)
    assert collisions_equal(got, want, with_ri=False)
    q, c = p.copy(), cfg.copy()
    n = eng.gravity_host(c, q)
    got = np.stack([q["ax"][:n], q["ay"][:n], q["az"][:n]], 1)
    assert np.array_equal(got.view(np.uint64), g["acc"].view(np.uint64))


def test_sei_step_bitwise(eng):
    p = ics.shearing_sheet(root_size=30.0, seed=8)
    cfg = ics.shearing_sheet_config(root_size=30.0, collision=abi.COLLISION_NONE)
    want, cw = checkers.oracle().integrator_step(cfg, p)
    eng.upload(p.copy())
    c = cfg.copy()
    eng.integrator_step(c)
    got = eng.download()
    assert bits_equal(got, want)
    assert c.t == cw.t and c.OMEGAZ == cw.OMEGAZ and c.dt_last_done == cw.dt_last_done


def test_disc_full_steps_bitwise(eng):
    """Config C4 shape: leapfrog + tree gravity + open boundary (particles leave the box)."""
    p = ics.selfgravity_disc(4000, seed=12)
    cfg = ics.selfgravity_disc_config(collision=abi.COLLISION_NONE)
    want, cw, _ = checkers.oracle().steps(cfg, p, 5)
    q, c = p.copy(), cfg.copy()
    n = eng.steps_host(c, q, 5)
    assert n == len(want)
    assert bits_equal(q[:n], want)
    assert c.t == cw.t


def test_sheet_steps_no_resolve_bitwise(eng):
    """Config C5 shape without the resolve loop: SEI + shear boundary + tree gravity (25 ghost boxes)
    + tree collision search every step."""
    p = ics.shearing_sheet(root_size=30.0, seed=9)
    cfg = ics.shearing_sheet_config(root_size=30.0)
    want, cw, _ = checkers.oracle().steps(cfg, p, 6, resolve=0)
    q, c = p.copy(), cfg.copy()
    n = eng.steps_host(c, q, 6)
    assert bits_equal(q[:n], want)
    want_col = checkers.oracle().collision_search(cw, want)
    got_col = eng.collisions_fetch()
    assert collisions_equal(got_col, want_col)


def test_disc_2pow20_full_size_bitwise(eng):
    """Config C4 recipe at N=2^20 (the largest size the CPU oracle finishes in seconds with OpenMP): the whole
    acceleration array is bit-identical, and the device tree satisfies the pre-order invariants."""
    n = 1 << 20
    p = ics.selfgravity_disc(n - 1, seed=42)
    cfg = ics.selfgravity_disc_config()
    want, _ = checkers.oracle().gravity(cfg, p)
    q, c = p.copy(), cfg.copy()
    m = eng.gravity_host(c, q)
    assert m == len(want)
    assert bits_equal(q[:m], want)
    cells = eng.tree(c)
    leaves = cells[cells["pt"] >= 0]
    assert len(leaves) == m and np.array_equal(np.sort(leaves["pt"]), np.arange(m))      # one leaf per particle
    internal = cells[cells["pt"] < 0]
    assert -internal["pt"][0] == m                                                          # root counts everything
    assert np.all(cells["skip"] > np.arange(len(cells))) and cells["skip"][0] == len(cells)
    assert abs(cells["m"][0] - p["m"].sum()) < 1e-12


def test_empty_and_tiny_inputs(eng):
    """N = 0 and N = 1 through every entry point (the reference's loops simply do nothing)."""
    for n in (0, 1):
        p = abi.particles(n)
        if n:
            p["x"], p["m"], p["r"] = 0.3, 1.0, 0.1
        for cfg in (ics.plummer_config(8), ics.plummer_config(8, gravity=abi.GRAVITY_COMPENSATED),
                    ics.selfgravity_disc_config(collision=abi.COLLISION_TREE),
                    ics.shearing_sheet_config(root_size=10.0), abi.default_config(collision=abi.COLLISION_DIRECT, root_size=5.0)):
            q, c = p.copy(), cfg.copy()
            m = eng.steps_host(c, q, 2)
            want, cw, _ = checkers.oracle().steps(cfg, p, 2)
            assert m == len(want) == n and bits_equal(q[:m], want) and c.t == cw.t
            assert len(eng.collision_search_host(cfg.copy(), p.copy())) == 0


def test_line_and_linetree_collision_lists_bitwise(eng):
    """REB_COLLISION_LINE / LINETREE on the GPU against the oracle (same cases as the oracle-vs-reference pin)."""
    from test_oracle_vs_reference import LINE_CASES
    for name, cfg, p in LINE_CASES:
        want = checkers.oracle().collision_search(cfg, p)
        got = eng.collision_search_host(cfg.copy(), np.ascontiguousarray(p))
        assert len(want) > 0 and len(got) == len(want), name
        assert collisions_equal(got, want, with_ri=(cfg.collision == abi.COLLISION_LINETREE)), name


def test_collision_subset_lists_bitwise(eng):
    """r->map / r->N_map / r->N_targets subsets (collision.c:53-58) on the GPU against the oracle, all four modes;
    the subset is cleared again afterwards and invalid subsets are refused."""
    from test_oracle_vs_reference import _subset_cases
    try:
        n_nonempty = 0
        for name, cfg, p, sub, nt in _subset_cases():
            want = checkers.oracle().collision_search_subset(cfg, p, sub, nt)
            eng.set_collision_subset(sub, nt)
            got = eng.collision_search_host(cfg.copy(), np.ascontiguousarray(p))
            tree = cfg.collision in (abi.COLLISION_TREE, abi.COLLISION_LINETREE)
            assert len(got) == len(want), name
            assert collisions_equal(got, want, with_ri=tree), name
            n_nonempty += len(want) > 0
        assert n_nonempty >= 20
        name, cfg, p, sub, nt = _subset_cases()[0]
        eng.set_collision_subset(np.array([0, len(p)], dtype=np.uint64))           # entry >= N
        with pytest.raises(ReboundCudaError):
            eng.collision_search_host(cfg.copy(), np.ascontiguousarray(p))
        eng.set_collision_subset(None, len(p) + 1)                                 # more targets than projectiles
        with pytest.raises(ReboundCudaError):
            eng.collision_search_host(cfg.copy(), np.ascontiguousarray(p))
    finally:
        eng.set_collision_subset()
    name, cfg, p, sub, nt = _subset_cases()[0]
    assert collisions_equal(eng.collision_search_host(cfg.copy(), np.ascontiguousarray(p)),
                            checkers.oracle().collision_search(cfg, p), with_ri=False)


def _pair_keys(col):
    """(p1, p2, ghost shift) of every entry as sortable rows."""
    k = np.stack([col["p1"].astype(np.float64), col["p2"].astype(np.float64), col["gb_x"], col["gb_y"], col["gb_vy"]], axis=1)
    return k[np.lexsort(k.T[::-1])]


def test_sheet_full_size_tree_and_direct_searches_agree(eng):
    """Config C5 recipe at N ~ 2^18 (beyond what the CPU oracle does in seconds): a size-independent property --
    the tree search and the all-pairs search report the same set of (p1, p2, ghost box) entries, every entry
    appears in both directions unless a ghost shift separates them, and the lists are reproducible."""
    root = 1327.5                       # 2x2 root boxes of this size hold ~2^18 particles (ics.shearing_sheet)
    p = ics.shearing_sheet(root_size=root, seed=3)
    assert 200_000 < len(p) < 330_000
    cfg_t = ics.shearing_sheet_config(root_size=root, t=12.5)
    cfg_d = ics.shearing_sheet_config(root_size=root, t=12.5, collision=abi.COLLISION_DIRECT)
    tree = eng.collision_search_host(cfg_t.copy(), p.copy())
    direct = eng.collision_search_host(cfg_d.copy(), p.copy())
    assert len(tree) == len(direct) > 1000
    assert np.array_equal(_pair_keys(tree), _pair_keys(direct))
    again = eng.collision_search_host(cfg_t.copy(), p.copy())
    assert tree.tobytes() == again.tobytes()
    # inside the main box (zero shift) overlap is symmetric: (i, j) present <=> (j, i) present
    main = tree[(tree["gb_x"] == 0) & (tree["gb_y"] == 0)]
    fwd = set(zip(main["p1"].tolist(), main["p2"].tolist()))
    assert all((j, i) in fwd for i, j in fwd)


def _box_particles(n, half, seed):
    rng = np.random.default_rng(seed)
    p = abi.particles(n)
    for f in ("x", "y", "z"):
        p[f] = rng.uniform(-half, half, n)
    p["m"] = rng.uniform(0.5, 1.5, n) / n
    return p


def quadrupole_cases():
    yield "disc", ics.selfgravity_disc_config(collision=abi.COLLISION_NONE), ics.selfgravity_disc(6000, seed=3)
    yield "sheet_ghosts", ics.shearing_sheet_config(root_size=30.0, t=3.3, collision=abi.COLLISION_NONE), ics.shearing_sheet(root_size=30.0, seed=4)
    yield "plummer_wide", ics.plummer_config(800, gravity=abi.GRAVITY_TREE, root_size=200.0, opening_angle2=1.0), ics.plummer(800, seed=5)
    yield "two_root_boxes", abi.default_config(gravity=abi.GRAVITY_TREE, root_size=5.0, N_root_x=2, N_root_y=2, N_root_z=2,
                                                opening_angle2=0.7, softening=0.01), _box_particles(900, 4.9, seed=6)


QUAD_CASES = list(quadrupole_cases())


@pytest.mark.parametrize("name,cfg,p", QUAD_CASES, ids=[c[0] for c in QUAD_CASES])
def test_quadrupole_tree_gravity_bitwise(eng, name, cfg, p):
    """cfg.quadrupole = 1 (a reference compiled with -DQUADRUPOLE, src/tree.c:148-198, 293-303): the oracle is pinned
    against that build in tests/test_oracle_vs_reference.py; STRICT accelerations are bit-identical, FAST within 1e-11."""
    cq = cfg.copy()
    cq.quadrupole = 1
    want, _ = checkers.oracle().gravity(cq, p)
    q = p.copy()
    n = eng.gravity_host(cq.copy(), q)
    assert n == len(want) and bits_equal(q[:n], want)
    mono = p.copy()
    eng.gravity_host(cfg.copy(), mono)
    assert not bits_equal(mono[:n], want)                    # the option is really on
    cf = cq.copy()
    cf.mode = abi.MODE_FAST
    f = p.copy()
    eng.gravity_host(cf, f)
    assert max_rel_acc_error(f[:n], want) <= 1e-11


def test_walk_variants_give_identical_bits(tmp_path):
    """REBOUND_B200_WALK = v1 / coop (the A/B kernels kept beside the default records walk) reproduce the default
    walk's accelerations bit for bit; the variant is latched per process, hence the subprocesses."""
    import subprocess
    import sys
    script = tmp_path / "walk_variant.py"
    script.write_text(
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
        "from rebound_b200 import abi, ics\n"
        "from rebound_b200.simulation import Engine\n"
        "eng = Engine(0)\n"
        "out = []\n"
        "for cfg, p in ((ics.selfgravity_disc_config(), ics.selfgravity_disc(20000, seed=2)),\n"
        "               (ics.shearing_sheet_config(root_size=60.0, t=12.3), ics.shearing_sheet(root_size=60.0, seed=5))):\n"
        "    for mode in (abi.MODE_STRICT, abi.MODE_FAST):\n"
        "        c = cfg.copy(); c.mode = mode\n"
        "        q = np.ascontiguousarray(p.copy())\n"
        "        eng.gravity_host(c, q)\n"
        "        out += [q['ax'], q['ay'], q['az']]\n"
        "np.concatenate(out).tofile(sys.argv[1])\n")
    res = {}
    for variant in ("", "rec", "v1", "coop"):
        env = dict(os.environ)
        env["REBOUND_B200_WALK"] = variant
        out = tmp_path / f"acc_{variant or 'default'}.bin"
        r = subprocess.run([sys.executable, str(script), str(out)], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        res[variant] = np.fromfile(out, dtype=np.uint64)
    assert len(res["rec"]) > 0
    assert np.array_equal(res["rec"], res["v1"])
    assert np.array_equal(res["rec"], res["coop"])
    # the default differs from "rec" only in FAST mode (group walk): the STRICT halves (first 3 of every 6 arrays) agree
    n1, n2 = 20001, (len(res["rec"]) - 6 * 20001) // 6
    strict = np.concatenate([np.arange(0, 3 * n1), 6 * n1 + np.arange(0, 3 * n2)])
    assert np.array_equal(res[""][strict], res["rec"][strict])
    assert not np.array_equal(res[""], res["rec"])


def test_extras_golden(eng):
    """The committed reference vectors of tests/golden/extras300.npz through the C ABI: LINE / LINETREE lists, r->map /
    N_targets subsets in all four search modes, the jerk kick, the exit checks (no oracle involved)."""
    import sys
    sys.path.insert(0, GOLD)
    from make_golden import extras_inputs
    g = np.load(os.path.join(GOLD, "extras300.npz"))
    q, sub, nt, base = extras_inputs()
    q = np.ascontiguousarray(q)
    try:
        for mode in (abi.COLLISION_DIRECT, abi.COLLISION_TREE, abi.COLLISION_LINE, abi.COLLISION_LINETREE):
            c = abi.default_config(collision=mode, **base)
            tree = mode in (abi.COLLISION_TREE, abi.COLLISION_LINETREE)
            for key, m, t in ((f"col_m{mode}", None, None), (f"col_m{mode}_map", sub, None),
                              (f"col_m{mode}_map_targets", sub, nt), (f"col_m{mode}_targets", None, nt)):
                eng.set_collision_subset(m, t)
                got = eng.collision_search_host(c.copy(), q.copy())
                want = np.frombuffer(g[key].tobytes(), dtype=abi.COLLISION_DTYPE)
                assert collisions_equal(got, want, with_ri=tree), key
    finally:
        eng.set_collision_subset()
    qa = np.frombuffer(g["jerk_in"].tobytes(), dtype=abi.PARTICLE_DTYPE).copy()
    cj = abi.default_config(softening=0.05, N_active=100, testparticle_type=1)
    out = qa.copy()
    eng.jerk_host(cj.copy(), out, 0.37)
    assert bits_equal(out, np.frombuffer(g["jerk_out"].tobytes(), dtype=abi.PARTICLE_DTYPE))
    cj2 = abi.default_config(softening=0.05, gravity_ignore_terms=abi.IGNORE_TERMS_INVOLVING_0)
    out = qa.copy()
    eng.jerk_host(cj2.copy(), out, -0.11)
    assert bits_equal(out, np.frombuffer(g["jerk_out_ignore0"].tobytes(), dtype=abi.PARTICLE_DTYPE))
    eng.upload(q)
    for i, mx in enumerate(g["exit_max"]):
        for j, mn in enumerate(g["exit_min"]):
            escape, encounter = eng.exit_check(float(mx), float(mn))
            assert (3 if encounter else (4 if escape else 0)) == int(g["exit_status"][i, j])


def test_key_prefix_builds_give_identical_trees(tmp_path):
    """The build sorts on a key prefix (fewer radix passes) and orders the particles that share it by exact pairwise
    descent; a prefix that is too short for the input aborts the build and falls back to full keys.  Forced prefixes
    of 2, 5 and 9 levels (REBOUND_B200_KEY_LEVELS, latched per process) must reproduce the full-key tree and
    accelerations bit for bit: 2 levels trips the fallback, 5 and 9 exercise long and short tie runs."""
    import subprocess
    import sys
    script = tmp_path / "prefix.py"
    script.write_text(
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
        "from rebound_b200 import abi, ics\n"
        "from rebound_b200.simulation import Engine\n"
        "eng = Engine(0)\n"
        "out = []\n"
        "p3 = ics.plummer(6000, seed=4)\n"
        "cases = [(ics.selfgravity_disc_config(), ics.selfgravity_disc(30000, seed=2)),\n"
        "         (ics.shearing_sheet_config(root_size=60.0, t=12.3), ics.shearing_sheet(root_size=60.0, seed=5)),\n"
        "         (ics.plummer_config(6000, gravity=abi.GRAVITY_TREE, root_size=200.0, opening_angle2=0.25), p3),\n"
        "         (ics.plummer_config(6000, gravity=abi.GRAVITY_TREE, root_size=50.0, N_root_x=2, N_root_y=3, N_root_z=2,\n"
        "                             boundary=abi.BOUNDARY_PERIODIC, N_ghost_x=1, N_ghost_y=1, N_ghost_z=1), p3)]\n"
        "for cfg, p in cases:\n"
        "    q = np.ascontiguousarray(p.copy())\n"
        "    n = eng.gravity_host(cfg.copy(), q)\n"
        "    out += [np.frombuffer(q[:n].tobytes(), dtype=np.uint8)]\n"
        "    eng.upload(np.ascontiguousarray(q[:n]))\n"
        "    out += [np.frombuffer(eng.tree(cfg.copy()).tobytes(), dtype=np.uint8)]\n"
        "    col = eng.collision_search_host(cfg.copy(), np.ascontiguousarray(q[:n]))\n"
        "    out += [np.frombuffer(col.tobytes(), dtype=np.uint8)]\n"
        "np.concatenate(out).tofile(sys.argv[1])\n")
    res = {}
    for levels in ("0", "2", "5", "9"):
        env = dict(os.environ)
        env["REBOUND_B200_KEY_LEVELS"] = levels
        out = tmp_path / f"tree_{levels}.bin"
        r = subprocess.run([sys.executable, str(script), str(out)], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        res[levels] = np.fromfile(out, dtype=np.uint8)
    assert len(res["0"]) > 1000000
    for levels in ("2", "5", "9"):
        assert np.array_equal(res["0"], res[levels]), levels
