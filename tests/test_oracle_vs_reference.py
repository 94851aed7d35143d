"""Pins the CPU restatement (oracle/oracle.c) against the UNMODIFIED reference compiled from
/root/reference/src into oracle/_ref (oracle/Makefile).  Everything is compared bit for bit.
The same cases are stored as fixtures (tests/golden) for machines without the reference build."""
import numpy as np
import pytest

import checkers
from checkers import bits_equal
from rebound_b200 import abi, ics

pytestmark = pytest.mark.needs_ref


def cases_direct():
    p = ics.plummer(300, seed=1)
    yield "plummer_basic", ics.plummer_config(300), p
    yield "plummer_comp", ics.plummer_config(300, gravity=abi.GRAVITY_COMPENSATED), p
    yield "nosoft", ics.plummer_config(300, softening=0.0), p
    for typ in (0, 1):
        for grav in (abi.GRAVITY_BASIC, abi.GRAVITY_COMPENSATED):
            q = ics.planetesimal_disk(200, seed=3)
            q["m"][10:] = 1e-9
            yield f"testp{typ}_g{grav}", ics.planetesimal_config(testparticle_type=typ, gravity=grav), q
    for terms in (1, 2):
        for grav in (abi.GRAVITY_BASIC, abi.GRAVITY_COMPENSATED):
            yield f"ignore{terms}_g{grav}", ics.plummer_config(300, gravity_ignore_terms=terms, gravity=grav), p
    # N_active = 1 with type 1: exercises the (j==1 && i==0) corner of gravity.c:263
    q = ics.planetesimal_disk(50, seed=4)
    q["m"][1:] = 1e-6
    yield "nactive1_ignore1", ics.planetesimal_config(N_active=1, testparticle_type=1, gravity_ignore_terms=1), q


@pytest.mark.parametrize("name,cfg,p", list(cases_direct()), ids=lambda v: v if isinstance(v, str) else "")
def test_direct_gravity_bitwise(name, cfg, p):
    ref, _ = checkers.reference().gravity(cfg, p)
    refomp, _ = checkers.reference(openmp=True).gravity(cfg, p)
    orc, _ = checkers.oracle().gravity(cfg, p)
    assert bits_equal(ref, refomp), "reference serial and OpenMP builds differ"
    assert bits_equal(orc, ref)


@pytest.mark.parametrize("name,cfg,p", list(cases_direct()), ids=lambda v: v if isinstance(v, str) else "")
def test_direct_gravity_rows_equal_the_reference(name, cfg, p):
    """orc_gravity_rows (the oracle's row-sampled direct sum used for the full-size spot checks of C3) gives the
    reference's bits for every row it is asked for."""
    ref, _ = checkers.reference().gravity(cfg, p)
    rows = np.array([0, 1, 2, len(p) // 2, len(p) - 1], dtype=np.uint64)
    got = checkers.oracle().gravity_rows(cfg, p, rows)
    want = np.stack([ref["ax"], ref["ay"], ref["az"]], 1)[rows.astype(np.int64)]
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_tree_session_samples_equal_the_full_evaluation():
    """The reference harness's sampled tree walk (bench.py's bounded CPU sample and the full-size spot checks of C4)
    leaves, on the sampled particles, exactly the accelerations of reb_gravity_tree_calculate_acceleration."""
    cfg = ics.selfgravity_disc_config()
    p = ics.selfgravity_disc(3000, seed=5)
    full, _ = checkers.reference().gravity(cfg, p)
    for chk in (checkers.reference(), checkers.reference(openmp=True)):
        s = chk.tree_session(cfg, p)
        assert s.N == len(full)
        sec, n = s.walk_sample(7, 3)
        assert n == len(range(3, s.N, 7)) and sec >= 0
        got = s.sample_acc(7, 3)
        t = s.close()
        want = np.stack([full["ax"], full["ay"], full["az"]], 1)[3::7]
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
        assert set(t) == {"boundary", "construct", "gravity_data", "delete", "rest"}


def test_basic_ghostboxes_matches_openmp_build():
    # With ghost boxes the reference's serial build applies the shifted pair antisymmetrically
    # (gravity.c:199-212) and is NOT bitwise equal to its own OpenMP build; the gather form is the
    # OpenMP build's (gravity.c:216-232).
    p = ics.plummer(200, seed=5)
    for b in (abi.BOUNDARY_PERIODIC, abi.BOUNDARY_OPEN, abi.BOUNDARY_SHEAR):
        cfg = ics.plummer_config(200, boundary=b, root_size=30.0, N_ghost_x=1, N_ghost_y=2, N_ghost_z=1,
                                 OMEGA=1.0, t=0.37)
        refomp, _ = checkers.reference(openmp=True).gravity(cfg, p)
        ref, _ = checkers.reference().gravity(cfg, p)
        orc, _ = checkers.oracle().gravity(cfg, p)
        assert bits_equal(orc, refomp)
        assert checkers.max_rel_acc_error(ref, refomp) < 1e-12


@pytest.mark.parametrize("order", [2, 4, 6, 8])
def test_leapfrog_steps_bitwise(order):
    p = ics.plummer(128, seed=7)
    cfg = ics.plummer_config(128, leapfrog_order=order, dt=1e-3)
    ref, cr, _ = checkers.reference().steps(cfg, p, 5)
    orc, co, _ = checkers.oracle().steps(cfg, p, 5)
    assert bits_equal(orc, ref)
    assert cr.t == co.t and cr.dt_last_done == co.dt_last_done


def test_leapfrog_bad_order():
    p = ics.plummer(8, seed=7)
    cfg = ics.plummer_config(8, leapfrog_order=3)
    with pytest.raises(checkers.CheckerError) as e1:
        checkers.reference().integrator_step(cfg, p)
    with pytest.raises(checkers.CheckerError) as e2:
        checkers.oracle().integrator_step(cfg, p)
    assert e1.value.msg == e2.value.msg == "Leapfrog order not supported."


def tree_cases():
    yield "disc", ics.selfgravity_disc_config(), ics.selfgravity_disc(2000, seed=2)
    yield "disc_theta1.5", ics.selfgravity_disc_config(opening_angle2=1.5), ics.selfgravity_disc(500, seed=3)
    p = ics.plummer(1000, seed=4)
    yield "plummer_box", ics.plummer_config(1000, gravity=abi.GRAVITY_TREE, root_size=200.0, opening_angle2=0.25), p
    yield "sheet", ics.shearing_sheet_config(root_size=40.0, t=123.4), ics.shearing_sheet(root_size=40.0, seed=5)
    yield "rootboxes_3d", ics.plummer_config(1000, gravity=abi.GRAVITY_TREE, root_size=50.0, N_root_x=2,
                                            N_root_y=3, N_root_z=2, boundary=abi.BOUNDARY_PERIODIC,
                                            N_ghost_x=1, N_ghost_y=1, N_ghost_z=1), p


@pytest.mark.parametrize("name,cfg,p", list(tree_cases()), ids=lambda v: v if isinstance(v, str) else "")
def test_tree_cells_bitwise(name, cfg, p):
    p, cfg = checkers.reference().boundary_check(cfg, p)
    ref = checkers.reference().tree_dump(cfg, p)
    orc = checkers.oracle().tree_dump(cfg, p)
    assert len(ref) == len(orc)
    assert ref.tobytes() == orc.tobytes()


@pytest.mark.parametrize("name,cfg,p", list(tree_cases()), ids=lambda v: v if isinstance(v, str) else "")
def test_tree_gravity_bitwise(name, cfg, p):
    ref, cr = checkers.reference().gravity(cfg, p)
    orc, co = checkers.oracle().gravity(cfg, p)
    assert len(ref) == len(orc)
    assert bits_equal(orc, ref)
    assert cr.N_active == co.N_active


def test_tree_errors():
    cfg = ics.selfgravity_disc_config()
    p = ics.selfgravity_disc(50, seed=2)
    bad = p.copy(); bad["x"][7] = bad["x"][3]; bad["y"][7] = bad["y"][3]; bad["z"][7] = bad["z"][3]
    nan = p.copy(); nan["y"][5] = np.nan
    noroot = ics.selfgravity_disc_config(root_size=-1.0)
    outside = p.copy(); outside["x"][9] = 100.0
    cfg_nob = ics.selfgravity_disc_config(boundary=abi.BOUNDARY_NONE)
    for c, q in ((cfg, bad), (cfg, nan), (noroot, p), (cfg_nob, outside)):
        with pytest.raises(checkers.CheckerError) as e1:
            checkers.reference().tree_dump(c, q)
        with pytest.raises(checkers.CheckerError) as e2:
            checkers.oracle().tree_dump(c, q)
        assert e1.value.msg == e2.value.msg


@pytest.mark.parametrize("boundary", [abi.BOUNDARY_OPEN, abi.BOUNDARY_PERIODIC, abi.BOUNDARY_SHEAR])
def test_boundary_bitwise(boundary):
    rng = np.random.default_rng(11)
    n = 500
    p = abi.particles(n)
    for f in ("x", "y", "z"):
        p[f] = rng.uniform(-14, 14, n)
    for f in ("vx", "vy", "vz"):
        p[f] = rng.normal(0, 1, n)
    p["m"] = 1.0
    cfg = abi.default_config(boundary=boundary, root_size=10.0, N_root_x=2, N_root_y=1, N_root_z=1,
                             OMEGA=0.7, t=3.3, N_active=40)
    ref, cr = checkers.reference().boundary_check(cfg, p)
    orc, co = checkers.oracle().boundary_check(cfg, p)
    assert len(ref) == len(orc)
    assert bits_equal(orc, ref)
    assert cr.N_active == co.N_active


def test_boundary_open_removes_everything():
    p = abi.particles(5)
    p["x"] = 100.0
    cfg = abi.default_config(boundary=abi.BOUNDARY_OPEN, root_size=10.0, N_active=3)
    ref, cr = checkers.reference().boundary_check(cfg, p)
    orc, co = checkers.oracle().boundary_check(cfg, p)
    assert len(ref) == len(orc) == 0
    assert cr.N_active == co.N_active


def collision_cases():
    p = ics.shearing_sheet(root_size=40.0, seed=5)
    yield "sheet_tree", ics.shearing_sheet_config(root_size=40.0, t=55.5), p
    yield "sheet_direct", ics.shearing_sheet_config(root_size=40.0, t=55.5, collision=abi.COLLISION_DIRECT), p
    rng = np.random.default_rng(3)
    n = 400
    q = abi.particles(n)
    for f in ("x", "y", "z"):
        q[f] = rng.uniform(-4.9, 4.9, n)
    for f in ("vx", "vy", "vz"):
        q[f] = rng.normal(0, 1, n)
    q["r"] = rng.uniform(0.05, 0.4, n)
    q["m"] = 1.0
    for col in (abi.COLLISION_DIRECT, abi.COLLISION_TREE):
        yield f"box_open_c{col}", abi.default_config(collision=col, root_size=10.0, boundary=abi.BOUNDARY_OPEN), q
        yield f"box_per_c{col}", abi.default_config(collision=col, root_size=10.0, boundary=abi.BOUNDARY_PERIODIC,
                                                    N_ghost_x=1, N_ghost_y=1, N_ghost_z=1), q
        yield f"box_2root_c{col}", abi.default_config(collision=col, root_size=5.0, N_root_x=2, N_root_y=2, N_root_z=2,
                                                      boundary=abi.BOUNDARY_PERIODIC, N_ghost_x=2, N_ghost_y=1), q


@pytest.mark.parametrize("name,cfg,p", list(collision_cases()), ids=lambda v: v if isinstance(v, str) else "")
def test_collision_list_bitwise(name, cfg, p):
    ref = checkers.reference().collision_search(cfg, p)
    orc = checkers.oracle().collision_search(cfg, p)
    assert len(ref) > 0
    assert len(ref) == len(orc)
    assert checkers.collisions_equal(ref, orc, with_ri=(cfg.collision == abi.COLLISION_TREE))


def subset_cases():
    """r->map / r->N_map / r->N_targets (collision.c:53-58): what MERCURIUS and TRACE set around encounter steps."""
    base = {name: (cfg, p) for name, cfg, p in collision_cases()}
    base.update({name: (cfg, p) for name, cfg, p in line_cases()})
    rng = np.random.default_rng(17)
    for name in ("box_open_c1", "box_per_c1", "box_open_c2", "box_2root_c2", "open_c4", "per_c4", "open_c5", "sheet_direct"):
        cfg, p = base[name]
        n = len(p)
        sub = rng.permutation(n)[: (2 * n) // 3]
        yield f"{name}_map", cfg, p, sub, None
        yield f"{name}_map_targets", cfg, p, sub, len(sub) // 4
        yield f"{name}_targets", cfg, p, None, n // 5
    cfg, p = base["box_open_c1"]
    yield "empty_map", cfg, p, np.zeros(0, dtype=np.uint64), None
    yield "zero_targets", cfg, p, None, 0
    yield "repeated_slots", cfg, p, np.array([5, 5, 7, 9, 7], dtype=np.uint64), None


SUBSET_CASES = None


def _subset_cases():
    global SUBSET_CASES
    if SUBSET_CASES is None:
        SUBSET_CASES = list(subset_cases())
    return SUBSET_CASES


def test_collision_subset_lists_bitwise():
    n_nonempty = 0
    for name, cfg, p, sub, nt in _subset_cases():
        ref = checkers.reference().collision_search_subset(cfg, p, sub, nt)
        orc = checkers.oracle().collision_search_subset(cfg, p, sub, nt)
        tree = cfg.collision in (abi.COLLISION_TREE, abi.COLLISION_LINETREE)
        assert checkers.collisions_equal(ref, orc, with_ri=tree), name
        n_nonempty += len(ref) > 0
    assert n_nonempty >= 20


def jerk_cases():
    """reb_gravity_basic_calculate_and_apply_jerk (gravity.c:850-924): state = positions, velocities and the
    accelerations of a preceding force evaluation."""
    def with_acc(cfg, p):
        q, _ = checkers.oracle().gravity(cfg, p)
        return q
    p = ics.plummer(700, seed=12)
    cfg = ics.plummer_config(700)
    yield "plummer", cfg, with_acc(cfg, p), 0.37
    for terms in (abi.IGNORE_TERMS_BETWEEN_0_AND_1, abi.IGNORE_TERMS_INVOLVING_0):
        c2 = ics.plummer_config(300, gravity_ignore_terms=terms)
        yield f"plummer_ignore{terms}", c2, with_acc(c2, p[:300]), -0.21
    q = ics.planetesimal_disk(600, seed=13)
    q["m"][10:] = 1e-9
    for tp in (0, 1):
        c3 = ics.planetesimal_config(testparticle_type=tp)
        yield f"testp_type{tp}", c3, with_acc(c3, q), 1e-3
    c4 = ics.planetesimal_config(testparticle_type=1, N_active=0)
    yield "no_active", c4, with_acc(ics.planetesimal_config(testparticle_type=1), q[:100]), 0.5
    yield "two", cfg, with_acc(cfg, p[:2]), 0.1
    yield "one", cfg, with_acc(cfg, p[:1]), 0.1


JERK_CASES = None


def _jerk_cases():
    global JERK_CASES
    if JERK_CASES is None:
        JERK_CASES = list(jerk_cases())
    return JERK_CASES


def test_apply_jerk_bitwise():
    for name, cfg, p, v in _jerk_cases():
        ref = checkers.reference().apply_jerk(cfg, p, v)
        orc = checkers.oracle().apply_jerk(cfg, p, v)
        assert checkers.bits_equal(ref, orc), name
        if len(p) > 1:
            assert not checkers.bits_equal(orc, p, fields=("vx", "vy", "vz")), name
        # the gather formulation the CUDA kernel uses (one ascending sum per particle) gives the same bits
        import ctypes as C
        g = p.copy()
        fn = checkers.oracle().lib.orc_apply_jerk_gather
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(abi.Config), C.c_void_p, C.c_uint64, C.c_double]
        assert fn(C.byref(cfg.copy()), abi.as_ptr(g), len(g), float(v)) == 0
        assert checkers.bits_equal(ref, g), name + " (gather)"


def exit_cases():
    """(particles, exit_max_distance, exit_min_distance) incl. thresholds that sit exactly on a particle / a pair."""
    rng = np.random.default_rng(23)
    n = 300
    q = abi.particles(n)
    for f in ("x", "y", "z"):
        q[f] = rng.normal(0, 1, n)
    q["m"] = 1.0 / n
    r = np.sqrt(q["x"] * q["x"] + q["y"] * q["y"] + q["z"] * q["z"])
    d = np.sqrt((q["x"][:, None] - q["x"][None, :]) ** 2 + (q["y"][:, None] - q["y"][None, :]) ** 2
                + (q["z"][:, None] - q["z"][None, :]) ** 2)
    dmin = d[np.triu_indices(n, 1)].min()
    for mx in (0.0, r.max() * 0.999, float(r.max()), np.nextafter(r.max(), 10.0), r.max() * 2):
        for mn in (0.0, dmin * 0.999, float(dmin), np.nextafter(dmin, 10.0), dmin * 1.5):
            yield q, float(mx), float(mn)
    yield q[:1], 0.5, 10.0
    yield q[:0], 0.5, 10.0
    w = q.copy()
    w["x"][7] = np.nan
    yield w, 3.0, 1e-3


def test_exit_checks_match_reference():
    seen = set()
    for q, mx, mn in exit_cases():
        ref = checkers.reference().exit_check(abi.default_config(), q, mx, mn)
        orc = checkers.oracle().exit_check(abi.default_config(), q, mx, mn)
        if len(q) == 0:
            assert ref == 2 and orc == 0       # REB_STATUS_NO_PARTICLES: the reference leaves before any check
            continue
        assert ref == orc, (mx, mn)
        seen.add(ref)
    assert seen == {0, 3, 4}


def test_sei_step_bitwise():
    p = ics.shearing_sheet(root_size=30.0, seed=8)
    cfg = ics.shearing_sheet_config(root_size=30.0, collision=abi.COLLISION_NONE)
    ref, cr = checkers.reference().integrator_step(cfg, p)
    orc, co = checkers.oracle().integrator_step(cfg, p)
    assert bits_equal(orc, ref)
    assert cr.t == co.t and cr.OMEGAZ == co.OMEGAZ


@pytest.mark.parametrize("resolve", [1, 2])
def test_shearing_sheet_full_steps_bitwise(resolve):
    # examples/shearing_sheet: SEI + shear boundary + tree gravity + tree collisions + hard spheres
    p = ics.shearing_sheet(root_size=30.0, seed=9)
    cfg = ics.shearing_sheet_config(root_size=30.0)
    mcv = 1.0 * ics.SHEET_OMEGA * 0.001
    ref, cr, ar = checkers.reference().steps(cfg, p, 20, resolve=resolve, minimum_collision_velocity=mcv)
    orc, co, ao = checkers.oracle().steps(cfg, p, 20, resolve=resolve, minimum_collision_velocity=mcv)
    assert ar["collisions_log_n"] > 0
    assert ar["collisions_log_n"] == ao["collisions_log_n"]
    assert ar["collisions_plog"] == ao["collisions_plog"]
    assert bits_equal(orc, ref)


def test_disc_full_steps_bitwise():
    p = ics.selfgravity_disc(1500, seed=12)
    cfg = ics.selfgravity_disc_config(collision=abi.COLLISION_NONE)
    ref, cr, _ = checkers.reference().steps(cfg, p, 5)
    orc, co, _ = checkers.oracle().steps(cfg, p, 5)
    assert len(ref) == len(orc)
    assert bits_equal(orc, ref)


def test_energy():
    p = ics.plummer(200, seed=1)
    cfg = ics.plummer_config(200)
    assert checkers.reference().energy(cfg, p) == checkers.oracle().energy(cfg, p)


def line_cases():
    """REB_COLLISION_LINE / LINETREE (collision.c:125-196, 270-331): trajectories over the last step."""
    rng = np.random.default_rng(3)
    n = 400
    q = abi.particles(n)
    for f in ("x", "y", "z"):
        q[f] = rng.uniform(-4.9, 4.9, n)
    for f in ("vx", "vy", "vz"):
        q[f] = rng.normal(0, 3, n)
    q["r"] = rng.uniform(0.02, 0.2, n)
    q["m"] = 1.0
    for col in (abi.COLLISION_LINE, abi.COLLISION_LINETREE):
        yield f"open_c{col}", abi.default_config(collision=col, root_size=10.0, boundary=abi.BOUNDARY_OPEN, dt_last_done=0.05), q
        yield f"per_c{col}", abi.default_config(collision=col, root_size=10.0, boundary=abi.BOUNDARY_PERIODIC, N_ghost_x=1,
                                                N_ghost_y=1, N_ghost_z=1, dt_last_done=0.03), q
        yield f"2root_negdt_c{col}", abi.default_config(collision=col, root_size=5.0, N_root_x=2, N_root_y=2, N_root_z=2,
                                                        boundary=abi.BOUNDARY_PERIODIC, N_ghost_x=2, N_ghost_y=1, dt_last_done=-0.02), q
        yield f"sheet_c{col}", ics.shearing_sheet_config(root_size=40.0, t=55.5, collision=col, dt_last_done=30.0), \
            ics.shearing_sheet(root_size=40.0, seed=5)


LINE_CASES = list(line_cases())


@pytest.mark.parametrize("name,cfg,p", LINE_CASES, ids=[c[0] for c in LINE_CASES])
def test_line_collision_list_bitwise(name, cfg, p):
    ref = checkers.reference().collision_search(cfg, p)
    orc = checkers.oracle().collision_search(cfg, p)
    assert len(ref) > 0
    assert checkers.collisions_equal(ref, orc, with_ri=(cfg.collision == abi.COLLISION_LINETREE))


def test_com_and_angular_momentum():
    """reb_simulation_com / reb_simulation_angular_momentum (src/tools.c:164-174, 376-408)."""
    for p, cfg in ((ics.plummer(777, seed=3), ics.plummer_config(777)),
                   (ics.planetesimal_disk(300, seed=4), ics.planetesimal_config())):
        p = p.copy()
        p["ax"] = np.linspace(-1, 1, len(p))
        assert checkers.reference().com(cfg, p) == checkers.oracle().com(cfg, p)
        assert checkers.reference().angular_momentum(cfg, p) == checkers.oracle().angular_momentum(cfg, p)


def test_gravity_cs_compensation_terms():
    """r->gravity_cs after reb_gravity_compensated_calculate_acceleration (src/gravity.c:293-414), which IAS15 reads."""
    cases = [(ics.plummer(300, seed=3), ics.plummer_config(300, gravity=abi.GRAVITY_COMPENSATED)),
             (ics.plummer(300, seed=3), ics.plummer_config(300, gravity=abi.GRAVITY_COMPENSATED, gravity_ignore_terms=1)),
             (ics.plummer(300, seed=3), ics.plummer_config(300, gravity=abi.GRAVITY_COMPENSATED, gravity_ignore_terms=2))]
    for typ in (0, 1):
        q = ics.planetesimal_disk(200, seed=4)
        q["m"][10:] = 1e-9
        cases.append((q, ics.planetesimal_config(testparticle_type=typ, gravity=abi.GRAVITY_COMPENSATED)))
    for p, cfg in cases:
        pr, cr = checkers.reference().gravity_cs(cfg, p)
        po, co = checkers.oracle().gravity_cs(cfg, p)
        assert bits_equal(po, pr)
        assert np.array_equal(co.view(np.uint64), cr.view(np.uint64))
        assert np.any(co != 0.0)


def test_quadrupole_tree_gravity():
    """-DQUADRUPOLE build of the reference (src/tree.c:148-198 moments, :293-303 force) against the oracle with
    cfg.quadrupole = 1; and the option really changes the result."""
    ref = checkers.reference(quadrupole=True)
    if ref is None:
        pytest.skip("oracle/_ref/libref_harness_quad.so not built")
    cases = [(ics.selfgravity_disc(2000, seed=3), ics.selfgravity_disc_config(collision=abi.COLLISION_NONE)),
             (ics.shearing_sheet(root_size=30.0, seed=4), ics.shearing_sheet_config(root_size=30.0, t=3.3, collision=abi.COLLISION_NONE)),
             (ics.plummer(500, seed=5), ics.plummer_config(500, gravity=abi.GRAVITY_TREE, root_size=200.0, opening_angle2=1.0))]
    for p, cfg in cases:
        cq = cfg.copy()
        cq.quadrupole = 1
        want, _ = ref.gravity(cq, p)
        got, _ = checkers.oracle().gravity(cq, p)
        assert bits_equal(got, want)
        mono, _ = checkers.oracle().gravity(cfg, p)
        assert not bits_equal(got, mono)
        # the quadrupole moves the tree force towards the direct sum
        if cfg.N_ghost_x == 0:
            cd = cfg.copy()
            cd.gravity = abi.GRAVITY_BASIC
            direct, _ = checkers.oracle().gravity(cd, p)
            assert checkers.max_rel_acc_error(got, direct) < checkers.max_rel_acc_error(mono, direct)
