"""GPU parity: direct-summation gravity and leapfrog through the C ABI against the CPU oracle
(oracle/oracle.c, pinned bitwise to the reference in test_oracle_vs_reference.py).

STRICT mode must be bit-identical; FAST mode within 1e-12 relative (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

import checkers
from checkers import bits_equal, max_rel_acc_error
from rebound_b200 import abi, ics
from rebound_b200.simulation import Engine, ReboundCudaError

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ACC = ("ax", "ay", "az")


@pytest.fixture(scope="module")
def eng():
    e = Engine(0)
    yield e
    e.close()


def gpu_gravity(eng, cfg, p):
    q = p.copy()
    c = cfg.copy()
    n = eng.gravity_host(c, q)
    return q[:n], c


def direct_cases():
    p = ics.plummer(1000, seed=1)
    yield "plummer_basic", ics.plummer_config(1000), p
    yield "plummer_comp", ics.plummer_config(1000, gravity=abi.GRAVITY_COMPENSATED), p
    yield "nosoft", ics.plummer_config(1000, softening=0.0), p
    yield "ragged_37", ics.plummer_config(37), ics.plummer(37, seed=2)
    yield "single", ics.plummer_config(1), ics.plummer(1, seed=2)
    yield "two", ics.plummer_config(2), ics.plummer(2, seed=2)
    yield "n5000", ics.plummer_config(5000), ics.plummer(5000, seed=3)
    yield "n20000_comp", ics.plummer_config(20000, gravity=abi.GRAVITY_COMPENSATED), ics.plummer(20000, seed=3)
    for typ in (0, 1):
        for grav in (abi.GRAVITY_BASIC, abi.GRAVITY_COMPENSATED):
            q = ics.planetesimal_disk(3000, seed=3)
            q["m"][10:] = 1e-9
            yield f"testp{typ}_g{grav}", ics.planetesimal_config(testparticle_type=typ, gravity=grav), q
    for terms in (1, 2):
        for grav in (abi.GRAVITY_BASIC, abi.GRAVITY_COMPENSATED):
            yield f"ignore{terms}_g{grav}", ics.plummer_config(1000, gravity_ignore_terms=terms, gravity=grav), p
    # N_active in the range of the producer/adder split kernel (256 <= N_active, N < 40960): ragged source
    # ranges per particle (type 1: actives see everybody, test particles see the actives only)
    pa = ics.plummer(2500, seed=6)
    for typ in (0, 1):
        for grav in (abi.GRAVITY_BASIC, abi.GRAVITY_COMPENSATED):
            yield f"nactive700_t{typ}_g{grav}", ics.plummer_config(2500, N_active=700, testparticle_type=typ, gravity=grav), pa
    yield "nactive700_ignore2", ics.plummer_config(2500, N_active=700, testparticle_type=1, gravity_ignore_terms=2), pa
    yield "n257", ics.plummer_config(257), ics.plummer(257, seed=7)
    yield "n40959", ics.plummer_config(40959), ics.plummer(40959, seed=8)
    yield "n40961", ics.plummer_config(40961), ics.plummer(40961, seed=8)
    q = ics.planetesimal_disk(50, seed=4)
    q["m"][1:] = 1e-6
    yield "nactive1_ignore1", ics.planetesimal_config(N_active=1, testparticle_type=1, gravity_ignore_terms=1), q
    for b in (abi.BOUNDARY_PERIODIC, abi.BOUNDARY_OPEN, abi.BOUNDARY_SHEAR):
        yield f"ghost_b{b}", ics.plummer_config(300, boundary=b, root_size=30.0, N_ghost_x=1, N_ghost_y=2,
                                                N_ghost_z=1, OMEGA=1.0, t=0.37), ics.plummer(300, seed=5)


CASES = list(direct_cases())


@pytest.mark.parametrize("name,cfg,p", CASES, ids=[c[0] for c in CASES])
def test_direct_strict_bitwise(eng, name, cfg, p):
    want, _ = checkers.oracle().gravity(cfg, p)
    got, _ = gpu_gravity(eng, cfg, p)
    assert bits_equal(got, want)                      # accelerations AND untouched fields


@pytest.mark.parametrize("name,cfg,p", CASES, ids=[c[0] for c in CASES])
def test_direct_fast_within_1e12(eng, name, cfg, p):
    want, _ = checkers.oracle().gravity(cfg, p)
    c = cfg.copy()
    c.mode = abi.MODE_FAST
    got, _ = gpu_gravity(eng, c, p)
    assert max_rel_acc_error(got, want) <= 1e-12


def test_golden_plummer_reference_vector(eng):
    """Accelerations the unmodified reference produced (tests/golden/make_golden.py)."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "direct_plummer512.npz")
    g = np.load(path)
    p = g["particles_in"].view(abi.PARTICLE_DTYPE).reshape(-1)
    for key, grav in (("basic", abi.GRAVITY_BASIC), ("compensated", abi.GRAVITY_COMPENSATED)):
        cfg = ics.plummer_config(512, gravity=grav)
        cfg.softening = float(g["softening"])
        got, _ = gpu_gravity(eng, cfg, p)
        want = g["acc_" + key]
        got_acc = np.stack([got[f] for f in ACC], axis=1)
        assert np.array_equal(got_acc.view(np.uint64), want.view(np.uint64))


@pytest.mark.parametrize("order", [2, 4, 6, 8])
def test_leapfrog_steps_bitwise(eng, order):
    p = ics.plummer(700, seed=7)
    cfg = ics.plummer_config(700, leapfrog_order=order, dt=1e-3)
    want, cw, _ = checkers.oracle().steps(cfg, p, 4)
    q = p.copy()
    c = cfg.copy()
    eng.steps_host(c, q, 4)
    assert bits_equal(q, want)
    assert c.t == cw.t and c.dt_last_done == cw.dt_last_done


def test_leapfrog_step_by_step_equals_batched(eng):
    p = ics.plummer(300, seed=8)
    cfg = ics.plummer_config(300, dt=1e-3)
    q1, c1 = p.copy(), cfg.copy()
    eng.steps_host(c1, q1, 6)
    q2, c2 = p.copy(), cfg.copy()
    for _ in range(6):
        eng.steps_host(c2, q2, 1)
    assert bits_equal(q1, q2) and c1.t == c2.t


def test_leapfrog_bad_order(eng):
    p = ics.plummer(8, seed=7)
    cfg = ics.plummer_config(8, leapfrog_order=3)
    with pytest.raises(ReboundCudaError) as e:
        eng.steps_host(cfg.copy(), p.copy(), 1)
    assert e.value.msg == "Leapfrog order not supported."


@pytest.mark.parametrize("grav", [abi.GRAVITY_BASIC, abi.GRAVITY_COMPENSATED])
@pytest.mark.parametrize("mode", [abi.MODE_STRICT, abi.MODE_FAST])
def test_planetesimal_fused_step(eng, grav, mode):
    """Config C2 shape: 10 massive bodies + test particles, testparticle_type 0 (fused step kernel)."""
    p = ics.planetesimal_disk(5000, seed=9)
    cfg = ics.planetesimal_config(gravity=grav, mode=mode)
    want, cw, _ = checkers.oracle().steps(cfg, p, 7)
    q, c = p.copy(), cfg.copy()
    eng.steps_host(c, q, 7)
    assert c.t == cw.t
    if mode == abi.MODE_STRICT:
        assert bits_equal(q, want)
    else:
        for f in ("x", "y", "z", "vx", "vy", "vz"):
            np.testing.assert_allclose(q[f], want[f], rtol=1e-11, atol=1e-13)
        assert max_rel_acc_error(q, want) <= 1e-12 * 50


def test_energy_error_matches_reference_path(eng):
    """Energy error after n steps equals the oracle's (bitwise trajectories => identical energies)."""
    p = ics.plummer(512, seed=11)
    cfg = ics.plummer_config(512)
    e0 = checkers.oracle().energy(cfg, p)
    want, _, _ = checkers.oracle().steps(cfg, p, 20)
    q, c = p.copy(), cfg.copy()
    eng.steps_host(c, q, 20)
    e_gpu = checkers.oracle().energy(cfg, q)
    e_ref = checkers.oracle().energy(cfg, want)
    assert e_gpu == e_ref
    assert abs((e_gpu - e0) / e0) < 1e-3


def test_linearity_in_G_and_translation_full_size(eng):
    """Size-independent properties at config C1's N=16384: a(2G) = 2 a(G) exactly in strict mode,
    sum_i m_i a_i ~ 0 (Newton's third law)."""
    n = 16384
    p = ics.plummer(n, seed=42)
    cfg = ics.plummer_config(n)
    a1, _ = gpu_gravity(eng, cfg, p)
    c2 = cfg.copy(); c2.G = 2.0
    a2, _ = gpu_gravity(eng, c2, p)
    for f in ACC:
        assert np.array_equal(a2[f], 2.0 * a1[f])
        assert abs(np.sum(p["m"] * a1[f])) < 1e-12 * np.sum(np.abs(p["m"] * a1[f]))
    # sampled rows against the oracle arithmetic (numpy restatement of gravity.c:222-230)
    rng = np.random.default_rng(0)
    for i in rng.integers(0, n, 16):
        dx, dy, dz = p["x"][i] - p["x"], p["y"][i] - p["y"], p["z"][i] - p["z"]
        r = np.sqrt(dx * dx + dy * dy + dz * dz + cfg.softening**2)
        pre = -cfg.G / (r * r * r) * p["m"]
        pre[i] = 0.0
        acc = 0.0
        for v in pre * dx:
            acc += v
        assert acc == a1["ax"][i]


def test_strict_math_selftest(eng):
    """The branch-free sqrt/divide used by the STRICT kernels equal __dsqrt_rn/__ddiv_rn bit for bit on
    3e8 operand pairs (ordinary magnitudes, arbitrary bit patterns, near-square rounding ties)."""
    n = 300_000_000
    r = eng.selftest_math(n, seed=12345)
    assert r["sqrt_mismatch"] == 0 and r["div_mismatch"] == 0
    assert r["sqrt_flagged"] < 1e-3 * n and r["div_flagged"] < 1e-3 * n


@pytest.mark.parametrize("n_test,steps", [((1 << 18) + 77, 5), ((1 << 19), 3)])
@pytest.mark.parametrize("grav", [abi.GRAVITY_BASIC, abi.GRAVITY_COMPENSATED])
def test_planetesimal_host_pipeline_bitwise(eng, n_test, steps, grav):
    """rebcu_steps_host on a large test-particle problem takes the chunk-pipelined path (copies overlapped
    with kernels, massive-body history replayed per chunk); it must equal the oracle bit for bit, leave the
    device state resident, and agree with the plain resident path."""
    p = ics.planetesimal_disk(n_test, seed=21)
    cfg = ics.planetesimal_config(gravity=grav)
    want, cw, _ = checkers.oracle().steps(cfg, p, steps)
    q, c = p.copy(), cfg.copy()
    eng.steps_host(c, q, steps)
    assert c.t == cw.t and c.dt_last_done == cw.dt_last_done
    assert bits_equal(q, want)
    assert bits_equal(eng.download(), want)          # device copy is the final state too
    c2 = cfg.copy()
    eng.upload(p.copy())
    eng.steps(c2, steps)
    assert bits_equal(eng.download(), want)


@pytest.mark.parametrize("mode", [abi.MODE_STRICT, abi.MODE_FAST])
def test_planetesimal_single_steps_equal_multistep_launch(eng, mode):
    """One step per call (tp_leapfrog_kernel, one launch per step) and n steps per call (massive-body history +
    tp_multistep_kernel, two launches in total) execute the same operation sequence per particle."""
    p = ics.planetesimal_disk(3000, seed=31)
    cfg = ics.planetesimal_config(mode=mode)
    eng.upload(p.copy())
    c1 = cfg.copy()
    for _ in range(6):
        eng.steps(c1, 1)
    a = eng.download().copy()
    eng.upload(p.copy())
    c2 = cfg.copy()
    eng.steps(c2, 6)
    b = eng.download().copy()
    assert c1.t == c2.t
    assert bits_equal(a, b)
    if mode == abi.MODE_STRICT:
        want, _, _ = checkers.oracle().steps(cfg, p, 6)
        assert bits_equal(a, want)


@pytest.mark.parametrize("n", [100, 1000, 20000, 30000])
def test_compensation_terms_bitwise(eng, n):
    """r->gravity_cs (the running Kahan compensation, src/gravity.c:297-341; read by IAS15) through every strict
    compensated kernel: one thread per particle (n < 256, n >= 20480) and the producer/adder split kernel."""
    p = ics.plummer(n, seed=12)
    cfg = ics.plummer_config(n, gravity=abi.GRAVITY_COMPENSATED)
    want_p, want_cs = checkers.oracle().gravity_cs(cfg, p)
    eng.upload(np.ascontiguousarray(p))
    eng.update_acceleration(cfg.copy())
    got_cs = eng.download_gravity_cs()
    assert bits_equal(eng.download(), want_p)
    assert np.array_equal(got_cs.view(np.uint64), want_cs.view(np.uint64))
    # FAST mode: the corrected sum a - cs agrees with the strict one to the tolerance of the mode
    c = cfg.copy()
    c.mode = abi.MODE_FAST
    eng.update_acceleration(c)
    cs_fast = eng.download_gravity_cs()
    assert np.all(np.isfinite(cs_fast)) and np.max(np.abs(cs_fast)) <= 1e-12 * np.max(np.abs(want_p["ax"]))


def test_compensation_terms_need_a_compensated_evaluation(eng):
    p = ics.plummer(64, seed=1)
    eng.upload(np.ascontiguousarray(p))
    eng.update_acceleration(ics.plummer_config(64))
    with pytest.raises(Exception):
        eng.download_gravity_cs()


@pytest.mark.parametrize("grav", [abi.GRAVITY_BASIC, abi.GRAVITY_COMPENSATED])
@pytest.mark.parametrize("n", [40, 600])
def test_coincident_particles_give_nan_where_the_reference_does(eng, grav, n):
    """Two particles at the same coordinates with zero softening: -G/0 * 0 = NaN in the reference's loop
    (gravity.c:222-230).  The GPU produces NaN for exactly the same particles and components (NaN sign/payload is
    the one thing not compared: x86 SSE returns the negative 'indefinite' NaN, the GPU the positive canonical one);
    everything finite stays bit-identical."""
    p = ics.plummer(n, seed=21)
    p["x"][7], p["y"][7], p["z"][7] = p["x"][3], p["y"][3], p["z"][3]
    cfg = ics.plummer_config(n, gravity=grav, softening=0.0)
    want, _ = checkers.oracle().gravity(cfg, p)
    got, _ = gpu_gravity(eng, cfg, p)
    for f in ("ax", "ay", "az"):
        assert np.array_equal(np.isnan(got[f]), np.isnan(want[f])), f
        ok = ~np.isnan(want[f])
        assert np.array_equal(got[f][ok].view(np.uint64), want[f][ok].view(np.uint64)), f
    assert np.isnan(want["ax"][3]) and np.isnan(want["ax"][7]) and not np.isnan(want["ax"][0])


def _type1_case(n, n_active, seed):
    q = ics.planetesimal_disk(n, seed=seed)
    q["m"][10:] = 1e-9
    if n_active > 10:                      # promote some planetesimals to (light) massive bodies
        q["m"][10:n_active] = 3e-7
    return q


@pytest.mark.parametrize("grav", [abi.GRAVITY_BASIC, abi.GRAVITY_COMPENSATED])
@pytest.mark.parametrize("n,n_active,terms", [(6000, 10, 0), (6000, 10, 1), (6000, 10, 2), (5000, 200, 0), (9000, 256, 0),
                                              (7000, 300, 0), (7000, 1000, 2), (6000, 600, 1)])
def test_type1_massive_rows_bitwise(eng, grav, n, n_active, terms):
    """testparticle_type 1 with few massive particles among many (N >= 4096): up to 256 massive rows go through the
    term-buffer + ordered-sum kernels, more of them through the producer/adder split kernel; accelerations (and the Kahan compensation) stay bit-identical,
    FAST stays within 1e-12."""
    p = _type1_case(n, n_active, seed=31)
    cfg = ics.planetesimal_config(testparticle_type=1, gravity=grav, N_active=n_active, gravity_ignore_terms=terms)
    want, _ = checkers.oracle().gravity(cfg, p)
    got, _ = gpu_gravity(eng, cfg, p)
    assert bits_equal(got, want)
    if grav == abi.GRAVITY_COMPENSATED:
        _, want_cs = checkers.oracle().gravity_cs(cfg, p)
        eng.upload(np.ascontiguousarray(p))
        eng.update_acceleration(cfg.copy())
        assert np.array_equal(eng.download_gravity_cs().view(np.uint64), want_cs.view(np.uint64))
    c = cfg.copy()
    c.mode = abi.MODE_FAST
    fast, _ = gpu_gravity(eng, c, p)
    assert max_rel_acc_error(fast, want) <= 1e-12


def test_type1_massive_rows_in_several_source_chunks(tmp_path):
    """The term buffer is filled and consumed chunk by chunk (REBOUND_B200_ROWCHUNK forces small chunks here); the
    running sums carry over between chunks."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import checkers
        from rebound_b200 import abi, ics
        from rebound_b200.simulation import Engine
        import test_gpu_direct as T
        eng = Engine(0)
        for grav in (abi.GRAVITY_BASIC, abi.GRAVITY_COMPENSATED):
            p = T._type1_case(7000, 10, seed=33)
            cfg = ics.planetesimal_config(testparticle_type=1, gravity=grav)
            want, _ = checkers.oracle().gravity(cfg, p)
            got, _ = T.gpu_gravity(eng, cfg, p)
            assert checkers.bits_equal(got, want)
        print("CHUNKS OK")
    """ % (ROOT, os.path.join(ROOT, "tests")))
    env = dict(os.environ, REBOUND_B200_ROWCHUNK="2048")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "CHUNKS OK" in r.stdout, r.stderr


def test_apply_jerk_bitwise(eng):
    """reb_gravity_basic_calculate_and_apply_jerk (gravity.c:850-924): the serial scatter order reproduced by one
    ascending sum per particle; same cases as the oracle-vs-reference pin, host-buffer and resident entry points."""
    from test_oracle_vs_reference import _jerk_cases
    for name, cfg, p, v in _jerk_cases():
        want = checkers.oracle().apply_jerk(cfg, p, v)
        q = np.ascontiguousarray(p.copy())
        eng.jerk_host(cfg.copy(), q, v)
        assert bits_equal(q, want), name
        eng.upload(np.ascontiguousarray(p))
        eng.apply_jerk(cfg.copy(), v)
        assert bits_equal(eng.download(), want), name + " (resident)"
    # a size with many CTAs and several tiles per row
    p = ics.plummer(3001, seed=21)
    cfg = ics.plummer_config(3001)
    p, _ = checkers.oracle().gravity(cfg, p)
    want = checkers.oracle().apply_jerk(cfg, p, 0.01)
    q = np.ascontiguousarray(p.copy())
    eng.jerk_host(cfg.copy(), q, 0.01)
    assert bits_equal(q, want)
