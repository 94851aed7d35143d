"""Synthetic initial conditions for the five BASELINE.json configurations.

numpy restatements of the reference's recipes (cited per function).  They do not have to be
bit-identical to the reference's generators: parity is judged on the hot path, which every
implementation (CUDA, oracle, reference harness) receives as the same particle array.
"""
import math

import numpy as np

from . import abi


def _fill(p, x, v, m, r=None):
    p["x"], p["y"], p["z"] = x[:, 0], x[:, 1], x[:, 2]
    p["vx"], p["vy"], p["vz"] = v[:, 0], v[:, 1], v[:, 2]
    p["m"] = m
    if r is not None:
        p["r"] = r
    return p


def plummer(n, M=1.0, R=1.0, seed=42):
    """Plummer sphere after Aarseth, Henon & Wielen 1974 (reb_simulation_add_plummer,
    src/tools.c:463-502) moved to its centre of mass (examples/selfgravity_plummer/problem.c:44)."""
    rng = np.random.default_rng(seed)
    E = 3.0 / 64.0 * math.pi * M * M / R
    rr = (rng.random(n) ** (-2.0 / 3.0) - 1.0) ** -0.5
    x2 = rng.random(n)
    x3 = rng.random(n) * 2 * math.pi
    z = (1 - 2 * x2) * rr
    rho = np.sqrt(np.maximum(rr * rr - z * z, 0.0))
    pos = np.stack([rho * np.cos(x3), rho * np.sin(x3), z], axis=1)
    q = np.empty(n)
    todo = np.arange(n)
    while todo.size:  # von Neumann rejection, tools.c:477-481
        x5 = rng.random(todo.size)
        qq = rng.random(todo.size)
        g = qq * qq * (1 - qq * qq) ** 3.5
        ok = 0.1 * x5 <= g
        q[todo[ok]] = qq[ok]
        todo = todo[~ok]
    v = q * math.sqrt(2.0) * (1 + rr * rr) ** -0.25
    x6 = rng.random(n)
    x7 = rng.random(n) * 2 * math.pi
    vz = (1 - 2 * x6) * v
    vr = np.sqrt(np.maximum(v * v - vz * vz, 0.0))
    vel = np.stack([vr * np.cos(x7), vr * np.sin(x7), vz], axis=1)
    pos *= 3 * math.pi / 64.0 * M * M / E
    vel *= math.sqrt(E * 64.0 / 3.0 / math.pi / M)
    m = np.full(n, M / n)
    pos -= pos.mean(axis=0)
    vel -= vel.mean(axis=0)
    return _fill(abi.particles(n), pos, vel, m)


def plummer_config(n, **kw):
    """examples/selfgravity_plummer/problem.c:26-45 (G=M=R=1)."""
    M = R = 1.0
    E = 3.0 / 64.0 * math.pi * M * M / R
    r0 = 16.0 / (3.0 * math.pi) * R
    t0 = 1.0 * M**2.5 * (4.0 * E) ** -1.5 * n / math.log(0.4 * n)
    c = abi.default_config(G=1.0, dt=2e-5 * t0, softening=0.01 * r0, gravity=abi.GRAVITY_BASIC,
                           integrator=abi.INTEGRATOR_LEAPFROG)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def planetesimal_disk(n_test, seed=42):
    """Config C2: a star, nine planets on circular orbits (the actives) and n_test massless
    planetesimals with a in U(0.4,20), e in U(0.01,0.2), random omega and f
    (cf. examples/solar_system_with_testparticles/problem.c:36-43)."""
    rng = np.random.default_rng(seed)
    n_act = 10
    n = n_act + n_test
    p = abi.particles(n)
    p["m"][0] = 1.0
    a_pl = np.arange(1, 10, dtype=np.float64)
    m_pl = np.linspace(1e-4, 1e-3, 9)
    ph = rng.random(9) * 2 * math.pi
    vk = np.sqrt(1.0 / a_pl)
    p["x"][1:10] = a_pl * np.cos(ph)
    p["y"][1:10] = a_pl * np.sin(ph)
    p["vx"][1:10] = -vk * np.sin(ph)
    p["vy"][1:10] = vk * np.cos(ph)
    p["m"][1:10] = m_pl
    a = rng.uniform(0.4, 20.0, n_test)
    e = rng.uniform(0.01, 0.2, n_test)
    om = rng.uniform(0, 2 * math.pi, n_test)
    f = rng.uniform(0, 2 * math.pi, n_test)
    rr = a * (1 - e * e) / (1 + e * np.cos(f))
    v0 = np.sqrt(1.0 / (a * (1 - e * e)))
    p["x"][n_act:] = rr * np.cos(om + f)
    p["y"][n_act:] = rr * np.sin(om + f)
    p["z"][n_act:] = rr * rng.normal(0, 0.01, n_test)
    p["vx"][n_act:] = -v0 * (np.sin(om + f) + e * np.sin(om))
    p["vy"][n_act:] = v0 * (np.cos(om + f) + e * np.cos(om))
    return p


def planetesimal_config(**kw):
    c = abi.default_config(G=1.0, dt=1e-2, softening=0.0, gravity=abi.GRAVITY_BASIC,
                           integrator=abi.INTEGRATOR_LEAPFROG, N_active=10, testparticle_type=0)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def _powerlaw(rng, lo, hi, slope, n):
    """reb_random_powerlaw, src/tools.c:77-81."""
    y = rng.random(n)
    if slope == -1:
        return np.exp(y * math.log(hi / lo) + math.log(lo))
    return (((hi ** (slope + 1)) - (lo ** (slope + 1))) * y + lo ** (slope + 1)) ** (1.0 / (slope + 1))


def selfgravity_disc(n, boxsize=10.2, disc_mass=0.2, seed=42):
    """Config C4: examples/selfgravity_disc/problem.c:25-56 (star + n disc particles)."""
    rng = np.random.default_rng(seed)
    p = abi.particles(n + 1)
    p["m"][0] = 1.0
    lo, hi = boxsize / 10.0, boxsize / 2.0 / 1.2
    a = _powerlaw(rng, lo, hi, -1.5, n)
    phi = rng.uniform(0, 2 * math.pi, n)
    mu = 1.0 + disc_mass * (a**-1.5 - lo**-1.5) / (hi**-1.5 - lo**-1.5)
    vkep = np.sqrt(mu / a)
    p["x"][1:] = a * np.cos(phi)
    p["y"][1:] = a * np.sin(phi)
    p["z"][1:] = a * rng.normal(0, math.sqrt(0.001), n)
    p["vx"][1:] = vkep * np.sin(phi)
    p["vy"][1:] = -vkep * np.cos(phi)
    p["m"][1:] = disc_mass / max(n, 1)
    return p


def selfgravity_disc_config(boxsize=10.2, **kw):
    c = abi.default_config(G=1.0, dt=3e-2, softening=0.02, gravity=abi.GRAVITY_TREE,
                           boundary=abi.BOUNDARY_OPEN, opening_angle2=0.25, root_size=boxsize,
                           integrator=abi.INTEGRATOR_LEAPFROG)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


SHEET_OMEGA = 0.00013143527
SHEET_G = 6.67428e-11


def shearing_sheet(root_size=100.0, n_root=2, surfacedensity=400.0, particle_density=400.0,
                   rmin=1.0, rmax=4.0, slope=-3.0, seed=42, n_max=None):
    """Config C5: examples/shearing_sheet/problem.c:26-91 (Saturn-ring patch, 2x2 root boxes)."""
    rng = np.random.default_rng(seed)
    bx = by = root_size * n_root
    total_mass = surfacedensity * bx * by
    mean_mass = particle_density * 4.0 / 3.0 * math.pi * np.mean(_powerlaw(rng, rmin, rmax, slope, 20000) ** 3)
    n_guess = int(total_mass / mean_mass * 1.1) + 16
    rad = _powerlaw(rng, rmin, rmax, slope, n_guess)
    mass = particle_density * 4.0 / 3.0 * math.pi * rad**3
    n = int(np.searchsorted(np.cumsum(mass), total_mass)) + 1
    if n_max is not None:
        n = min(n, n_max)
    rad, mass = rad[:n], mass[:n]
    p = abi.particles(n)
    p["x"] = rng.uniform(-bx / 2, bx / 2, n)
    p["y"] = rng.uniform(-by / 2, by / 2, n)
    p["z"] = rng.normal(0, 1.0, n)
    p["vy"] = -1.5 * p["x"] * SHEET_OMEGA
    p["r"] = rad
    p["m"] = mass
    return p


def shearing_sheet_config(root_size=100.0, n_root=2, n_ghost=2, **kw):
    c = abi.default_config(G=SHEET_G, OMEGA=SHEET_OMEGA, softening=0.1,
                           dt=1e-3 * 2 * math.pi / SHEET_OMEGA, opening_angle2=0.5,
                           gravity=abi.GRAVITY_TREE, collision=abi.COLLISION_TREE,
                           boundary=abi.BOUNDARY_SHEAR, integrator=abi.INTEGRATOR_SEI,
                           root_size=root_size, N_root_x=n_root, N_root_y=n_root, N_root_z=1,
                           N_ghost_x=n_ghost, N_ghost_y=n_ghost, N_ghost_z=0)
    for k, v in kw.items():
        setattr(c, k, v)
    return c
