"""Host-side mirror of the reference's simulation interface for the hot path.

`Engine` is a 1:1 ctypes binding of include/rebound_b200.h (the drop-in C ABI).  `Simulation`
mirrors the part of the reference's `rebound.Simulation` that drives the hot path -- the attribute
names (`gravity`, `collision`, `boundary`, `integrator`, `N_active`, `testparticle_type`,
`softening`, `opening_angle2`, `root_size`, `N_ghost_x`, ...), `add`, `steps`, `integrate`,
`synchronize` (rebound/simulation.py:1276-1325 in the reference) and its error behaviour (pending
error => RuntimeError with the reference's message text).

There is no CPU path: if the CUDA library is missing or no device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librebound_b200.so")


class LibraryMissing(ImportError):
    pass


class ReboundCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{abi.ERRORS.get(code, code)}: {msg}")
        self.code = code
        self.msg = msg


_lib = None
_fn = None


def load_library():
    """Loads rebound_b200/librebound_b200.so and binds every symbol declared in the header."""
    global _lib, _fn
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C rebound_b200/csrc).  rebound_b200 has no CPU fallback."
            )
        _lib = C.CDLL(LIB_PATH)
        sigs = dict(abi.PRODUCT_SIGNATURES)
        _fn = abi.bind(_lib, "rebcu_", sigs)
    return _fn


EXCHANGE_CB = C.CFUNCTYPE(None, C.c_void_p)
COLLISION_CB = C.CFUNCTYPE(C.c_int, C.c_void_p)


class Engine:
    """One rebcu_handle (= one simulation's device state)."""

    def __init__(self, device=0, stream=None, devices=None):
        """devices: a list of device indices -> one multi-GPU group handle (rebcu_create_group); a device may repeat."""
        self.f = load_library()
        if devices is not None and len(devices) > 1:
            arr = (C.c_int * len(devices))(*devices)
            self.h = self.f["create_group"](arr, len(devices))
        else:
            self.h = self.f["create"](device if devices is None else devices[0], stream)
        if not self.h:
            raise ReboundCudaError(-1, f"rebcu_create(device={device if devices is None else devices}) failed: no usable CUDA device")
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.f["destroy"](self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, err):
        if err != 0:
            raise ReboundCudaError(err, self.f["last_error"](self.h).decode())

    # residency
    def upload(self, p):
        assert p.dtype == abi.PARTICLE_DTYPE and p.flags.c_contiguous
        self._check(self.f["upload"](self.h, abi.as_ptr(p), len(p)))

    def download(self, out=None):
        n = self.N
        if out is None:
            out = abi.particles(n)
        self._check(self.f["download"](self.h, abi.as_ptr(out), len(out)))
        return out[:n]

    @property
    def N(self):
        return int(self.f["N"](self.h))

    def synchronize(self):
        self._check(self.f["synchronize"](self.h))

    def download_gravity_cs(self):
        """r->gravity_cs of the last COMPENSATED force evaluation, shape (N, 3)."""
        n = self.N
        out = np.zeros((max(n, 1), 3), dtype=np.float64)
        self._check(self.f["download_gravity_cs"](self.h, out.ctypes.data_as(C.POINTER(C.c_double)), n))
        return out[:n]

    def device_field(self, field):
        return self.f["device_field"](self.h, field)

    # hot path
    def update_acceleration(self, cfg):
        self._check(self.f["update_acceleration"](self.h, C.byref(cfg)))

    def integrator_step(self, cfg):
        self._check(self.f["integrator_step"](self.h, C.byref(cfg)))

    def boundary_check(self, cfg):
        self._check(self.f["boundary_check"](self.h, C.byref(cfg)))

    def apply_jerk(self, cfg, v):
        """reb_gravity_basic_calculate_and_apply_jerk (src/gravity.c:850-924) on the resident state."""
        self._check(self.f["apply_jerk"](self.h, C.byref(cfg), float(v)))

    def jerk_host(self, cfg, p, v):
        self._check(self.f["jerk_host"](self.h, C.byref(cfg), abi.as_ptr(p), len(p), float(v)))

    def steps(self, cfg, n):
        err = self.f["steps"](self.h, C.byref(cfg), n)
        if err == 1:        # REBCU_INTERRUPTED: the reference's Python raises KeyboardInterrupt for REB_STATUS_SIGINT
            raise KeyboardInterrupt
        self._check(err)

    def set_interrupt_flag(self, flag):
        """flag: a ctypes.c_int the caller raises above 1 to stop rebcu_steps (the reference's reb_sigint), or None."""
        self._interrupt_flag = flag      # keep it alive
        self._check(self.f["set_interrupt_flag"](self.h, C.byref(flag) if flag is not None else None))

    def collision_search(self, cfg, cap=None):
        cap = cap or max(1024, 4 * self.N)
        out = np.zeros(cap, dtype=abi.COLLISION_DTYPE)
        n = C.c_uint64(0)
        self._check(self.f["collision_search"](self.h, C.byref(cfg), abi.as_ptr(out), cap, C.byref(n)))
        if n.value > cap:
            out = np.zeros(n.value, dtype=abi.COLLISION_DTYPE)
            self._check(self.f["collisions_fetch"](self.h, abi.as_ptr(out), n.value, C.byref(n)))
        return out[: n.value]

    def exit_check(self, exit_max_distance=0.0, exit_min_distance=0.0):
        """(escape, encounter) flags of run_heartbeat's exit conditions (src/simulation.c:242-272) on the resident state."""
        a, b = C.c_int(0), C.c_int(0)
        self._check(self.f["exit_check"](self.h, float(exit_max_distance), float(exit_min_distance), C.byref(a), C.byref(b)))
        return bool(a.value), bool(b.value)

    def set_collision_subset(self, map=None, n_targets=None):
        """r->map / r->N_map / r->N_targets for the following searches (src/collision.c:53-58); no arguments resets."""
        m = None if map is None else np.ascontiguousarray(map, dtype=np.uint64)
        self._check(self.f["set_collision_subset"](self.h, None if m is None else abi.as_ptr(m), 0 if m is None else len(m),
                                                   abi.SIZE_MAX if n_targets is None else int(n_targets)))

    def collisions_fetch(self):
        n = C.c_uint64(0)
        self._check(self.f["collisions_fetch"](self.h, None, 0, C.byref(n)))
        out = np.zeros(max(1, n.value), dtype=abi.COLLISION_DTYPE)
        self._check(self.f["collisions_fetch"](self.h, abi.as_ptr(out), len(out), C.byref(n)))
        return out[: n.value]

    def tree(self, cfg):
        self._check(self.f["tree_build"](self.h, C.byref(cfg)))
        n = int(self.f["tree_cell_count"](self.h))
        out = np.zeros(max(n, 1), dtype=abi.TREECELL_DTYPE)
        self._check(self.f["tree_fetch"](self.h, abi.as_ptr(out), len(out)))
        return out[:n]

    # host-buffer drop-ins
    def gravity_host(self, cfg, p):
        n = C.c_uint64(len(p))
        self._check(self.f["gravity_host"](self.h, C.byref(cfg), abi.as_ptr(p), C.byref(n)))
        return n.value

    def steps_host(self, cfg, p, n_steps):
        n = C.c_uint64(len(p))
        self._check(self.f["steps_host"](self.h, C.byref(cfg), abi.as_ptr(p), C.byref(n), n_steps))
        return n.value

    def collision_search_host(self, cfg, p, cap=None):
        cap = cap or max(1024, 4 * len(p))
        out = np.zeros(cap, dtype=abi.COLLISION_DTYPE)
        n = C.c_uint64(0)
        self._check(self.f["collision_search_host"](self.h, C.byref(cfg), abi.as_ptr(p), len(p), abi.as_ptr(out), cap, C.byref(n)))
        if n.value > cap:
            return self.collision_search_host(cfg, p, cap=n.value)
        return out[: n.value]

    # device-side hard-sphere resolve (not bit-identical: see csrc/resolve.cu)
    def set_device_resolve(self, enable=True, restitution=None, minimum_collision_velocity=0.0, rand_seed=0):
        """restitution: None (elastic), a float (constant), or (a, b, c, lo, hi) for eps = clamp(a*pow(|v|*b, c), lo, hi)."""
        rest = None
        if restitution is not None:
            rest = abi.Restitution()
            if isinstance(restitution, (int, float)):
                rest.kind, rest.a, rest.lo, rest.hi = abi.RESTITUTION_CONSTANT, float(restitution), 0.0, 1.0
            else:
                rest.kind = abi.RESTITUTION_POWERLAW
                rest.a, rest.b, rest.c, rest.lo, rest.hi = (float(v) for v in restitution)
        self._check(self.f["set_device_resolve"](self.h, 1 if enable else 0, C.byref(rest) if rest is not None else None,
                                                 float(minimum_collision_velocity), int(rand_seed)))

    def collision_resolve(self, cfg):
        self._check(self.f["collision_resolve"](self.h, C.byref(cfg)))

    def collision_resolve_pairs(self, resolver, rand_seed, plog=0.0, log_n=0):
        """rebcu_collision_resolve_pairs on the list of the last search: `resolver(pair)` (a Python callable) gets one
        abi.ResolvePair at a time and fills v1, v2, plog_term, logged.  Returns (rand_seed, plog, log_n, rounds)."""
        def batch(_user, pairs, n):
            for j in range(n):
                resolver(pairs[j])
            return 0
        cb = abi.PAIR_RESOLVER(batch)
        seed, pl, ln, rounds = C.c_uint(rand_seed), C.c_double(plog), C.c_uint64(log_n), C.c_int(0)
        self._check(self.f["collision_resolve_pairs"](self.h, C.byref(seed), C.cast(cb, C.c_void_p), None, C.byref(pl), C.byref(ln), C.byref(rounds)))
        return int(seed.value), pl.value, int(ln.value), int(rounds.value)

    def collision_stats(self):
        plog, n, seed, rounds = C.c_double(0), C.c_uint64(0), C.c_uint(0), C.c_int(0)
        self._check(self.f["collision_stats"](self.h, C.byref(plog), C.byref(n), C.byref(seed), C.byref(rounds)))
        return {"collisions_plog": plog.value, "collisions_log_n": int(n.value), "rand_seed": int(seed.value), "rounds": int(rounds.value)}

    # sharding
    def set_shard(self, rank, world):
        self._check(self.f["set_shard"](self.h, rank, world))

    def shard_range(self):
        b, e = C.c_uint64(0), C.c_uint64(0)
        self.f["shard_range"](self.h, C.byref(b), C.byref(e))
        return b.value, e.value

    # native exchange (csrc/comm.cu)
    def comm_unique_id(self):
        buf = (C.c_ubyte * 128)()
        if self.f["comm_unique_id"](buf) != 0:
            raise ReboundCudaError(-1, "ncclGetUniqueId failed")
        return bytes(buf)

    def comm_init_rank(self, unique_id, rank, world):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._check(self.f["comm_init_rank"](self.h, buf, rank, world))

    def set_sharded_build(self, mode):
        """0: replicated tree builds, 1: per-rank subtree builds whenever possible, 2: automatic (N >= 2^18)."""
        self._check(self.f["set_sharded_build"](self.h, int(mode)))

    def comm_destroy(self):
        self._check(self.f["comm_destroy"](self.h))

    def comm_stats(self):
        b, n, t = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        self._check(self.f["comm_stats"](self.h, C.byref(b), C.byref(n), C.byref(t)))
        return {"bytes_received": int(b.value), "exchanges": int(n.value), "transport": {0: None, 1: "nccl", 2: "local"}[t.value]}

    def exchange(self, need=abi.EXCHANGE_ALL):
        self._check(self.f["exchange"](self.h, int(need)))

    def upload_shard(self, block, n_total):
        assert block.dtype == abi.PARTICLE_DTYPE and block.flags.c_contiguous
        self._check(self.f["upload_shard"](self.h, abi.as_ptr(block), int(n_total)))

    def download_shard(self, out):
        self._check(self.f["download_shard"](self.h, abi.as_ptr(out), len(out)))
        return out

    def set_exchange_callback(self, fn):
        cb = EXCHANGE_CB(lambda _u: fn()) if fn else None
        self._keep.append(cb)
        self._check(self.f["set_exchange_callback"](self.h, C.cast(cb, C.c_void_p) if cb else None, None))

    @property
    def exchange_request(self):
        """Bit mask of abi.EXCHANGE_*: what the exchange callback running right now has to gather."""
        return int(self.f["exchange_request"](self.h))

    def collisions_segments(self):
        """Entries per segment of the local collision list (rebcu_collisions_segments)."""
        counts = (C.c_uint64 * 32)()
        n = C.c_uint64(0)
        self._check(self.f["collisions_segments"](self.h, counts, 32, C.byref(n)))
        return [int(counts[i]) for i in range(n.value)]

    def set_collision_callback(self, fn):
        cb = COLLISION_CB(lambda _u: int(fn() or 0)) if fn else None
        self._keep.append(cb)
        self._check(self.f["set_collision_callback"](self.h, C.cast(cb, C.c_void_p) if cb else None, None))

    # diagnostics on the resident state
    def energy(self, cfg):
        """(kinetic, potential, total) -- reb_simulation_energy without energy_offset."""
        out = (C.c_double * 3)()
        self._check(self.f["energy"](self.h, C.byref(cfg), out))
        return out[0], out[1], out[2]

    def com(self):
        """Centre of mass as a dict m, x, y, z, vx, vy, vz, ax, ay, az (reb_simulation_com)."""
        out = (C.c_double * 10)()
        self._check(self.f["com"](self.h, out))
        return dict(zip(("m", "x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az"), out))

    def angular_momentum(self):
        out = (C.c_double * 3)()
        self._check(self.f["angular_momentum"](self.h, out))
        return out[0], out[1], out[2]

    def tree_walk_stats(self, cfg):
        """Work counters of the tree walk on the current tree (rebcu_tree_walk_stats)."""
        out = (C.c_uint64 * 6)()
        self._check(self.f["tree_walk_stats"](self.h, C.byref(cfg), out))
        keys = ("interactions", "visits", "group_entries", "group_visits", "groups", "cells")
        return dict(zip(keys, (int(v) for v in out)))

    def measure_fp64_peak(self):
        out = C.c_double(0)
        self._check(self.f["measure_fp64_peak"](self.h, C.byref(out)))
        return out.value

    # instrumentation
    @property
    def launch_count(self):
        return int(self.f["launch_count"](self.h))

    def selftest_math(self, n_samples, seed=1):
        out = (C.c_uint64 * 4)()
        self._check(self.f["selftest_math"](self.h, n_samples, seed, out))
        return {"sqrt_mismatch": int(out[0]), "div_mismatch": int(out[1]), "sqrt_flagged": int(out[2]), "div_flagged": int(out[3])}

    def selftest_sort(self, keys, vals, bits=64):
        """Sorts (keys uint64, vals uint32) in place on the device with the tree build's radix sort."""
        assert keys.dtype == np.uint64 and vals.dtype == np.uint32 and len(keys) == len(vals)
        self._check(self.f["selftest_sort"](self.h, abi.as_ptr(keys), abi.as_ptr(vals), len(keys), bits))

    def selftest_scan(self, values):
        assert values.dtype == np.uint32
        self._check(self.f["selftest_scan"](self.h, abi.as_ptr(values), len(values)))

    def timing_enable(self, on=True):
        self._check(self.f["timing_enable"](self.h, 1 if on else 0))

    def timing_reset(self):
        self._check(self.f["timing_reset"](self.h))

    def timing_read(self):
        ms = (C.c_double * 8)()
        n = (C.c_uint64 * 8)()
        self._check(self.f["timing_read"](self.h, ms, n, 8))
        names = ("direct", "kickdrift", "treebuild", "treewalk", "collision", "boundary", "pack", "exchange")
        return {k: {"ms": ms[i], "launches": int(n[i])} for i, k in enumerate(names)}


_GRAVITY = {"none": 0, "basic": 1, "compensated": 2, "tree": 3}
_COLLISION = {"none": 0, "direct": 1, "tree": 2, "line": 4, "linetree": 5}
_BOUNDARY = {"none": 0, "open": 1, "periodic": 2, "shear": 3}
_INTEGRATOR = {"none": 0, "leapfrog": 1, "sei": 2}
_ENUMS = {"gravity": _GRAVITY, "collision": _COLLISION, "boundary": _BOUNDARY, "integrator": _INTEGRATOR}
_CFG_FIELDS = {name for name, _ in abi.Config._fields_}


class Escape(Exception):
    """A particle is farther than exit_max_distance from the origin (the reference's rebound.Escape, REB_STATUS_ESCAPE)."""


class Encounter(Exception):
    """Two particles are closer than exit_min_distance (the reference's rebound.Encounter, REB_STATUS_ENCOUNTER)."""


class Simulation:
    """The hot-path subset of the reference's `rebound.Simulation`, resident on one B200.

    Enum-valued attributes accept the reference's string names (rebound/simulation.py:18-20).
    Particles added with `add` live in a host array until the next step, then stay in HBM until
    `synchronize()` / `particles` is read (the is_synchronized protocol, src/simulation.c:633-637).
    """

    def __init__(self, device=0, stream=None):
        object.__setattr__(self, "_cfg", abi.default_config())
        object.__setattr__(self, "_engine", Engine(device, stream))
        object.__setattr__(self, "_host", abi.particles(0))
        object.__setattr__(self, "_host_valid", True)   # host copy is current
        object.__setattr__(self, "_dev_valid", False)   # device copy is current
        object.__setattr__(self, "steps_done", 0)
        object.__setattr__(self, "collision_resolve", None)
        object.__setattr__(self, "exit_max_distance", 0.0)      # src/rebound.h: 0 = check off
        object.__setattr__(self, "exit_min_distance", 0.0)

    def __setattr__(self, name, value):
        if name in _ENUMS:
            if isinstance(value, str):
                if value not in _ENUMS[name]:
                    raise ValueError(f"{name} '{value}' not found")
                value = _ENUMS[name][value]
            setattr(self._cfg, name, value)
        elif name == "N_active":
            self._cfg.N_active = abi.SIZE_MAX if value in (-1, None) else value
        elif name in _CFG_FIELDS:
            setattr(self._cfg, name, value)
        else:
            object.__setattr__(self, name, value)

    def __getattr__(self, name):
        if name in _CFG_FIELDS:
            return getattr(self._cfg, name)
        raise AttributeError(name)

    @property
    def config(self):
        return self._cfg

    @property
    def engine(self):
        return self._engine

    @property
    def N(self):
        return self._engine.N if self._dev_valid else len(self._host)

    def add(self, particles=None, **kw):
        """reb_simulation_add (src/particle.c:47-76): appends particles; marks the device copy stale."""
        self.synchronize()
        if particles is None:
            particles = abi.particles(1)
            for k, v in kw.items():
                particles[k] = v
        object.__setattr__(self, "_host", np.concatenate([self._host, particles.astype(abi.PARTICLE_DTYPE)]))
        object.__setattr__(self, "_dev_valid", False)

    @property
    def particles(self):
        self.synchronize()
        return self._host

    def did_modify_particles(self):
        """The caller edited `particles` in place (r->did_modify_particles, src/rebound.h:247)."""
        object.__setattr__(self, "_dev_valid", False)
        object.__setattr__(self, "_host_valid", True)

    def synchronize(self):
        if not self._host_valid:
            object.__setattr__(self, "_host", self._engine.download())
            object.__setattr__(self, "_host_valid", True)

    def _to_device(self):
        if not self._dev_valid:
            self._engine.upload(np.ascontiguousarray(self._host))
            object.__setattr__(self, "_dev_valid", True)

    def update_acceleration(self):
        self._to_device()
        self._engine.update_acceleration(self._cfg)
        object.__setattr__(self, "_host_valid", False)

    def steps(self, n):
        """reb_simulation_steps (src/simulation.c:504-513)."""
        self._to_device()
        self._engine.steps(self._cfg, int(n))
        object.__setattr__(self, "_host_valid", False)
        object.__setattr__(self, "steps_done", self.steps_done + int(n))

    def _exit_check(self):
        """run_heartbeat's exit conditions (src/simulation.c:242-272), tested on the device."""
        self._to_device()
        escape, encounter = self._engine.exit_check(self.exit_max_distance, self.exit_min_distance)
        if encounter:       # tested after the escape condition, so its status is the one that survives
            raise Encounter("Two particles had a close encounter (d<exit_min_distance).")
        if escape:
            raise Escape("A particle escaped (r>exit_max_distance).")

    def step(self):
        self.steps(1)

    def integrate(self, tmax):
        """reb_simulation_integrate with exact_finish_time=0 semantics (src/simulation.c:465): whole
        steps of size dt until t >= tmax."""
        c = self._cfg
        if c.dt == 0:
            raise RuntimeError("dt is zero")
        if tmax == c.t:
            return
        # simulation.c:377-380: dt takes the sign of the direction of integration
        c.dt = float(np.copysign(c.dt, tmax - c.t))
        # count the steps with the integrator's own update of t: two half drifts per step (integrator_leapfrog.c:81,
        # integrator_sei.c:107,116), which round differently from t += dt
        half = c.dt / 2.0
        n = 0
        t = c.t
        while (t < tmax) if c.dt > 0 else (t > tmax):
            t += half
            t += half
            n += 1
        if not (self.exit_max_distance or self.exit_min_distance):
            if n:
                self.steps(n)
            return
        # the reference runs the heartbeat (and with it the exit checks) before the first step and after every
        # step (src/simulation.c:392, 431-432); the step that trips a condition is completed
        self._exit_check()
        for _ in range(n):
            self.steps(1)
            self._exit_check()

    def collision_search(self):
        self._to_device()
        return self._engine.collision_search(self._cfg)

    def set_device_resolve(self, enable=True, restitution=None, minimum_collision_velocity=0.0, rand_seed=0):
        """Hard-sphere resolve on the device after every step's collision search (Engine.set_device_resolve); the
        counterpart of `sim.collision_resolve = "hardsphere"` for a resident simulation.  Not bit-identical."""
        self._engine.set_device_resolve(enable, restitution, minimum_collision_velocity, rand_seed)

    @property
    def collision_stats(self):
        """collisions_plog / collisions_log_n as advanced by the device resolver."""
        return self._engine.collision_stats()

    def energy(self):
        """reb_simulation_energy (src/tools.c:108-162), evaluated on the device."""
        self._to_device()
        return self._engine.energy(self._cfg)[2]

    def com(self):
        """reb_simulation_com (src/tools.c:401-408), evaluated on the device."""
        self._to_device()
        return self._engine.com()

    def angular_momentum(self):
        """reb_simulation_angular_momentum (src/tools.c:164-174), evaluated on the device."""
        self._to_device()
        return self._engine.angular_momentum()

    def tree(self):
        self._to_device()
        return self._engine.tree(self._cfg)

    def close(self):
        self._engine.close()
