"""ctypes / numpy mirror of include/rebound_b200.h.

The reference's Python layer is itself a ctypes mirror of its C structs
(rebound/simulation.py:1365-1469, rebound/particle.py:1031-1045); this module plays the same role
for the structs of the hot-path C ABI.  It contains no arithmetic.
"""
import ctypes as C

import numpy as np

SIZE_MAX = 2**64 - 1

# struct reb_particle, src/rebound.h:86-104 (112 bytes)
PARTICLE_DTYPE = np.dtype(
    [(n, "<f8") for n in ("x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az", "m", "r")]
    + [("name", "<u8"), ("ap", "<u8"), ("sim", "<u8")]
)
assert PARTICLE_DTYPE.itemsize == 112

# struct reb_collision, src/rebound.h:144-149 (72 bytes)
COLLISION_DTYPE = np.dtype(
    [("p1", "<u8"), ("p2", "<u8")]
    + [("gb_" + n, "<f8") for n in ("x", "y", "z", "vx", "vy", "vz")]
    + [("ri", "<u8")]
)
assert COLLISION_DTYPE.itemsize == 72

# rebcu_treecell
TREECELL_DTYPE = np.dtype(
    [(n, "<f8") for n in ("x", "y", "z", "w", "m", "mx", "my", "mz")]
    + [("pt", "<i4"), ("skip", "<i4"), ("depth", "<i4"), ("rootbox", "<i4")]
)
assert TREECELL_DTYPE.itemsize == 80

COLLISION_NONE, COLLISION_DIRECT, COLLISION_TREE = 0, 1, 2
COLLISION_LINE, COLLISION_LINETREE = 4, 5
BOUNDARY_NONE, BOUNDARY_OPEN, BOUNDARY_PERIODIC, BOUNDARY_SHEAR = 0, 1, 2, 3
GRAVITY_NONE, GRAVITY_BASIC, GRAVITY_COMPENSATED, GRAVITY_TREE = 0, 1, 2, 3
IGNORE_TERMS_NONE, IGNORE_TERMS_BETWEEN_0_AND_1, IGNORE_TERMS_INVOLVING_0 = 0, 1, 2
INTEGRATOR_NONE, INTEGRATOR_LEAPFROG, INTEGRATOR_SEI = 0, 1, 2
MODE_STRICT, MODE_FAST = 0, 1
EXCHANGE_POSITIONS, EXCHANGE_VELOCITIES, EXCHANGE_ALL = 1, 2, 4
TRANSPORT_AUTO, TRANSPORT_NCCL, TRANSPORT_LOCAL = 0, 1, 2
N_FIELDS = 14  # x y z vx vy vz ax ay az m r name ap sim

ERRORS = {
    -1: "CUDA",
    -2: "ARG",
    -3: "ROOT_SIZE",
    -4: "OUTSIDE_BOX",
    -5: "NONFINITE",
    -6: "SAME_COORDINATES",
    -7: "LEAPFROG_ORDER",
    -8: "CAPACITY",
    -9: "CELL_SIZE_ZERO",
    -10: "NOT_RESIDENT",
}


class Restitution(C.Structure):
    """rebcu_restitution: closed-form restitution law for the device-side resolver."""

    _fields_ = [("kind", C.c_int32), ("pad_", C.c_int32), ("a", C.c_double), ("b", C.c_double), ("c", C.c_double),
                ("lo", C.c_double), ("hi", C.c_double)]


RESTITUTION_CONSTANT, RESTITUTION_POWERLAW = 0, 1


class Vec6d(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("x", "y", "z", "vx", "vy", "vz")]


class ResolvePair(C.Structure):
    """rebcu_resolve_pair: one collision handed to the caller's resolver by rebcu_collision_resolve_pairs."""

    _fields_ = [("k", C.c_uint64), ("p1", C.c_uint64), ("p2", C.c_uint64), ("gb", Vec6d), ("s1", C.c_double * 8), ("s2", C.c_double * 8),
                ("v1", C.c_double * 3), ("v2", C.c_double * 3), ("plog_term", C.c_double), ("logged", C.c_uint64)]


PAIR_RESOLVER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(ResolvePair), C.c_uint64)


class Config(C.Structure):
    """rebcu_config: the scalar fields of struct reb_simulation the hot path reads."""

    _fields_ = [
        ("t", C.c_double),
        ("G", C.c_double),
        ("softening", C.c_double),
        ("OMEGA", C.c_double),
        ("OMEGAZ", C.c_double),
        ("dt", C.c_double),
        ("dt_last_done", C.c_double),
        ("opening_angle2", C.c_double),
        ("root_size", C.c_double),
        ("N_active", C.c_uint64),
        ("testparticle_type", C.c_int32),
        ("gravity_ignore_terms", C.c_int32),
        ("N_root_x", C.c_int32),
        ("N_root_y", C.c_int32),
        ("N_root_z", C.c_int32),
        ("N_ghost_x", C.c_int32),
        ("N_ghost_y", C.c_int32),
        ("N_ghost_z", C.c_int32),
        ("boundary", C.c_int32),
        ("gravity", C.c_int32),
        ("collision", C.c_int32),
        ("integrator", C.c_int32),
        ("leapfrog_order", C.c_int32),
        ("mode", C.c_int32),
        ("quadrupole", C.c_int32),
    ]

    def copy(self):
        c = Config()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(Config))
        return c


def default_config(**kw):
    """Defaults of reb_simulation_create (src/simulation.c:98-122)."""
    c = Config()
    c.t = 0.0
    c.G = 1.0
    c.softening = 0.0
    c.OMEGA = 0.0
    c.OMEGAZ = -1.0
    c.dt = 0.001
    c.dt_last_done = 0.0
    c.opening_angle2 = 0.25
    c.root_size = -1.0
    c.N_active = SIZE_MAX
    c.N_root_x = c.N_root_y = c.N_root_z = 1
    c.gravity = GRAVITY_BASIC
    c.integrator = INTEGRATOR_LEAPFROG
    c.leapfrog_order = 2
    c.mode = MODE_STRICT
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def particles(n):
    return np.zeros(n, dtype=PARTICLE_DTYPE)


def as_ptr(arr):
    return arr.ctypes.data_as(C.c_void_p)


_P = C.c_void_p
_U64P = C.POINTER(C.c_uint64)
_CFG = C.POINTER(Config)
_DBLP = C.POINTER(C.c_double)

# Signatures shared by the CPU checkers (prefix orc_ / refh_); see oracle/oracle.c, oracle/ref_harness.c.
CHECKER_SIGNATURES = {
    "gravity": (C.c_int, [_CFG, _P, _U64P]),
    "gravity_timed": (C.c_int, [_CFG, _P, _U64P, C.c_int, _DBLP]),
    "gravity_cs": (C.c_int, [_CFG, _P, _U64P, _DBLP]),
    "boundary_check": (C.c_int, [_CFG, _P, _U64P]),
    "integrator_step": (C.c_int, [_CFG, _P, _U64P]),
    "collision_search": (C.c_int, [_CFG, _P, C.c_uint64, _P, C.c_uint64, _U64P]),
    "collision_search_subset": (C.c_int, [_CFG, _P, C.c_uint64, _P, C.c_uint64, C.c_uint64, _P, C.c_uint64, _U64P]),
    "steps": (C.c_int, [_CFG, _P, _U64P, C.c_uint64, C.c_int, C.c_double, _DBLP]),
    "exit_check": (C.c_int, [_CFG, _P, C.c_uint64, C.c_double, C.c_double]),
    "apply_jerk": (C.c_int, [_CFG, _P, C.c_uint64, C.c_double]),
    "energy": (C.c_double, [_CFG, _P, C.c_uint64]),
    "com": (None, [_CFG, _P, C.c_uint64, _DBLP]),
    "angular_momentum": (None, [_CFG, _P, C.c_uint64, _DBLP]),
    "tree_dump": (C.c_int, [_CFG, _P, C.c_uint64, _P, C.c_uint64, _U64P]),
    "last_error": (C.c_char_p, []),
    "openmp_threads": (C.c_int, []),
    "set_threads": (None, [C.c_int]),
}

# Signatures of the product C ABI (prefix rebcu_); one entry per declaration in include/rebound_b200.h.
PRODUCT_SIGNATURES = {
    "create": (_P, [C.c_int, _P]),
    "destroy": (None, [_P]),
    "last_error": (C.c_char_p, [_P]),
    "version": (C.c_int, []),
    "device_count": (C.c_int, []),
    "stream": (_P, [_P]),
    "synchronize": (C.c_int, [_P]),
    "host_register": (C.c_int, [_P, C.c_uint64]),
    "host_unregister": (C.c_int, [_P]),
    "upload": (C.c_int, [_P, _P, C.c_uint64]),
    "download": (C.c_int, [_P, _P, C.c_uint64]),
    "download_acc": (C.c_int, [_P, _P, C.c_uint64]),
    "download_gravity_cs": (C.c_int, [_P, _DBLP, C.c_uint64]),
    "N": (C.c_uint64, [_P]),
    "device_field": (_P, [_P, C.c_int]),
    "update_acceleration": (C.c_int, [_P, _CFG]),
    "integrator_step": (C.c_int, [_P, _CFG]),
    "boundary_check": (C.c_int, [_P, _CFG]),
    "collision_search": (C.c_int, [_P, _CFG, _P, C.c_uint64, _U64P]),
    "steps": (C.c_int, [_P, _CFG, C.c_uint64]),
    "collisions_fetch": (C.c_int, [_P, _P, C.c_uint64, _U64P]),
    "set_collision_subset": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64]),
    "tree_build": (C.c_int, [_P, _CFG]),
    "tree_cell_count": (C.c_uint64, [_P]),
    "tree_fetch": (C.c_int, [_P, _P, C.c_uint64]),
    "gravity_host": (C.c_int, [_P, _CFG, _P, _U64P]),
    "collision_search_host": (C.c_int, [_P, _CFG, _P, C.c_uint64, _P, C.c_uint64, _U64P]),
    "steps_host": (C.c_int, [_P, _CFG, _P, _U64P, C.c_uint64]),
    "set_device_resolve": (C.c_int, [_P, C.c_int, C.POINTER(Restitution), C.c_double, C.c_uint]),
    "collision_resolve": (C.c_int, [_P, _CFG]),
    "collision_stats": (C.c_int, [_P, _DBLP, _U64P, C.POINTER(C.c_uint), C.POINTER(C.c_int)]),
    "collision_resolve_pairs": (C.c_int, [_P, C.POINTER(C.c_uint), _P, _P, _DBLP, _U64P, C.POINTER(C.c_int)]),
    "set_shard": (C.c_int, [_P, C.c_int, C.c_int]),
    "shard_range": (None, [_P, _U64P, _U64P]),
    "set_exchange_callback": (C.c_int, [_P, _P, _P]),
    "exchange_request": (C.c_int, [_P]),
    "collisions_segments": (C.c_int, [_P, _U64P, C.c_uint64, _U64P]),
    "set_collision_callback": (C.c_int, [_P, _P, _P]),
    "set_interrupt_flag": (C.c_int, [_P, _P]),
    "energy": (C.c_int, [_P, _CFG, _DBLP]),
    "com": (C.c_int, [_P, _DBLP]),
    "angular_momentum": (C.c_int, [_P, _DBLP]),
    "exit_check": (C.c_int, [_P, C.c_double, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "apply_jerk": (C.c_int, [_P, _CFG, C.c_double]),
    "jerk_host": (C.c_int, [_P, _CFG, _P, C.c_uint64, C.c_double]),
    "measure_fp64_peak": (C.c_int, [_P, _DBLP]),
    "launch_count": (C.c_uint64, [_P]),
    "timing_enable": (C.c_int, [_P, C.c_int]),
    "timing_read": (C.c_int, [_P, _DBLP, _U64P, C.c_int]),
    "timing_reset": (C.c_int, [_P]),
    "tree_walk_stats": (C.c_int, [_P, _CFG, _U64P]),
    "create_group": (_P, [C.POINTER(C.c_int), C.c_int]),
    "group_size": (C.c_int, [_P]),
    "comm_unique_id": (C.c_int, [_P]),
    "exchange": (C.c_int, [_P, C.c_int]),
    "upload_shard": (C.c_int, [_P, _P, C.c_uint64]),
    "download_shard": (C.c_int, [_P, _P, C.c_uint64]),
    "comm_init_rank": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "comm_init_all": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int]),
    "comm_destroy": (C.c_int, [_P]),
    "set_sharded_build": (C.c_int, [_P, C.c_int]),
    "comm_stats": (C.c_int, [_P, _U64P, _U64P, C.POINTER(C.c_int)]),
    "selftest_math": (C.c_int, [_P, C.c_uint64, C.c_uint64, _U64P]),
    "selftest_sort": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_int]),
    "selftest_scan": (C.c_int, [_P, _P, C.c_uint64]),
}


def bind(lib, prefix, signatures):
    """Attach restype/argtypes to every `prefix+name` symbol; missing symbols raise AttributeError."""
    out = {}
    for name, (res, args) in signatures.items():
        fn = getattr(lib, prefix + name)
        fn.restype = res
        fn.argtypes = args
        out[name] = fn
    return out
