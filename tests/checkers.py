"""CPU checkers for the parity tests: the oracle restatement (oracle/liboracle.so) and, when it has
been built in this container, the unmodified reference behind oracle/ref_harness.c
(oracle/_ref/libref_harness*.so).  TEST INFRASTRUCTURE ONLY -- the product never imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

from rebound_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class CheckerError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.msg = msg


class Checker:
    """Uniform Python face of liboracle.so (prefix orc_) and libref_harness*.so (prefix refh_)."""

    def __init__(self, path, prefix, kind):
        self.lib = C.CDLL(path)
        self.f = abi.bind(self.lib, prefix, abi.CHECKER_SIGNATURES)
        self.kind = kind
        self.path = path

    def _check(self, err):
        if err != 0:
            raise CheckerError(err, self.f["last_error"]().decode())

    def gravity(self, cfg, p):
        p = p.copy()
        c = cfg.copy()
        n = C.c_uint64(len(p))
        self._check(self.f["gravity"](C.byref(c), abi.as_ptr(p), C.byref(n)))
        return p[: n.value], c

    def gravity_cs(self, cfg, p):
        """(particles, r->gravity_cs as (N,3)) after one COMPENSATED force evaluation."""
        p = p.copy()
        c = cfg.copy()
        n = C.c_uint64(len(p))
        cs = np.zeros((max(len(p), 1), 3), dtype=np.float64)
        self._check(self.f["gravity_cs"](C.byref(c), abi.as_ptr(p), C.byref(n), cs.ctypes.data_as(C.POINTER(C.c_double))))
        return p[: n.value], cs[: n.value]

    def gravity_timed(self, cfg, p, n_evals=1):
        p = p.copy()
        c = cfg.copy()
        n = C.c_uint64(len(p))
        sec = C.c_double(0)
        self._check(self.f["gravity_timed"](C.byref(c), abi.as_ptr(p), C.byref(n), n_evals, C.byref(sec)))
        return p[: n.value], sec.value

    def boundary_check(self, cfg, p):
        p = p.copy()
        c = cfg.copy()
        n = C.c_uint64(len(p))
        self._check(self.f["boundary_check"](C.byref(c), abi.as_ptr(p), C.byref(n)))
        return p[: n.value], c

    def integrator_step(self, cfg, p):
        p = p.copy()
        c = cfg.copy()
        n = C.c_uint64(len(p))
        self._check(self.f["integrator_step"](C.byref(c), abi.as_ptr(p), C.byref(n)))
        return p[: n.value], c

    def collision_search(self, cfg, p, cap=None):
        p = p.copy()
        c = cfg.copy()
        cap = cap or max(64, 8 * len(p))
        out = np.zeros(cap, dtype=abi.COLLISION_DTYPE)
        n = C.c_uint64(0)
        self._check(self.f["collision_search"](C.byref(c), abi.as_ptr(p), len(p), abi.as_ptr(out), cap, C.byref(n)))
        if n.value > cap:
            return self.collision_search(cfg, p, cap=n.value)
        return out[: n.value]

    def collision_search_subset(self, cfg, p, map=None, n_targets=None, cap=None):
        """reb_collision_search with r->map / r->N_map / r->N_targets set (collision.c:53-58)."""
        p = p.copy()
        c = cfg.copy()
        cap = cap or max(64, 8 * len(p))
        out = np.zeros(cap, dtype=abi.COLLISION_DTYPE)
        n = C.c_uint64(0)
        m = None if map is None else np.ascontiguousarray(map, dtype=np.uint64)
        self._check(self.f["collision_search_subset"](
            C.byref(c), abi.as_ptr(p), len(p), None if m is None else abi.as_ptr(m), 0 if m is None else len(m),
            abi.SIZE_MAX if n_targets is None else n_targets, abi.as_ptr(out), cap, C.byref(n)))
        if n.value > cap:
            return self.collision_search_subset(cfg, p, map, n_targets, cap=n.value)
        return out[: n.value]

    def steps(self, cfg, p, n_steps, resolve=0, minimum_collision_velocity=0.0):
        p = p.copy()
        c = cfg.copy()
        n = C.c_uint64(len(p))
        aux = (C.c_double * 3)()
        self._check(self.f["steps"](C.byref(c), abi.as_ptr(p), C.byref(n), n_steps, resolve,
                                    minimum_collision_velocity, aux))
        return p[: n.value], c, {"collisions_log_n": int(aux[0]), "collisions_plog": aux[1], "seconds": aux[2]}

    def apply_jerk(self, cfg, p, v):
        """reb_gravity_basic_calculate_and_apply_jerk (gravity.c:850-924): velocities after the jerk kick."""
        p = p.copy()
        self._check(self.f["apply_jerk"](C.byref(cfg.copy()), abi.as_ptr(p), len(p), float(v)))
        return p

    def exit_check(self, cfg, p, exit_max_distance=0.0, exit_min_distance=0.0):
        """Status after run_heartbeat's exit checks: 4 escape, 3 encounter (wins over escape), 0 neither."""
        p = p.copy()
        return self.f["exit_check"](C.byref(cfg.copy()), abi.as_ptr(p), len(p), float(exit_max_distance), float(exit_min_distance))

    def energy(self, cfg, p):
        p = p.copy()
        c = cfg.copy()
        return self.f["energy"](C.byref(c), abi.as_ptr(p), len(p))

    def com(self, cfg, p):
        p = p.copy()
        out = (C.c_double * 10)()
        self.f["com"](C.byref(cfg.copy()), abi.as_ptr(p), len(p), out)
        return dict(zip(("m", "x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az"), out))

    def angular_momentum(self, cfg, p):
        p = p.copy()
        out = (C.c_double * 3)()
        self.f["angular_momentum"](C.byref(cfg.copy()), abi.as_ptr(p), len(p), out)
        return out[0], out[1], out[2]

    def tree_dump(self, cfg, p):
        p = p.copy()
        c = cfg.copy()
        cap = 4 * len(p) + 64
        while True:
            out = np.zeros(cap, dtype=abi.TREECELL_DTYPE)
            n = C.c_uint64(0)
            self._check(self.f["tree_dump"](C.byref(c), abi.as_ptr(p), len(p), abi.as_ptr(out), cap, C.byref(n)))
            if n.value <= cap:
                return out[: n.value]
            cap = n.value

    def gravity_rows(self, cfg, p, rows):
        """Oracle port only: accelerations (n_rows, 3) of the given particles from all their sources (direct sum)."""
        fn = self.lib.orc_gravity_rows
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(abi.Config), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_double)]
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.zeros((max(len(rows), 1), 3), dtype=np.float64)
        self._check(fn(C.byref(cfg.copy()), abi.as_ptr(p), len(p), abi.as_ptr(rows), len(rows), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out[: len(rows)]

    def tree_session(self, cfg, p):
        """Reference only: the serial phases of one tree force evaluation on the full problem, each timed, and a tree
        that stays alive for sampled walks (oracle/ref_harness.c: refh_tree_open / _walk_sample / _close)."""
        return TreeSession(self, cfg, p)

    def threads(self):
        return self.f["openmp_threads"]()

    def set_threads(self, n):
        self.f["set_threads"](n)


class TreeSession:
    def __init__(self, chk, cfg, p):
        lib = chk.lib
        lib.refh_tree_open.restype = C.c_void_p
        lib.refh_tree_open.argtypes = [C.POINTER(abi.Config), C.c_void_p, C.c_uint64, C.POINTER(C.c_double)]
        lib.refh_tree_session_N.restype = C.c_uint64
        lib.refh_tree_session_N.argtypes = [C.c_void_p]
        lib.refh_tree_walk_sample.restype = C.c_int
        lib.refh_tree_walk_sample.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        lib.refh_tree_sample_acc.restype = C.c_int
        lib.refh_tree_sample_acc.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_double), C.c_uint64]
        lib.refh_tree_close.restype = C.c_int
        lib.refh_tree_close.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        self.lib = lib
        sec = (C.c_double * 3)()
        c = cfg.copy()
        self.h = lib.refh_tree_open(C.byref(c), abi.as_ptr(p), len(p), sec)
        if not self.h:
            raise CheckerError(-1, chk.f["last_error"]().decode())
        self.seconds = {"boundary": sec[0], "construct": sec[1], "gravity_data": sec[2]}
        self.N = int(lib.refh_tree_session_N(self.h))

    def walk_sample(self, stride, offset=0):
        """Walks every stride-th particle from `offset`: (seconds, number walked)."""
        sec = C.c_double(0)
        n = C.c_uint64(0)
        self.lib.refh_tree_walk_sample(self.h, stride, offset, C.byref(sec), C.byref(n))
        return sec.value, int(n.value)

    def sample_acc(self, stride, offset=0):
        n = (self.N - offset + stride - 1) // stride if offset < self.N else 0
        out = np.zeros((max(n, 1), 3), dtype=np.float64)
        k = self.lib.refh_tree_sample_acc(self.h, stride, offset, out.ctypes.data_as(C.POINTER(C.c_double)), n)
        return out[:k]

    def close(self):
        """{delete, rest}: reb_tree_delete, and one reb_simulation_steps(r,1) without gravity (drift, kick, drift,
        boundary check)."""
        if self.h:
            sec = (C.c_double * 2)()
            self.lib.refh_tree_close(self.h, sec)
            self.h = None
            self.seconds.update({"delete": sec[0], "rest": sec[1]})
        return self.seconds


def build_oracle():
    """Compiles oracle/liboracle.so (and oracle/_ref when the reference sources are present)."""
    subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True, capture_output=True)


_cache = {}


def oracle():
    if "oracle" not in _cache:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, capture_output=True)
        _cache["oracle"] = Checker(path, "orc_", "port")
    return _cache["oracle"]


def reference(openmp=False, quadrupole=False):
    """The unmodified reference, or None if oracle/_ref has not been built (it is built in the
    authoring container, where /root/reference exists, and travels to the GPU box as a binary).
    quadrupole=True: the serial build compiled with -DQUADRUPOLE."""
    key = "ref_quad" if quadrupole else ("ref_omp" if openmp else "ref")
    if key not in _cache:
        name = "libref_harness_quad.so" if quadrupole else ("libref_harness_omp.so" if openmp else "libref_harness.so")
        path = os.path.join(ORACLE_DIR, "_ref", name)
        _cache[key] = Checker(path, "refh_", "reference") if os.path.exists(path) else None
    return _cache[key]


DOUBLE_FIELDS = ("x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az", "m", "r")


def bits_equal(a, b, fields=DOUBLE_FIELDS):
    """True iff the given double fields agree bit for bit (NaN-safe, distinguishes -0.0)."""
    if len(a) != len(b):
        return False
    return all(np.array_equal(a[f].view(np.uint64), b[f].view(np.uint64)) for f in fields)


def max_rel_acc_error(a, b):
    """max_i |a_i - b_i| / |b_i| over acceleration vectors."""
    da = np.stack([a["ax"] - b["ax"], a["ay"] - b["ay"], a["az"] - b["az"]], axis=1)
    nb = np.sqrt(b["ax"] ** 2 + b["ay"] ** 2 + b["az"] ** 2)
    nd = np.sqrt((da**2).sum(axis=1))
    ok = nb > 0
    return float((nd[ok] / nb[ok]).max()) if ok.any() else 0.0


def collisions_equal(a, b, with_ri=True):
    """Bitwise list equality.  `ri` is left uninitialised by the reference's DIRECT search
    (src/collision.c:114-117 never writes it), so it is only compared for TREE lists."""
    if len(a) != len(b):
        return False
    fields = [f for f in abi.COLLISION_DTYPE.names if with_ri or f != "ri"]
    return all(np.array_equal(a[f].view(np.uint64), b[f].view(np.uint64)) for f in fields)
