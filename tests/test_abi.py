"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/rebound_b200.h declares, the struct mirrors have the reference's sizes, and the product
fails loudly without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from rebound_b200 import abi, simulation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "rebound_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rebcu_[a-z_A-Z0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    names = declared_functions()
    assert len(names) >= 30
    assert sorted("rebcu_" + n for n in abi.PRODUCT_SIGNATURES) == names


def test_library_exports_every_declared_symbol():
    fn = simulation.load_library()
    lib = C.CDLL(simulation.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert fn["version"]() >= 100


def test_struct_sizes_match_reference_layout():
    assert abi.PARTICLE_DTYPE.itemsize == 112      # struct reb_particle, rebound.h:86-104
    assert abi.COLLISION_DTYPE.itemsize == 72      # struct reb_collision, rebound.h:144-149
    assert C.sizeof(abi.Config) == 144             # 9 doubles + uint64 + 15 int32, padded to 8 bytes (rebcu_config)


def test_no_cpu_fallback():
    fn = simulation.load_library()
    if fn["device_count"]() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(simulation.ReboundCudaError):
        simulation.Engine(0)
    with pytest.raises(simulation.ReboundCudaError):
        simulation.Simulation()


def test_product_does_not_reference_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "rebound_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "liboracle" not in text and "libref_harness" not in text, f


def test_dropin_librebound_exports_reference_symbols():
    """The drop-in librebound (when built) exports the hot-path symbols under the reference's names and,
    like the reference's CI demands (.github/workflows/c.yml:17-22), only reb_-prefixed symbols."""
    import subprocess
    lib = os.path.join(ROOT, "rebound_b200", "_dropin", "librebound.so")
    if not os.path.exists(lib):
        pytest.skip("drop-in librebound not built")
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if l.split()[-2] in "TDRB"]
    assert all(s.startswith("reb_") for s in syms), [s for s in syms if not s.startswith("reb_")]
    for s in ("reb_gravity_basic_calculate_acceleration", "reb_gravity_compensated_calculate_acceleration",
              "reb_gravity_tree_calculate_acceleration", "reb_boundary_check", "reb_collision_search",
              "reb_integrator_leapfrog", "reb_integrator_sei", "reb_simulation_integrate", "reb_tree_construct",
              "reb_boundary_get_ghostbox", "reb_integrator_leapfrog_lf4_a"):
        assert s in syms, s
