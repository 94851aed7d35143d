"""The reference's own Python package (ctypes) on top of the drop-in librebound: `import rebound` finds the
library through importlib (rebound/__init__.py:33-38), checks sizeof(struct reb_simulation)
(rebound/simulation.py:1478-1482) and raises RuntimeError from reb_simulation_error messages
(rebound/simulation.py:259-266).  Needs /root/reference (not present on the GPU box), runs without a GPU: paths
that stay on the reference's C code work, the replaced hot path refuses to run on the CPU."""
import os
import shutil
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "rebound_b200", "_dropin", "librebound.so")
REF = "/root/reference"

pytestmark = [pytest.mark.needs_ref,
              pytest.mark.skipif(not (os.path.exists(DROPIN) and os.path.isdir(os.path.join(REF, "rebound"))),
                                 reason="needs the reference checkout and the built drop-in library")]

SCRIPT = r"""
import rebound
print("LIB", rebound.__libpath__)
sim = rebound.Simulation()
sim.add(m=1.); sim.add(m=1e-3, a=1.); sim.add(m=1e-3, a=2.3)
sim.integrator = "whfast"
sim.integrator.kernel = "lazy"         # Jacobi-coordinate forces: reb_gravity_jacobi_*, not a replaced symbol
sim.dt = 0.01
e0 = sim.energy()
sim.integrate(3.0)
print("WHFAST", sim.t, abs((sim.energy() - e0) / e0))
sim2 = rebound.Simulation()
sim2.add(m=1.); sim2.add(m=1e-3, a=1.)
sim2.integrator = "leapfrog"
sim2.dt = 0.01
try:
    sim2.steps(1)
    print("LEAPFROG ran")
except RuntimeError as e:
    print("LEAPFROG RuntimeError:", e)
"""


def test_reference_python_package_loads_the_dropin(tmp_path):
    lib_dir = tmp_path / "pkg" / "_dropin"
    lib_dir.mkdir(parents=True)
    shutil.copy(DROPIN, lib_dir / "librebound.so")
    shutil.copy(os.path.join(ROOT, "rebound_b200", "librebound_b200.so"), tmp_path / "pkg" / "librebound_b200.so")
    env = dict(os.environ, PYTHONPATH=f"{lib_dir}:{REF}")
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert f"LIB {lib_dir}" in out
    whfast = [l for l in out.splitlines() if l.startswith("WHFAST")][0].split()
    assert float(whfast[1]) >= 3.0 and float(whfast[2]) < 1e-6
    if torch.cuda.is_available():
        assert "LEAPFROG ran" in out
    else:
        assert "LEAPFROG RuntimeError:" in out and "no usable CUDA device" in out and "no CPU fallback" in out
