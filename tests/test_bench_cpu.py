"""bench.py on the CPU: the reference arm (`--impl reference`) prints one JSON line with the contract's keys on the same
`config` the GPU arm prints, and the workload table matches BASELINE.json's configurations."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.needs_ref


def run_reference_arm(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", *extra],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_tree_headline_at_reduced_size():
    """The headline workload (C4, tree) at 2^16: serial phases timed on the full problem, sampled OpenMP walks."""
    d = run_reference_arm("--n-log2", "16")
    assert d["impl"] == "reference" and d["metric"] == "particle-steps/s (tree)" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert set(d["config"]) == {"workload", "N", "inner_steps_per_step"} and d["config"]["N"] == 1 << 16
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    phases = d["cpu_baseline"]["detail"]["phases_s"]
    assert set(phases) == {"boundary", "construct", "gravity_data", "delete", "rest"}
    # value = N / (fixed phases + scaled sample walk): recomputable from the line
    det = d["cpu_baseline"]["detail"]
    fixed = sum(phases.values())
    est = [fixed + w * det["N"] / det["walk_sample_particles"] for w in det["walk_sample_s"][1:]]
    assert abs(d["value"] - det["N"] * len(est) / sum(est)) / d["value"] < 5e-3


def test_gpu_arm_and_reference_arm_describe_the_same_config():
    import bench

    for name, lg in (("c4", 16), ("c3", 12), ("c1", 0)):
        w = bench.workload(name, lg)
        assert w["N"] == len(w["p"])
        cfg = {"workload": w["desc"], "N": int(w["N"]), "inner_steps_per_step": w["inner"]}
        assert json.dumps(cfg)          # what both arms print under "config"
    w = bench.workload("c2")
    assert w["N"] == (1 << 20) + 10 and w["cfg"].N_active == 10 and w["cfg"].testparticle_type == 0
    assert bench.workload("c5", 14)["cfg"].N_ghost_x == 2
