#!/usr/bin/env python
"""Per-kernel-class device timings of the other BASELINE.json configurations (not the bench line).
usage: python tools/measure.py [c1|c1fast|c3s|c4_20|c4_22|c4_24|c5_18|c5_20|c5r_20 ...]
(c5r_N: C5 with the device-side hard-sphere resolve, Bridges restitution)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from rebound_b200 import abi, ics  # noqa: E402
from rebound_b200.simulation import Engine  # noqa: E402


def case(name):
    if name.startswith("c1"):
        n = 16384
        cfg = ics.plummer_config(n)
        if "fast" in name:
            cfg.mode = abi.MODE_FAST
        return ics.plummer(n, seed=42), cfg, 5, n * n - n
    if name.startswith("d"):            # direct BASIC, Plummer sphere, N = 2^k  (d17, d17fast, ...)
        n = 1 << int(name[1:].replace("fast", ""))
        cfg = ics.plummer_config(n)
        if "fast" in name:
            cfg.mode = abi.MODE_FAST
        return ics.plummer(n, seed=42), cfg, 2, n * n - n
    if name.startswith("c3s"):          # C3 recipe on one GPU at reduced N
        n = 1 << int(name.split("_")[1].replace("fast","")) if "_" in name else 1 << 18
        cfg = ics.plummer_config(n, gravity=abi.GRAVITY_COMPENSATED)
        if "fast" in name:
            cfg.mode = abi.MODE_FAST
        return ics.plummer(n, seed=42), cfg, 1, n * n - n
    if name.startswith("c4"):
        n = 1 << int(name.split("_")[1].replace("fast",""))
        cfg = ics.selfgravity_disc_config()
        if "fast" in name:
            cfg.mode = abi.MODE_FAST
        return ics.selfgravity_disc(n - 1, seed=42), cfg, 2, None
    if name.startswith("c5"):
        n = 1 << int(name.split("_")[1].replace("fast",""))
        rs = 2655.0 * (n / 2**20) ** 0.5          # root_size ~ 2655 m gives N ~ 2^20 with 2x2 root boxes
        cfg = ics.shearing_sheet_config(root_size=rs)
        p = ics.shearing_sheet(root_size=rs, seed=42)
        if "fast" in name:
            cfg.mode = abi.MODE_FAST
        return p, cfg, 2, None
    raise SystemExit(name)


def main():
    names = sys.argv[1:] or ["c1", "c1fast", "c4_20"]
    eng = Engine(0)
    for name in names:
        p, cfg, steps, inter = case(name)
        eng.upload(np.ascontiguousarray(p))
        c = cfg.copy()
        device_resolve = name.startswith("c5r")
        if device_resolve:
            eng.set_device_resolve(True, restitution=(0.32, 100.0, -0.234, 0.0, 1.0),
                                   minimum_collision_velocity=1.0 * ics.SHEET_OMEGA * 0.001, rand_seed=42)
        eng.steps(c, 1)
        eng.synchronize()
        eng.timing_enable(True)
        eng.timing_reset()
        t0 = time.perf_counter()
        eng.steps(c, steps)
        eng.synchronize()
        wall = time.perf_counter() - t0
        tim = eng.timing_read()
        eng.timing_enable(False)
        out = {"case": name, "N": int(len(p)), "N_after": eng.N, "steps": steps, "wall_ms_per_step": 1e3 * wall / steps,
               "particle_steps_per_s": len(p) * steps / wall,
               "kernel_ms_per_step": {k: round(v["ms"] / steps, 4) for k, v in tim.items() if v["launches"]},
               "launches_per_step": {k: v["launches"] / steps for k, v in tim.items() if v["launches"]}}
        if inter:
            out["interactions_per_s"] = inter * steps / wall
        if cfg.gravity == abi.GRAVITY_TREE:
            eng.update_acceleration(c)
            out["walk_stats"] = eng.tree_walk_stats(c)
        if device_resolve:
            out["resolve"] = eng.collision_stats()
            eng.set_device_resolve(False)
        elif cfg.collision:
            out["collisions_last_step"] = int(len(eng.collisions_fetch()))
        print(json.dumps(out), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
