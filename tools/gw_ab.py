#!/usr/bin/env python
"""A/B of the group-walk kernel variants (REBOUND_B200_GW_VARIANT, read once per process): runs tools/measure.py once
per variant and prints walk time, groups that finished as groups, list entries.
usage: python tools/gw_ab.py <variants, e.g. -hijklmnopq> <cases...>   ('-' = the default kernel)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(variant, cases):
    env = dict(os.environ)
    env.pop("REBOUND_B200_GW_VARIANT", None)
    if variant != "-":
        env["REBOUND_B200_GW_VARIANT"] = variant
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "measure.py")] + cases, env=env, capture_output=True, text=True, timeout=900)
    rows = []
    for line in r.stdout.splitlines():
        try:
            d = json.loads(line)
        except Exception:
            continue
        ws = d.get("walk_stats", {})
        rows.append({"variant": variant, "case": d["case"], "walk_ms": d["kernel_ms_per_step"].get("treewalk"), "build_ms": d["kernel_ms_per_step"].get("treebuild"),
                     "groups": ws.get("groups"), "entries": ws.get("group_entries"), "visited": ws.get("group_visits"),
                     "interactions": ws.get("interactions")})
    if r.returncode != 0:
        rows.append({"variant": variant, "error": r.stderr[-400:]})
    return rows


def main():
    variants, cases = sys.argv[1], sys.argv[2:]
    for v in variants:
        for row in run(v, cases):
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
