#!/usr/bin/env python
"""Wall-clock probe of the host-buffer path (rebcu_steps_host) on config C2 with a pinned AoS.
usage: [REBOUND_B200_CHUNKS=n] [REBOUND_B200_PIPE_TRACE=1] python tools/pipe_probe.py [inner_steps] [calls]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rebound_b200 import abi, ics  # noqa: E402
from rebound_b200.simulation import Engine  # noqa: E402


def main():
    inner = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    p = ics.planetesimal_disk(1 << 20, seed=42)
    cfg = ics.planetesimal_config()
    host = torch.empty(len(p) * abi.PARTICLE_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
    hp = host.numpy().view(abi.PARTICLE_DTYPE)
    hp[:] = p
    eng = Engine(0)
    c = cfg.copy()
    for _ in range(3):
        eng.steps_host(c, hp, inner)
    trace = os.environ.pop("REBOUND_B200_PIPE_TRACE", None)
    ts = []
    for _ in range(calls):
        t0 = time.perf_counter()
        eng.steps_host(c, hp, inner)
        ts.append(1e3 * (time.perf_counter() - t0))
    eng.upload(hp)
    eng.steps(c, inner)
    eng.synchronize()
    t0 = time.perf_counter()
    for _ in range(calls):
        eng.steps(c, inner)
    eng.synchronize()
    res = 1e3 * (time.perf_counter() - t0) / calls
    print(f"chunks={os.environ.get('REBOUND_B200_CHUNKS', 'default')} inner={inner}: host-buffer call median {np.median(ts):.3f} ms "
          f"min {min(ts):.3f} ms; resident {res:.3f} ms", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
