"""Prints the per-range timeline of the chunk-pipelined host path (rebcu_steps_host on C2): REBOUND_B200_PIPE_TRACE=1.
usage: [REBOUND_B200_CHUNKS=n] python tools/pipe_trace.py"""
import os, sys, torch, numpy as np
sys.path.insert(0, ".")
from rebound_b200 import ics, abi
from rebound_b200.simulation import Engine
st = torch.cuda.Stream()
eng = Engine(0, st.cuda_stream)
n = 1 << 20
p = ics.planetesimal_disk(n); cfg = ics.planetesimal_config()
host = torch.empty(len(p) * 112, dtype=torch.uint8, pin_memory=True)
hp = host.numpy().view(abi.PARTICLE_DTYPE); hp[:] = p
for _ in range(3): eng.steps_host(cfg.copy(), hp, 100)
os.environ["REBOUND_B200_PIPE_TRACE"] = "1"
eng.steps_host(cfg.copy(), hp, 100)
del os.environ["REBOUND_B200_PIPE_TRACE"]
import time
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10): eng.steps_host(cfg.copy(), hp, 100)
torch.cuda.synchronize(); print("ms/call", (time.perf_counter()-t)*100)
