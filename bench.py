#!/usr/bin/env python
"""bench.py -- benchmark of the hot path (contract: DESIGN.md "Measurement").

Headline workload (every --gpus N): BASELINE.json configs[3] ("c4"): self-gravitating disc N = 2^24,
REB_GRAVITY_TREE opening_angle2 = 0.25, leapfrog, open boundary -- ONE problem, strong scaling: rank r owns the
contiguous target block [N r/W, N (r+1)/W), positions are all-gathered over NCCL inside the engine between drift and
force (csrc/comm.cu).  One bench step = one reb_simulation_steps(r, 1).  Metric: particle-steps/s = N / step time.

  value   device-resident: particles already in HBM, CUDA events on the engine's stream around K steps, max over ranks.
  e2e     the same steps through the host-buffer C-ABI calls the drop-in shim makes: pinned host AoS -> H2D -> step ->
          D2H every step (N=1: rebcu_steps_host on the whole array; N>1: every rank moves its own block,
          rebcu_upload_shard / rebcu_download_shard, the rest travels over NVLink).
  configs the other BASELINE.json configurations, each with value / e2e / roofline / cpu_baseline:
          c1 Plummer 16384 BASIC, c2 10 massive + 2^20 test particles, c5 shearing sheet 2^20 (N=1 only);
          c3 Plummer COMPENSATED direct sum sharded with the NCCL position all-gather (every N; reduced N, law stated).
  --impl reference : the reference's own CPU implementation (oracle/_ref: the unmodified reference compiled with
          OpenMP) on the headline workload, all host threads, each step a bounded sample (sampled tree walk on the full
          2^24-particle tree, serial phases timed in full).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

BRIDGES = (0.32, 100.0, -0.234, 0.0, 1.0)        # examples/shearing_sheet/problem.c:96-103
FLOP_PER_INTERACTION = {"basic": 20.0, "compensated": 29.0}
# FP64-pipe instructions per pair term (SASS counts): strict = IEEE sqrt + divide expanded
DP_PER_PAIR = {("basic", "strict"): 36.0, ("basic", "fast"): 16.0, ("compensated", "strict"): 45.0, ("compensated", "fast"): 25.0}


# ------------------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------------------
def workload(name, n_log2=0):
    """BASELINE.json configs -> dict(p, cfg, N, inner, units (metric units per inner step), metric, unit, desc, kind)."""
    from rebound_b200 import abi, ics

    if name == "c1":
        n = 16384
        return dict(p=ics.plummer(n, seed=42), cfg=ics.plummer_config(n), N=n, inner=100, units=float(n) * n - n,
                    metric="pairwise interactions/s (direct)", unit="interactions/s", kind="basic",
                    desc="C1 Plummer sphere N=16384, REB_GRAVITY_BASIC, leapfrog, 100 steps per call")
    if name == "c2":
        p = ics.planetesimal_disk(1 << 20, seed=42)
        n = len(p)
        return dict(p=p, cfg=ics.planetesimal_config(), N=n, inner=100, units=float(n) * 10 - 10,
                    metric="pairwise interactions/s (direct)", unit="interactions/s", kind="basic",
                    desc="C2 planetesimal disk: N_active=10 + 2^20 test particles, REB_GRAVITY_BASIC, testparticle_type=0, leapfrog, 100 steps per call")
    if name == "c3":
        lg = n_log2 or 20
        n = 1 << lg
        note = "" if lg == 22 else f" (C3 recipe at N=2^{lg} instead of 2^22: the pair rate does not depend on N once the GPU is full; cost law N^2)"
        return dict(p=ics.plummer(n, seed=42), cfg=ics.plummer_config(n, gravity=abi.GRAVITY_COMPENSATED), N=n, inner=1,
                    units=float(n) * n - n, metric="pairwise interactions/s (direct)", unit="interactions/s", kind="compensated",
                    desc=f"C3 Plummer sphere N=2^{lg}, REB_GRAVITY_COMPENSATED direct summation, leapfrog, target blocks sharded with an NCCL position all-gather" + note)
    if name == "c4":
        lg = n_log2 or 24
        n = 1 << lg
        return dict(p=ics.selfgravity_disc(n - 1, seed=42), cfg=ics.selfgravity_disc_config(), N=n, inner=1, units=float(n),
                    metric="particle-steps/s (tree)", unit="particle-steps/s", kind="tree",
                    desc=f"C4 self-gravitating disc N=2^{lg}, REB_GRAVITY_TREE opening_angle2=0.25, leapfrog, open boundary (examples/selfgravity_disc scaled up)")
    if name == "c5":
        lg = n_log2 or 20
        rs = 2655.0 * (2.0 ** (lg - 20)) ** 0.5        # root_size ~ 2655 m gives N ~ 2^20 with 2x2 root boxes
        p = ics.shearing_sheet(root_size=rs, seed=42)
        n = len(p)
        return dict(p=p, cfg=ics.shearing_sheet_config(root_size=rs), N=n, inner=1, units=float(n),
                    metric="particle-steps/s (tree)", unit="particle-steps/s", kind="sheet", root_size=rs,
                    desc=f"C5 shearing sheet N~2^{lg} ({n} particles), REB_GRAVITY_TREE + REB_COLLISION_TREE, 25 ghost boxes, SEI, hard-sphere resolve with Bridges restitution (examples/shearing_sheet scaled up)")
    raise SystemExit(f"unknown workload {name}")


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
def _nvml_handle(gpu_index):
    """NVML handle of the CUDA device `gpu_index` of this process (honours a numeric CUDA_VISIBLE_DEVICES)."""
    import pynvml

    pynvml.nvmlInit()
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    idx = gpu_index
    try:
        ids = [int(x) for x in vis.split(",") if x.strip() != ""]
        if ids:
            idx = ids[gpu_index]
    except ValueError:
        pass
    return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)


class ClockSampler:
    """Samples the SM clock and the clock event (throttle) reasons of one GPU DURING the timed region: NVML every
    2 ms from a thread; falls back to an nvidia-smi loop."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.samples = []          # (sm_mhz, reasons bit mask)
        self.max_mhz = None
        self.stop_flag = False

    def start(self):
        try:
            self.nvml, self.handle = _nvml_handle(self.gpu)
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)), int(reasons_fn(self.handle))))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            sm = [s for s, _ in self.samples]
            mask = 0
            for _, m in self.samples:
                mask |= m
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(k for k, bit in self.BITS.items() if mask & bit), "samples": len(sm), "source": "nvml, 2 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def bind_to_gpu_numa_node(gpu_index):
    """One process per GPU: run on the cores next to this GPU so that the pinned host buffers (first touch) and the
    copy threads are local to its PCIe root."""
    try:
        nvml, handle = _nvml_handle(gpu_index)
        bus = nvml.nvmlDeviceGetPciInfo(handle).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:            # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------------------
# the reference's CPU path (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_checker():
    import checkers

    ref = checkers.reference(openmp=True)
    if ref is not None:
        return ref, "reference"
    return checkers.oracle(), "port"


def cpu_tree_session(w, chk, n_walks, stride):
    """C4 on the CPU: the serial phases of one step on the FULL problem, timed once, + n_walks sampled OpenMP walks.
    Returns (list of estimated full-step seconds, detail)."""
    s = chk.tree_session(w["cfg"], w["p"])
    walks = []
    for k in range(n_walks):
        sec, n_s = s.walk_sample(stride, k % stride)
        walks.append((sec, n_s))
    t = s.close()
    fixed = t["boundary"] + t["construct"] + t["gravity_data"] + t["delete"] + t["rest"]
    est = [fixed + sec * (s.N / max(1, n_s)) for sec, n_s in walks]
    detail = {"phases_s": {k: round(v, 4) for k, v in t.items()}, "walk_sample_particles": walks[0][1] if walks else 0,
              "walk_sample_s": [round(x[0], 4) for x in walks], "N": s.N,
              "law": "step = boundary + construct + gravity_data + delete + rest (each timed once on the full problem) + sampled OpenMP walk x N / n_sample"}
    return est, detail


def cpu_baseline(name, w, budget_s=8.0):
    """The reference's OpenMP build on the box's host cores, a bounded sample of workload `name`."""
    from rebound_b200 import abi, ics

    chk, kind = cpu_reference_checker()
    cores = os.cpu_count() or 1
    chk.set_threads(cores)
    cfg, p = w["cfg"], w["p"]
    if name in ("c1", "c2"):
        chk.steps(cfg, p, 1)                  # thread pool start-up, page faults
        _, _, aux = chk.steps(cfg, p, 2)
        n_cpu = max(1, min(20 * w["inner"], int(budget_s / max(aux["seconds"] / 2, 1e-6))))
        _, _, aux = chk.steps(cfg, p, n_cpu)
        return {"value": w["units"] * n_cpu / aux["seconds"], "unit": w["unit"], "cores": chk.threads(), "kind": kind,
                "sample": f"reb_simulation_steps(r,{n_cpu}) on the full workload ({aux['seconds']:.1f} s)"}
    if name == "c3":
        # all N targets, the first N_active = n_src particles as sources (testparticle_type 0): the reference's own
        # compensated pair loop on a bounded number of pairs; sources stay cache resident, which favours the CPU
        n = w["N"]
        c = cfg.copy()
        n_src = 4096
        c.N_active = n_src
        c.testparticle_type = 0
        _, sec = chk.gravity_timed(c, p, 1)
        reps = max(1, min(8, int(budget_s / max(sec, 1e-6))))
        _, sec = chk.gravity_timed(c, p, reps)
        pairs = float(n) * n_src - n_src
        return {"value": pairs / sec, "unit": w["unit"], "cores": chk.threads(), "kind": kind,
                "sample": f"{reps} force evaluations of all N=2^{int(np.log2(n))} targets against N_active={n_src} sources "
                          f"({pairs:.3g} pairs each, {sec:.2f} s): the unmodified compensated loop on a bounded pair count"}
    if name == "c4":
        if kind != "reference":
            # no compiled reference on this box: full steps of the port at reduced N
            small = workload("c4", 18)
            _, _, aux = chk.steps(small["cfg"], small["p"], 1)
            return {"value": small["N"] / aux["seconds"], "unit": w["unit"], "cores": chk.threads(), "kind": kind,
                    "sample": f"one full step at N=2^18 ({aux['seconds']:.1f} s); particle-steps/s falls with log N"}
        est, detail = cpu_tree_session(w, chk, 3, 256)
        sec = statistics.median(est)
        return {"value": w["N"] / sec, "unit": w["unit"], "cores": chk.threads(), "kind": kind,
                "sample": f"estimated {sec:.1f} s per step at N={w['N']}: " + detail["law"], "detail": detail}
    if name == "c5":
        # full steps with the reference's hard-sphere resolver at N ~ 2^17 (a full 2^20 step takes the CPU the better
        # part of a minute); per-particle cost grows ~ log N, so the rate at 2^20 is lower than reported here
        small = workload("c5", 17)
        min_v = ics.SHEET_OMEGA * 0.001
        chk.steps(small["cfg"], small["p"], 1, resolve=2, minimum_collision_velocity=min_v)
        _, _, aux = chk.steps(small["cfg"], small["p"], 2, resolve=2, minimum_collision_velocity=min_v)
        return {"value": small["N"] * 2 / aux["seconds"], "unit": w["unit"], "cores": chk.threads(), "kind": kind,
                "sample": f"2 full steps (SEI + tree gravity with 25 ghost boxes + tree collision search + hard-sphere resolve) at "
                          f"N={small['N']} ({aux['seconds']:.1f} s); cost per particle grows ~log N, so this is an upper bound for N~2^20"}
    raise ValueError(name)


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation on the headline workload (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name = args.workload
    w = workload(name, args.n_log2)
    chk, kind = cpu_reference_checker()
    cores = os.cpu_count() or 1
    chk.set_threads(cores)
    config = {"workload": w["desc"], "N": int(w["N"]), "inner_steps_per_step": w["inner"]}
    if name == "c4" and kind == "reference":
        est, detail = cpu_tree_session(w, chk, args.warmup + args.steps, 256)
        est = est[args.warmup:]
        t = sum(est)
        value = w["units"] * len(est) / t
        sample = (f"{len(est)} steps, each = the serial phases of the reference's tree step (timed once on the full N={w['N']} problem) "
                  f"+ one OpenMP tree walk over every 256th particle scaled by N/n_sample")
        extra = {"detail": detail}
    else:
        cfg, p = w["cfg"], w["p"]
        if name == "c3":
            cb = cpu_baseline(name, w, budget_s=20.0)
            line = {"impl": "reference", "metric": w["metric"], "value": cb["value"], "unit": w["unit"], "n_gpus": args.gpus, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": 1e3 * w["units"] / cb["value"], "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": cb,
                    "e2e": {"value": cb["value"], "unit": w["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            print(json.dumps(line))
            return
        resolve = 2 if name == "c5" else 0
        from rebound_b200 import ics
        mcv = ics.SHEET_OMEGA * 0.001 if name == "c5" else 0.0
        chk.steps(cfg, p, 1, resolve=resolve, minimum_collision_velocity=mcv)
        _, _, aux = chk.steps(cfg, p, 2, resolve=resolve, minimum_collision_velocity=mcv)
        per_step = max(aux["seconds"] / 2, 1e-6)
        inner_cpu = max(1, min(w["inner"], int(2.0 / per_step)))
        for _ in range(args.warmup):
            chk.steps(cfg, p, inner_cpu, resolve=resolve, minimum_collision_velocity=mcv)
        t = 0.0
        for _ in range(args.steps):
            _, _, aux = chk.steps(cfg, p, inner_cpu, resolve=resolve, minimum_collision_velocity=mcv)
            t += aux["seconds"]
        value = w["units"] * inner_cpu * args.steps / t
        t = t * w["inner"] / inner_cpu
        sample = f"{args.steps} x reb_simulation_steps(r,{inner_cpu}) on the full workload"
        extra = {}
        est = [0] * args.steps
    line = {
        "impl": "reference", "metric": w["metric"], "value": value, "unit": w["unit"],
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / max(1, len(est)),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": w["unit"], "cores": chk.threads(), "kind": kind, "sample": sample, **extra},
        "e2e": {"value": value, "unit": w["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# GPU measurement helpers
# ------------------------------------------------------------------------------------------------------------
class Ctx:
    """One rank's measuring context: torch stream, engine, process group."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.numa_cpus = bind_to_gpu_numa_node(self.local_rank) if self.world > 1 else None
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        # An explicit (non-default) torch stream: the engine launches on it and torch.cuda.Event times it.
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        from rebound_b200.simulation import Engine

        self.eng = Engine(self.local_rank, self.stream.cuda_stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        self.sharded = False

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = self.torch.tensor([ms], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def attach(self):
        """Shards the engine over the ranks (NCCL inside the engine); once per process."""
        if self.world > 1 and not self.sharded:
            from rebound_b200 import distributed as D

            D.attach(self.eng, self.dev)
            self.sharded = True

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def pinned(self, n_particles):
        from rebound_b200 import abi

        host = self.torch.empty(max(1, n_particles) * abi.PARTICLE_DTYPE.itemsize, dtype=self.torch.uint8, pin_memory=True)
        return host, host.numpy().view(abi.PARTICLE_DTYPE)

    def close(self):
        self.eng.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def time_resident(ctx, cfg, inner, steps, warmup, flush_l2):
    """K timed calls of rebcu_steps(inner) on the resident state; returns (ms total, max over ranks; launches)."""
    eng = ctx.eng
    c = cfg.copy()
    for _ in range(warmup):
        eng.steps(c, inner)
    ctx.barrier()
    l0 = eng.launch_count
    if flush_l2:
        ev = [(ctx.event(), ctx.event()) for _ in range(steps)]
        for a, b in ev:
            ctx.flush.fill_(1)              # L2 flush between timed iterations (256 MiB > 126 MB L2), untimed
            a.record(ctx.stream)
            eng.steps(c, inner)
            b.record(ctx.stream)
        ctx.barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
    else:
        a, b = ctx.event(), ctx.event()
        a.record(ctx.stream)
        for _ in range(steps):
            eng.steps(c, inner)
        b.record(ctx.stream)
        ctx.barrier()
        ms = a.elapsed_time(b)
    return ctx.max_over_ranks(ms), int(eng.launch_count - l0), c


def time_e2e(ctx, cfg, p, inner, steps, warmup=1):
    """The same steps through the host-buffer calls: every step moves its inputs H2D from pinned memory and its result D2H."""
    from rebound_b200 import distributed as D

    eng = ctx.eng
    n = len(p)
    c = cfg.copy()
    if ctx.world == 1:
        keep, hp = ctx.pinned(n)
        hp[:] = p
        call = lambda: eng.steps_host(c, hp, inner)                 # noqa: E731
        h2d = d2h = n * 112
    else:
        b, e = D.shard_range(n, ctx.rank, ctx.world)
        keep, hp = ctx.pinned(e - b)
        hp[:] = p[b:e]

        def call():
            eng.upload_shard(hp, n)
            eng.steps(c, inner)
            eng.download_shard(hp)
        h2d = d2h = (e - b) * 112
    for _ in range(warmup):
        call()
    ctx.barrier()
    a, b_ = ctx.event(), ctx.event()
    a.record(ctx.stream)
    for _ in range(steps):
        call()
    b_.record(ctx.stream)
    ctx.barrier()
    ms = ctx.max_over_ranks(a.elapsed_time(b_))
    return ms, int(h2d), int(d2h)


def kernel_classes(ctx, cfg, inner):
    """Device time per kernel class of one call (CUDA events on the engine's stream around every launch group)."""
    eng = ctx.eng
    c = cfg.copy()
    eng.timing_enable(True)
    eng.timing_reset()
    eng.steps(c, inner)
    tim = eng.timing_read()
    eng.timing_enable(False)
    return {k: v for k, v in tim.items() if v["launches"] or v["ms"] > 0}


def load_peaks():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    return hbm, src


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernels, from the committed ncu captures."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return {}


def fp64_roofline(kernel, flop_alg, k_ms, fp64_peak, dp_instr, sm_mhz, share, traffic, note):
    """FP64-bound kernel: achieved = algorithmic flop (BASELINE.md's count per interaction) / kernel time against the DFMA
    peak measured on this box in this run; pipe_frac = FP64-pipe instructions actually issued / the measured DFMA issue rate."""
    ach = flop_alg / (k_ms * 1e-3) / 1e12
    r = {"bound": "fp64", "kernel": kernel, "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
         "peak_source": "DFMA microbenchmark run inside this bench (rebcu_measure_fp64_peak; MEASURED_PEAKS.json has no FP64 entry); "
                        f"nominal 148 SM x 64 DFMA/clk x 2 x {sm_mhz:.0f} MHz = {148 * 64 * 2 * sm_mhz * 1e6 / 1e12:.1f}",
         "traffic": traffic, "kernel_ms": k_ms, "kernel_share_of_step": share,
         "algorithmic_flop_per_launch": flop_alg, "note": note}
    if dp_instr:
        r["pipe_frac"] = dp_instr / (k_ms * 1e-3) / (fp64_peak * 1e12 / 2.0)
    return r


def measure_config(ctx, name, args, steps, warmup, headline=False):
    """Measures one BASELINE.json configuration; returns its block (rank 0 assembles; all ranks must call)."""
    from rebound_b200 import abi, ics

    w = workload(name, args.n_log2 if headline or name == args.workload else 0)
    eng, p, inner = ctx.eng, w["p"], w["inner"]
    n = w["N"]
    kind = w["kind"]
    sharded = ctx.world > 1
    big = n * 48 > (126 << 20)                     # resident x,v larger than L2
    hbm_peak, hbm_src = load_peaks()
    # the captures are single-GPU launches of the default problem sizes: null on a sharded run or with --n-log2
    traffic = {} if (sharded or args.n_log2) else load_traffic()
    out = {"metric": w["metric"], "unit": w["unit"], "config": {"workload": w["desc"], "N": int(n), "inner_steps_per_step": inner}}
    min_v = ics.SHEET_OMEGA * 0.001
    modes = [("fast", abi.MODE_FAST), ("strict", abi.MODE_STRICT)]
    if args.mode == "strict":
        modes = [("strict", abi.MODE_STRICT)]
    results = {}
    clocks = None
    for mname, mode in modes:
        cfg = w["cfg"].copy()
        cfg.mode = mode
        eng.upload(np.ascontiguousarray(p))
        if name == "c5":
            eng.set_device_resolve(True, restitution=BRIDGES, minimum_collision_velocity=min_v, rand_seed=42)
        first = mname == modes[0][0]
        k = steps if first else max(2, steps // 3)
        sampler = None
        if first:
            sampler = ClockSampler(ctx.local_rank)
        # warm-up happens inside time_resident; the sampler covers the timed region plus that warm-up's tail
        if sampler:
            sampler.start()
        cs0 = eng.comm_stats() if sharded else None
        ms, launches, c_end = time_resident(ctx, cfg, inner, k, warmup if first else 1, flush_l2=not big)
        if sampler:
            clocks = sampler.stop()
        value = w["units"] * inner * k / (ms * 1e-3)
        results[mname] = {"value": value, "ms_per_step": ms / k, "steps": k, "gpu_launches": launches}
        tim = kernel_classes(ctx, c_end, inner)
        results[mname]["kernel_ms_per_step"] = {kk: round(v["ms"], 4) for kk, v in tim.items()}
        results[mname]["_tim"] = tim
        if kind in ("tree", "sheet"):
            eng.update_acceleration(c_end)             # a tree (and group-walk counters) consistent with the current positions
            results[mname]["walk_stats"] = eng.tree_walk_stats(c_end)
        if sharded:
            cs = eng.comm_stats()
            n_calls = k + (warmup if first else 1)
            results[mname]["_comm"] = {"transport": cs["transport"],
                                       "bytes_received_per_step": (cs["bytes_received"] - cs0["bytes_received"]) // max(1, n_calls),
                                       "collectives_per_step": (cs["exchanges"] - cs0["exchanges"]) / max(1, n_calls)}
    head = modes[0][0]
    R = results[head]
    out.update({"value": R["value"], "ms_per_step": R["ms_per_step"], "mode": head + (" (FMA + rsqrt; group walk for the tree)" if head == "fast" else " (bit-identical to the reference)"),
                "gpu_launches": R["gpu_launches"], "kernel_ms_per_step": R["kernel_ms_per_step"],
                "l2": "inputs larger than L2" if big else "flushed between timed steps (256 MiB fill), untimed"})
    if "strict" in results and head != "strict":
        S = results["strict"]
        out["strict"] = {"value": S["value"], "ms_per_step": S["ms_per_step"], "mode": "strict (bit-identical to the reference)",
                         "kernel_ms_per_step": S["kernel_ms_per_step"]}

    # ---- end to end: host buffers through the C ABI, every step H2D + D2H ----
    cfg = w["cfg"].copy()
    cfg.mode = modes[0][1]
    e2e_steps = steps if headline else max(2, steps // 3)
    ms_e2e, h2d, d2h = time_e2e(ctx, cfg, p, inner, e2e_steps)
    out["e2e"] = {"value": w["units"] * inner * e2e_steps / (ms_e2e * 1e-3), "unit": w["unit"], "ms_per_step": ms_e2e / e2e_steps,
                  "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                  "call": "rebcu_steps_host" if ctx.world == 1 else "rebcu_upload_shard + rebcu_steps + rebcu_download_shard (own block per rank)"}
    if name == "c5":
        eng.set_device_resolve(False)

    # ---- roofline of the dominant kernel ----
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp64_peak = eng.measure_fp64_peak()
    for mname in results:
        tim = results[mname].pop("_tim")
        total = max(1e-9, sum(v["ms"] for v in tim.values()))
        tag = "fast" if mname == "fast" else "strict"
        if kind in ("basic", "compensated"):
            cls = "direct"
            k_ms = tim[cls]["ms"] / max(1, tim[cls]["launches"])
            per_launch_units = w["units"] * inner / max(1, tim[cls]["launches"]) / ctx.world      # this rank's target block
            kern = ("tp_multistep_kernel" if name == "c2" else ("direct_fast_kernel" if tag == "fast" else "direct_strict_kernel"))
            roof = fp64_roofline(kern, FLOP_PER_INTERACTION[kind] * per_launch_units, k_ms, fp64_peak,
                                 DP_PER_PAIR[(kind, tag)] * per_launch_units, sm_mhz, tim[cls]["ms"] / total, traffic.get(f"{name}_{tag}"),
                                 f"{FLOP_PER_INTERACTION[kind]:.0f} flop per ordered interaction (BASELINE.md); {DP_PER_PAIR[(kind, tag)]:.0f} FP64-pipe instructions issued per pair in this mode")
        else:
            cls = "treewalk"
            st = results[mname]["walk_stats"]
            k_ms = tim[cls]["ms"]                      # walk_pack + walk of one force evaluation
            grouped = tag == "fast" and st["groups"] > 0        # FAST without ghost boxes: the group walk
            evaluated = st["group_entries"] * 32 if grouped else st["interactions"]
            # FP64-pipe instructions: 16 (FAST, fast_math.cuh) / 36 (STRICT) per pair term; the per-particle walks add 8 per
            # visited cell (the group walk's traversal and exact-criterion tests are not counted: a lower bound)
            dp = 16.0 * evaluated if grouped else (16.0 if tag == "fast" else 36.0) * evaluated + 8.0 * st["visits"]
            kern = "walk_group_kernel" if grouped else ("walk_rec_kernel<FAST>" if tag == "fast" else "walk_rec_kernel")
            roof = fp64_roofline(kern, 20.0 * st["interactions"], k_ms, fp64_peak, dp, sm_mhz, tim[cls]["ms"] / total,
                                 traffic.get(f"{name}_{tag}"),
                                 "algorithmic work = the interactions of the reference's per-particle opening criterion on this tree "
                                 f"({st['interactions'] / max(1, n / ctx.world):.0f} per particle) x 20 flop; the group walk evaluates "
                                 f"{st['group_entries'] * 32 / max(1, st['interactions']):.2f}x as many pair terms (group criterion = every particle of the group accepts the cell)" if grouped else
                                 "algorithmic work = accepted cells + leaves of the per-particle walk x 20 flop; "
                                 f"{16 if tag == 'fast' else 36} FP64-pipe instructions per interaction + 8 per visited cell")
            roof["interactions_per_particle"] = st["interactions"] / max(1, n / ctx.world)
            if grouped:
                roof["evaluated_pair_terms_per_particle"] = st["group_entries"] * 32 / max(1, n / ctx.world)
            # the build is the HBM-bound part of the step (keys 36 B + sort 4 x 24 B + cells 1.5 x 64 + 32 B per particle)
            if "treebuild" in tim:
                b_ms = tim["treebuild"]["ms"]
                alg = 260.0 * n * (2 if kind == "sheet" else 1)
                roof["build_hbm"] = {"bound": "hbm", "kernels": "key, radix sort, tie/lcp, emit, adopt, moment", "achieved": alg / (b_ms * 1e-3) / 1e9,
                                     "peak": hbm_peak, "unit": "GB/s", "frac": alg / (b_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                                     "algorithmic_bytes": alg, "ms": b_ms, "traffic": traffic.get(f"{name}_build"),
                                     "note": "260 B per particle per build (SURVEY.md 8d) x ALL particles, also when the build is sharded over the ranks"}
        if mname == head:
            out["roofline"] = roof
        else:
            out["strict"]["roofline"] = roof
        results[mname].pop("walk_stats", None)
    if sharded:
        cs = results[head].get("_comm")
        tim_x = R["kernel_ms_per_step"].get("exchange", 0.0)
        out["exchange"] = {"transport": cs["transport"], "bytes_received_per_rank_per_step": cs["bytes_received_per_step"],
                           "collectives_per_step": cs["collectives_per_step"], "exchange_ms_per_step_rank0": tim_x,
                           "collectives": "ncclAllGather of x,y,z in place on the engine's stream between drift and force"
                                          + ("; tree: bucket table, traversal records (40 B per cell) and sorted permutation of the per-rank builds" if kind == "tree" else "")}
    out["clocks"] = clocks
    return out, w


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-log2", type=int, default=0, help="log2 of N for the selected workload (c3: default 20, c4: 24, c5: 20)")
    ap.add_argument("--mode", default="fast", choices=["strict", "fast"], help="arithmetic mode of the headline value (the other one is reported beside it)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard-build", type=int, default=2, choices=[0, 1, 2],
                    help="tree builds of a sharded run: 0 every rank builds the whole tree, 1 per-rank subtree builds, 2 automatic (default)")
    ap.add_argument("--no-configs", action="store_true", help="skip the blocks of the other configurations")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    ctx = Ctx()
    rank, world = ctx.rank, ctx.world
    name = args.workload
    if world > 1 and name in ("c1", "c2", "c5"):
        raise SystemExit(f"{name} is a single-GPU configuration; the sharded paths are c3 and c4")
    # ---- headline ----
    ctx.attach()                                   # N > 1: NCCL communicator inside the engine, one rank per GPU
    ctx.eng.set_sharded_build(args.shard_build)
    head, w = measure_config(ctx, name, args, args.steps, args.warmup, headline=True)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(name, w)
    # ---- the other configurations ----
    blocks = {}
    if not args.no_configs:
        others = [c for c in (("c1", "c2", "c3", "c5") if world == 1 else ("c3",)) if c != name]
        if name != "c4" and world == 1:
            others.append("c4")
        for c in others:
            saved = args.n_log2
            args.n_log2 = 0
            if c == "c4":
                args.n_log2 = 22                    # as a side block the tree config runs at 2^22
            blk, wc = measure_config(ctx, c, args, max(2, min(args.steps, 3)), 1)
            args.n_log2 = saved
            if rank == 0 and world == 1 and not args.no_cpu_baseline:
                blk["cpu_baseline"] = cpu_baseline(c, wc)
            blocks[c] = blk
    if rank == 0:
        line = {
            "metric": head["metric"], "value": head["value"], "unit": head["unit"], "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if name in ("c3", "c4") else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": head["config"],
            "detail": {"mode": head["mode"], "l2": head["l2"],
                       "sharding": ("one problem, contiguous target blocks per rank, NCCL all-gather of x,y,z inside the engine between drift and force; "
                                    + ("every rank builds the whole tree" if args.shard_build == 0 else "every rank sorts and builds the subtrees of its own key range, "
                                       "traversal records all-gathered, top of the tree filled in by everyone")
                                    + "; every rank walks its own block in key order") if world > 1 else "single GPU, no exchange",
                       "host_affinity": (f"each rank bound to the {ctx.numa_cpus} cores local to its GPU" if ctx.numa_cpus else "unbound")},
            "clocks": head["clocks"],
            "e2e": head["e2e"],
            "gpu_launches": head["gpu_launches"],
            "roofline": head["roofline"],
            "cpu_baseline": cpu,
            "kernel_ms_per_step": head["kernel_ms_per_step"],
        }
        if "strict" in head:
            line["strict"] = head["strict"]
        if "exchange" in head:
            line["exchange"] = head["exchange"]
        for blk in blocks.values():
            blk.pop("clocks", None)
        line["configs"] = blocks
        print(json.dumps(line))
    ctx.close()


if __name__ == "__main__":
    main()
