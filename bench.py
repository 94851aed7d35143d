#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract: see DESIGN.md "Measurement").

Default workload = BASELINE.json configs[1] ("c2"): N_active=10 massive bodies + 2^20 test
particles, REB_GRAVITY_BASIC, testparticle_type=0, leapfrog.  One bench "step" is one
reb_simulation_steps(r, 100)-sized batch (100 leapfrog steps, the count configs[0] quotes).
Metric: pairwise interactions/s = [N*N_active - N_active] * force evaluations / time  (BASELINE.md).

  value  device-resident: particles already in HBM, CUDA events around K batches.
  e2e    the same batches through the host-buffer C-ABI call rebcu_steps_host (what the shim's
         reb_simulation_steps would call): pinned host AoS -> H2D -> 100 steps -> D2H, per batch.
  --impl reference : the reference's own CPU path (oracle/_ref OpenMP build when present, else the
         oracle port) on the same workload, all host threads.
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

INNER_STEPS = {"c2": 100, "c1": 100, "c1fast": 100, "c2fast": 100, "c3": 1, "c4": 1}


def workload(name, rank=0):
    """Returns (particles, config, interactions per force evaluation, description)."""
    from rebound_b200 import abi, ics

    if name.startswith("c2"):
        n_test = 1 << 20
        p = ics.planetesimal_disk(n_test, seed=42 + rank)
        cfg = ics.planetesimal_config()
        if name.endswith("fast"):
            cfg.mode = abi.MODE_FAST
        n = len(p)
        inter = n * 10 - 10
        desc = "C2 planetesimal disk: N_active=10 + 2^20 test particles, REB_GRAVITY_BASIC, testparticle_type=0, leapfrog"
    elif name.startswith("c1"):
        n = 16384
        p = ics.plummer(n, seed=42 + rank)
        cfg = ics.plummer_config(n)
        if name.endswith("fast"):
            cfg.mode = abi.MODE_FAST
        inter = n * n - n
        desc = "C1 Plummer sphere N=16384, REB_GRAVITY_BASIC, leapfrog"
    else:
        raise SystemExit(f"unknown workload {name}")
    return p, cfg, inter, desc


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
    2 ms from a thread (the timed region of the default run is ~30 ms); falls back to an nvidia-smi loop."""

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
    copy threads are local to its PCIe root; without it 8 ranks push their host traffic through one socket."""
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


def cpu_reference_checker():
    import checkers

    ref = checkers.reference(openmp=True)
    if ref is not None:
        return ref, "reference"
    return checkers.oracle(), "port"


def run_reference_arm(args):
    """The reference's own CPU implementation on the same workload (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p, cfg, inter, desc = workload(args.workload)
    chk, kind = cpu_reference_checker()
    cores = os.cpu_count() or 1
    chk.set_threads(cores)
    inner = args.inner or INNER_STEPS[args.workload]
    # bound the CPU work per bench step to ~2 s: reduce the inner step count, never the problem size
    chk.steps(cfg, p, 1)                      # thread pool start-up, page faults
    _, _, aux = chk.steps(cfg, p, 2)
    per_step = max(aux["seconds"] / 2, 1e-6)
    inner_cpu = max(1, min(inner, int(2.0 / per_step)))
    for _ in range(args.warmup):
        chk.steps(cfg, p, inner_cpu)
    t = 0.0
    for _ in range(args.steps):
        _, _, aux = chk.steps(cfg, p, inner_cpu)
        t += aux["seconds"]
    value = inter * inner_cpu * args.steps / t
    line = {
        "impl": "reference", "metric": "pairwise interactions/s (direct)", "value": value, "unit": "interactions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "inner_steps_per_step": inner_cpu, "N": int(len(p))},
        "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": chk.threads(), "kind": kind,
                         "sample": f"{args.steps} x reb_simulation_steps(r,{inner_cpu}) on the full workload"},
        "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_sharded(args):
    """BASELINE.json configs[2] / configs[3] (`--workload c3|c4`): ONE problem sharded over the ranks in
    contiguous target blocks (strong scaling), positions all-gathered over NCCL between drift and force
    (rebound_b200/distributed.py).  c3: Plummer sphere, REB_GRAVITY_COMPENSATED direct summation;
    c4: self-gravitating disc, REB_GRAVITY_TREE theta^2=0.25, open boundary.  `--n-log2` scales N."""
    import torch
    import torch.distributed as dist

    from rebound_b200 import abi, distributed as D, ics
    from rebound_b200.simulation import Engine

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    if args.workload == "c3":
        n = 1 << (args.n_log2 or 22)
        p = ics.plummer(n, seed=42)
        cfg = ics.plummer_config(n, gravity=abi.GRAVITY_COMPENSATED)
        units_per_step = float(n) * n - n
        metric, unit = "pairwise interactions/s (direct)", "interactions/s"
        desc = f"C3 Plummer sphere N=2^{int(np.log2(n))}, REB_GRAVITY_COMPENSATED direct summation, leapfrog"
    else:
        n = 1 << (args.n_log2 or 24)
        p = ics.selfgravity_disc(n - 1, seed=42)
        cfg = ics.selfgravity_disc_config()
        units_per_step = float(n)
        metric, unit = "particle-steps/s (tree)", "particle-steps/s"
        desc = f"C4 self-gravitating disc N=2^{int(np.log2(n))}, REB_GRAVITY_TREE opening_angle2=0.25, leapfrog, open boundary"
    if args.mode == "fast":
        cfg.mode = abi.MODE_FAST
    inner = args.inner or 1
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng = Engine(local_rank, stream.cuda_stream)
    eng.upload(np.ascontiguousarray(p))
    if world > 1:
        D.attach(eng, dev)
    c = cfg.copy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        eng.steps(c, inner)
    barrier()
    launches0 = eng.launch_count
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(args.steps):
        eng.steps(c, inner)
    t1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = units_per_step * inner * args.steps / (ms_total * 1e-3)
    eng.timing_enable(True)
    eng.timing_reset()
    eng.steps(c, inner)
    tim = eng.timing_read()
    eng.timing_enable(False)
    if rank == 0:
        print(json.dumps({
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "inner_steps_per_step": inner, "N_total": int(n),
                       "mode": "strict (bit-identical to the reference)" if cfg.mode == 0 else "fast",
                       "sharding": "contiguous target blocks per rank, NCCL all-gather of x,y,z between drift and force",
                       "l2": "inputs larger than L2" if n >= (1 << 22) else "L2 not flushed (state fits L2)"},
            "clocks": clocks, "gpu_launches": int(eng.launch_count - launches0),
            "kernel_ms_per_step_rank0": {k: v["ms"] for k, v in tim.items() if v["launches"]},
            "e2e": None, "roofline": None, "cpu_baseline": None,
        }))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-log2", type=int, default=0, help="log2 of N for the sharded workloads c3/c4")
    ap.add_argument("--mode", default="strict", choices=["strict", "fast"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(INNER_STEPS))
    ap.add_argument("--inner", type=int, default=0, help="leapfrog steps per bench step (default 100)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the informational measurements of the other configurations")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.workload in ("c3", "c4"):
        run_sharded(args)
        return

    import torch
    import torch.distributed as dist

    from rebound_b200 import abi
    from rebound_b200.simulation import Engine

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    numa_cpus = None
    if world > 1:
        numa_cpus = bind_to_gpu_numa_node(local_rank)      # before anything allocates host memory or starts threads
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    p, cfg, inter, desc = workload(args.workload, rank)
    inner = args.inner or INNER_STEPS[args.workload]
    n = len(p)
    # An explicit (non-default) torch stream: the engine launches on it and torch.cuda.Event times it.
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng = Engine(local_rank, stream.cuda_stream)

    # pinned host AoS (the role of r->particles after rebcu_host_register)
    host = torch.empty(n * abi.PARTICLE_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
    hp = host.numpy().view(abi.PARTICLE_DTYPE)
    hp[:] = p
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm ----------------
    eng.upload(hp)
    c = cfg.copy()
    for _ in range(args.warmup):
        eng.steps(c, inner)
    barrier()
    launches0 = eng.launch_count
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.fill_(1)                      # L2 flush between timed iterations (256 MiB > 126 MB L2), untimed
        a.record(stream)
        eng.steps(c, inner)
        b.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = eng.launch_count - launches0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t_dev = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_total = float(t_dev.item())
    value = world * inter * inner * args.steps / (ms_total * 1e-3)

    # ---------------- end-to-end arm: host buffers through rebcu_steps_host ----------------
    c = cfg.copy()
    hp[:] = p
    for _ in range(2):
        eng.steps_host(c, hp, inner)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(args.steps):
        eng.steps_host(c, hp, inner)       # H2D of the AoS + `inner` steps + D2H of the AoS, synchronous
    t1.record(stream)
    barrier()
    e2e_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * inter * inner * args.steps / (float(e2e_ms.item()) * 1e-3)

    # ---------------- roofline of the dominant kernel (per-launch, CUDA events on this stream) ---------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    eng.upload(hp)
    c = cfg.copy()
    eng.timing_enable(True)
    eng.timing_reset()
    eng.steps(c, inner)
    tim = eng.timing_read()
    eng.timing_enable(False)
    dom = max(tim, key=lambda k: tim[k]["ms"])
    k_ms = tim[dom]["ms"] / max(1, tim[dom]["launches"])
    steps_in_launch = 1
    if args.workload.startswith("c2"):
        # tp_multistep_kernel: ONE launch advances every test particle through all `inner` steps with x,v in
        # registers: 48 B read + 72 B written per particle per launch (x,v in; x,v,a out), DESIGN.md section 3.
        steps_in_launch = inner
        # 96 B per particle-step (x,v in; x,v out) x the particle-steps one launch processes.
        alg_bytes = 96.0 * n * inner
        roof = {"bound": "hbm", "kernel": "tp_multistep_kernel", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "peak_source": peak_src,
                "traffic": 67.76e6,          # dram__bytes_read + write per launch, ncu --set full (profiles/r01_ms_ncu.txt)
                "algorithmic_bytes_per_launch": alg_bytes, "resident_state_bytes_per_launch": 120.0 * n,
                "limiting_resource": "fp64 pipe",
                "note": f"algorithmic bytes = 96 B per particle-step (SURVEY 8d) x N x {inner} steps per launch; the kernel keeps "
                        "x,v in registers across all steps of the launch, so the DRAM traffic it really causes is 120 B per particle "
                        "per LAUNCH (`traffic`, far below the algorithmic bytes) and what bounds it is the FP64 pipe (strict IEEE "
                        "sqrt+divide: 36 DP instructions per interaction), see fp64.pipe_frac / fp64.frac_of_measured_peak"}
    else:
        flops = 20.0 * inter                 # 20 flop per interaction
        roof = {"bound": "hbm", "kernel": "direct_strict_kernel", "achieved": 32.0 * n / (k_ms * 1e-3) / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "peak_source": peak_src, "traffic": None,
                "note": "FP64-pipe bound kernel; HBM traffic negligible, see fp64", "alg_tflops": flops / (k_ms * 1e-3) / 1e12}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["kernel_ms"] = k_ms
    roof["kernel_share_of_step"] = tim[dom]["ms"] / max(1e-9, sum(v["ms"] for v in tim.values()))
    # FP64 pipe view: interactions/s of the kernel alone x 20 flop against 148 SM x 64 DFMA/clk x 2 x clock
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    fp64_peak = 148 * 64 * 2 * sm_mhz * 1e6 / 1e12
    fp64_measured = eng.measure_fp64_peak()          # DFMA microbenchmark on this box, TFLOP/s
    k_inter_per_s = inter * steps_in_launch / (k_ms * 1e-3)
    dp_per_inter = 36.0 if cfg.mode == 0 else 22.0      # FP64-pipe instructions per interaction (SASS count)
    roof["fp64"] = {"achieved_tflops": 20.0 * k_inter_per_s / 1e12, "peak_tflops": fp64_peak,
                    "peak_source": "148 SM x 64 DFMA/clk x 2 flop x sampled SM clock (nominal)",
                    "frac": 20.0 * k_inter_per_s / 1e12 / fp64_peak,
                    "measured_peak_tflops": fp64_measured,
                    "frac_of_measured_peak": 20.0 * k_inter_per_s / 1e12 / fp64_measured,
                    "pipe_frac_of_measured_peak": k_inter_per_s * dp_per_inter * 2.0 / 1e12 / fp64_measured,
                    "pipe_frac": k_inter_per_s * dp_per_inter / (148 * 64 * sm_mhz * 1e6),
                    "pipe_frac_note": "issued FP64-pipe instructions per lane-slot: interactions/s x DP instructions per "
                                      "interaction / (148 SM x 64 lanes x clock); 20 flop/interaction is BASELINE.md's algorithmic count"}

    # ---------------- other BASELINE.json configurations, device-resident, informational ----------------
    others = None
    if rank == 0 and world == 1 and args.workload == "c2" and not args.no_extra:
        from rebound_b200 import ics

        def timed(fn, reps):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) * 1e-3 / reps

        others = {}
        n1 = 16384
        p1 = ics.plummer(n1, seed=42)
        for tag, mode in (("c1_plummer16384_basic_strict", abi.MODE_STRICT), ("c1_plummer16384_basic_fast", abi.MODE_FAST)):
            c1 = ics.plummer_config(n1, mode=mode)
            eng.upload(np.ascontiguousarray(p1))
            s1 = timed(lambda: eng.steps(c1, 10), 3) / 10
            others[tag] = {"interactions_per_s": (n1 * n1 - n1) / s1, "ms_per_step": s1 * 1e3}
        n4 = 1 << 20
        p4 = ics.selfgravity_disc(n4 - 1, seed=42)
        c4 = ics.selfgravity_disc_config()
        eng.upload(np.ascontiguousarray(p4))
        s4 = timed(lambda: eng.steps(c4, 2), 2) / 2
        others["c4_disc_2pow20_tree_strict"] = {"particle_steps_per_s": n4 / s4, "ms_per_step": s4 * 1e3}
        eng.upload(hp)

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        chk, kind = cpu_reference_checker()
        chk.set_threads(os.cpu_count() or 1)
        chk.steps(cfg, p, 1)                  # thread pool start-up, page faults
        _, _, aux = chk.steps(cfg, p, 2)
        # about 10 s of CPU work on the full-size workload (more steps, never fewer particles)
        n_cpu = max(1, min(50 * inner, int(10.0 / max(aux["seconds"] / 2, 1e-6))))
        _, _, aux = chk.steps(cfg, p, n_cpu)
        cpu = {"value": inter * n_cpu / aux["seconds"], "unit": "interactions/s", "cores": chk.threads(), "kind": kind,
               "sample": f"reb_simulation_steps(r,{n_cpu}) on the full workload ({aux['seconds']:.1f} s)"}

    if rank == 0:
        line = {
            "metric": "pairwise interactions/s (direct)", "value": value, "unit": "interactions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "inner_steps_per_step": inner, "N_per_gpu": int(n),
                       "mode": "strict (bit-identical to the reference)" if cfg.mode == 0 else "fast",
                       "l2": "flushed between timed steps (256 MiB fill), untimed",
                       "sharding": "test particles sharded per rank, massive bodies replicated, no collective",
                       "host_affinity": (f"each rank bound to the {numa_cpus} cores local to its GPU" if numa_cpus else "unbound")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "interactions/s", "h2d_bytes_per_step": int(n * 112), "d2h_bytes_per_step": int(n * 112)},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "other_configs": others,
        }
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
