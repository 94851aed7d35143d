"""Multi-GPU plumbing.

Direct summation and the tree walk shard the *targets*: rank r of W owns the contiguous index block
[N*r/W, N*(r+1)/W) (`rebcu_set_shard`) and computes forces, kicks and drifts only for it.  Between the
drift and the force evaluation every rank needs the other blocks' new positions: that is the one real
exchange step of the path -- an all-gather of x, y, z (24 B per particle), the role the reference's MPI
build gives to reb_communication_mpi_distribute_* (src/communication_mpi.c:109-181, 354-438).
Two rarer events widen the exchange: a collision search also needs the other blocks' velocities, and an
open-boundary removal needs every field, because the compaction shifts particles across block borders.

The exchange itself runs INSIDE the engine (csrc/comm.cu): ncclAllGather on the engine's stream, ordered
between the drift and the force kernels without any host round trip.  What is left here is plumbing:
  * `attach`       one process per GPU: ships NCCL's unique id from rank 0 to the others over
                   torch.distributed and hands it to rebcu_comm_init_rank;
  * `LocalGroup`   several engines in ONE process (one thread each), NCCL via ncclCommInitAll or -- when engines
                   share a device, e.g. to test the sharded kernels on a single GPU -- the LOCAL transport;
  * `attach(..., transport="callback")`  the exchange as a Python callback on torch tensors (`BlockExchange`);
                   runs on CPU tensors with the gloo backend, which is how tests/test_distributed_cpu.py
                   covers the block arithmetic without a GPU;
  * collision lists are merged on the host (`gather_collisions`, `merge_collision_segments`).
"""
import threading

import numpy as np
import torch
import torch.distributed as dist

from . import abi


def shard_range(n, rank, world):
    """The block rebcu_shard_range reports (csrc/context.cu: engine_shard)."""
    return n * rank // world, n * (rank + 1) // world


class _DevicePtr:
    """Zero-copy torch view of a device array owned by the engine (CUDA array interface v2)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_view(ptr, n, device, typestr="<f8"):
    return torch.as_tensor(_DevicePtr(ptr, n, typestr), device=device)


def exchange_fields(request):
    """Field indices (rebcu_device_field) the exchange callback has to gather for a rebcu_exchange_request mask."""
    if request & abi.EXCHANGE_ALL:
        return list(range(abi.N_FIELDS))
    fields = [0, 1, 2]
    if request & abi.EXCHANGE_VELOCITIES:
        fields += [3, 4, 5]
    return fields


class BlockExchange:
    """All-gathers the owners' blocks of the given full-length tensors in place.

    `fields` are full-length (N) tensors, identical in layout on every rank; after the call every rank
    holds every owner's block.  Equal blocks use one all_gather_into_tensor per field; ragged blocks
    (N not divisible by W) fall back to one broadcast per owner."""

    def __init__(self, fields, group=None):
        self.fields = list(fields)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n = int(self.fields[0].numel())
        self.ranges = [shard_range(self.n, r, self.world) for r in range(self.world)]
        sizes = {e - b for b, e in self.ranges}
        self.even = len(sizes) == 1
        self.calls = 0

    def __call__(self):
        self.calls += 1
        b, e = self.ranges[self.rank]
        for f in self.fields:
            if self.even:
                # out-of-place into a scratch tensor, then copy back: in-place all-gather on overlapping
                # input/output views is not portable across backends
                mine = f[b:e].clone()
                dist.all_gather_into_tensor(f, mine, group=self.group)
            else:
                for r, (rb, re) in enumerate(self.ranges):
                    if re > rb:
                        dist.broadcast(f[rb:re], src=dist.get_global_rank(self.group, r) if self.group else r, group=self.group)


class ExchangeState:
    """What `attach` returns: counters of the exchange.  state["calls"] = exchanges so far, state["fields"] = 8-byte
    fields gathered so far (3 per position exchange, 6 with velocities, 14 for everything)."""

    def __init__(self, engine, counters=None):
        self.engine = engine
        self.counters = counters

    def _native(self):
        st = self.engine.comm_stats()
        n = self.engine.N
        b, e = self.engine.shard_range()
        other = max(1, n - (e - b))
        return {"calls": st["exchanges"], "fields": st["bytes_received"] // (8 * other), "bytes_received": st["bytes_received"],
                "transport": st["transport"]}

    def __getitem__(self, key):
        return (self.counters if self.counters is not None else self._native())[key]

    def get(self, key, default=None):
        d = self.counters if self.counters is not None else self._native()
        return d.get(key, default)


def attach(engine, device, group=None, transport="nccl"):
    """Shards `engine` over the process group (one process per GPU) and installs the exchange.  Call after upload.
    transport "nccl": the engine's own NCCL communicator (rank 0's unique id travels over torch.distributed);
    "callback": the exchange as a Python callback on torch tensors (any torch.distributed backend)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if transport == "nccl":
        uid = [engine.comm_unique_id() if rank == 0 else None]
        src = dist.get_global_rank(group, 0) if group else 0
        dist.broadcast_object_list(uid, src=src, group=group)
        engine.comm_init_rank(uid[0], rank, world)
        return ExchangeState(engine)
    engine.set_shard(rank, world)
    state = {"calls": 0, "fields": 0}

    def exchange():
        # The engine must launch on torch's current stream (pass torch.cuda.current_stream().cuda_stream
        # of an explicit stream to Engine) so that the collective orders behind the drift kernel.  Views are rebuilt
        # per call: the engine may swap its SoA block (open-boundary compaction) or change N.
        n = engine.N
        ks = exchange_fields(engine.exchange_request)       # x, y, z unless a collision search / removal asks for more
        # 64-bit integer views: a byte-exact transport also for the tag fields (name, ap, sim)
        fields = [device_view(engine.device_field(k), n, device, "<i8") for k in ks]
        BlockExchange(fields, group)()
        state["calls"] += 1
        state["fields"] += len(ks)

    engine.set_exchange_callback(exchange)
    return ExchangeState(engine, state)


def gather_owned(engine, device=None, group=None):
    """After stepping: every rank's owned block of every field gathered everywhere, so that rank 0 can download a
    complete state.  Native transport: rebcu_exchange(ALL); callback transport: the same through the callback."""
    if engine.comm_stats()["transport"] is not None:
        engine.exchange(abi.EXCHANGE_ALL)
        return
    n = engine.N
    fields = [device_view(engine.device_field(k), n, device) for k in range(9)]
    BlockExchange(fields, group)()


class LocalGroup:
    """Several engines driven by the threads of ONE process: `run(fn)` calls fn(rank, engine) on every engine from its
    own thread (the engine's calls block in barriers / collectives until all ranks have made them).  Engines on
    distinct devices use NCCL (ncclCommInitAll); engines that share a device use the LOCAL transport."""

    def __init__(self, engines, transport=abi.TRANSPORT_AUTO):
        import ctypes as C

        self.engines = list(engines)
        arr = (C.c_void_p * len(self.engines))(*[e.h for e in self.engines])
        err = self.engines[0].f["comm_init_all"](arr, len(self.engines), int(transport))
        if err != 0:
            self.engines[0]._check(err)

    def run(self, fn):
        out = [None] * len(self.engines)
        errs = []

        def work(r):
            try:
                out[r] = fn(r, self.engines[r])
            except BaseException as e:      # noqa: BLE001 -- re-raised in the caller's thread
                errs.append(e)

        ts = [threading.Thread(target=work, args=(r,)) for r in range(len(self.engines))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]
        return out

    def collisions(self):
        """The complete collision list of the last sharded search in the reference's serial order."""
        lists = [e.collisions_fetch() for e in self.engines]
        segs = [e.collisions_segments() for e in self.engines]
        return merge_collision_segments(lists, segs)


def merge_collision_segments(lists, segments):
    """The complete list in the reference's serial order from per-rank lists (rebcu_collisions_segments):
    segment by segment (ghost box major for DIRECT/LINE, a single segment for TREE/LINETREE), ranks in
    order inside a segment (projectile blocks are contiguous and ascending in rank)."""
    n_seg = max((len(s) for s in segments), default=0)
    starts = [np.concatenate([[0], np.cumsum(s)]).astype(np.int64) if len(s) else np.zeros(1, np.int64) for s in segments]
    parts = []
    for g in range(n_seg):
        for r, lst in enumerate(lists):
            if g < len(segments[r]) and segments[r][g]:
                parts.append(lst[starts[r][g]:starts[r][g + 1]])
    if not parts:
        return np.zeros(0, dtype=abi.COLLISION_DTYPE)
    return np.concatenate(parts)


def gather_collisions(engine, device, group=None):
    """After a sharded collision search (rebcu_collision_search / rebcu_steps): every rank receives the
    complete collision list in the reference's serial order.  `device` is the tensor device used for the
    transport ("cuda:k" with NCCL, "cpu" with gloo)."""
    world = dist.get_world_size(group)
    local = engine.collisions_fetch()
    seg = engine.collisions_segments()
    n_seg = torch.tensor([len(seg)], dtype=torch.int64, device=device)
    dist.all_reduce(n_seg, op=dist.ReduceOp.MAX, group=group)
    n_seg = int(n_seg.item())
    mine = torch.zeros(n_seg + 1, dtype=torch.int64)
    mine[: len(seg)] = torch.tensor(seg, dtype=torch.int64)
    mine[n_seg] = len(local)
    mine = mine.to(device)
    counts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(counts, mine, group=group)
    counts = [c.cpu().numpy() for c in counts]
    longest = max(int(c[n_seg]) for c in counts)
    if longest == 0:
        return np.zeros(0, dtype=abi.COLLISION_DTYPE)
    size = abi.COLLISION_DTYPE.itemsize
    buf = torch.zeros(longest * size, dtype=torch.uint8)
    if len(local):
        buf[: len(local) * size] = torch.from_numpy(local.view(np.uint8).reshape(-1).copy())
    buf = buf.to(device)
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    lists = [b.cpu().numpy()[: int(c[n_seg]) * size].view(abi.COLLISION_DTYPE) for b, c in zip(bufs, counts)]
    return merge_collision_segments(lists, [list(map(int, c[:n_seg])) for c in counts])
