"""Multi-GPU plumbing: one process per GPU, torch.distributed for the exchange.

Direct summation and the tree walk shard the *targets*: rank r of W owns the contiguous index block
[N*r/W, N*(r+1)/W) (`rebcu_set_shard`) and computes forces, kicks and drifts only for it.  Between the
drift and the force evaluation every rank needs the other blocks' new positions: that is the one real
exchange step of the path -- an all-gather of x, y, z (24 B per particle), the role the reference's MPI
build gives to reb_communication_mpi_distribute_* (src/communication_mpi.c:109-181, 354-438).
The engine calls back into `BlockExchange.__call__` at exactly that point (rebcu_set_exchange_callback).

Everything here is plumbing on torch tensors; it runs on CPU tensors with the gloo backend as well, which
is how tests/test_distributed_cpu.py covers it without a GPU.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """The block rebcu_shard_range reports (csrc/context.cu: engine_shard)."""
    return n * rank // world, n * (rank + 1) // world


class _DevicePtr:
    """Zero-copy torch view of a device array owned by the engine (CUDA array interface v2)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def device_view(ptr, n, device):
    return torch.as_tensor(_DevicePtr(ptr, n), device=device)


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


def attach(engine, device, group=None):
    """Shards `engine` over the process group and installs the position exchange.  Call after upload."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    engine.set_shard(rank, world)
    state = {"calls": 0}

    def exchange():
        # The engine must launch on torch's current stream (pass torch.cuda.current_stream().cuda_stream
        # of an explicit stream to Engine) so that NCCL orders behind the drift kernel.  Views are rebuilt
        # per call: the engine may swap its SoA block (open-boundary compaction) or change N.
        n = engine.N
        fields = [device_view(engine.device_field(k), n, device) for k in range(3)]      # x, y, z
        BlockExchange(fields, group)()
        state["calls"] += 1

    engine.set_exchange_callback(exchange)
    return state


def gather_owned(engine, device, group=None):
    """After stepping: every rank's owned block of x..vz (and ax..az) gathered everywhere, so that rank 0
    can download a complete state.  Returns nothing; the engine's arrays are updated in place."""
    n = engine.N
    fields = [device_view(engine.device_field(k), n, device) for k in range(9)]
    BlockExchange(fields, group)()
