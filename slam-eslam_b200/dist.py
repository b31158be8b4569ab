"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the exchange steps."""
import torch
import torch.distributed as td


class Comm:
    def __init__(self, world, rank, local_rank, backend):
        self.world, self.rank, self.local_rank = world, rank, local_rank
        self.device = torch.device("cuda", local_rank) if backend == "nccl" else torch.device("cpu")
        self.hooks = None

    def barrier(self):
        td.barrier()

    def max_float(self, v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())


def init(world, rank, local_rank, backend="nccl"):
    if not td.is_initialized():
        td.init_process_group(backend=backend, world_size=world, rank=rank)
    return Comm(world, rank, local_rank, backend)
