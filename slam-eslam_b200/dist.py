"""Multi-GPU plumbing: one process per GPU.  torch.distributed carries the rendezvous (the 128-byte NCCL unique id,
barriers, max-over-ranks of timings); the filter's own exchange steps run inside libeslam_gpu.so on NCCL."""
import ctypes as C

import torch
import torch.distributed as td

from . import abi


class Comm:
    def __init__(self, world, rank, local_rank, backend):
        self.world, self.rank, self.local_rank, self.backend = world, rank, local_rank, backend
        self.device = torch.device("cuda", local_rank) if backend == "nccl" else torch.device("cpu")

    def barrier(self):
        td.barrier()

    def max_float(self, v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    def broadcast_bytes(self, data, n):
        t = torch.zeros(n, dtype=torch.uint8, device=self.device)
        if self.rank == 0:
            t.copy_(torch.tensor(list(data), dtype=torch.uint8))
        td.broadcast(t, src=0)
        return bytes(t.cpu().tolist())

    def shard(self, n_local):
        return abi.Shard(rank=self.rank, world=self.world, global_count=n_local * self.world)

    def connect(self, estimator):
        """Create the filter's NCCL communicator: rank 0 makes the unique id, everyone joins."""
        lib = abi.load_library()
        buf = (C.c_uint8 * 128)()
        if self.rank == 0:
            assert lib.eslam_gpu_nccl_unique_id(buf) == abi.OK
        uid = self.broadcast_bytes(bytes(buf), 128)
        estimator.commInit(uid)


def init(world, rank, local_rank, backend="nccl"):
    if not td.is_initialized():
        td.init_process_group(backend=backend, world_size=world, rank=rank)
    return Comm(world, rank, local_rank, backend)
