"""TEST INFRASTRUCTURE: a communicator with the interface of ``smash_b200.distributed.NcclComm`` on top of a
``torch.distributed`` gloo group, so that the host-side logic of the N > 1 paths runs on a CPU-only machine."""
import numpy as np
import torch
import torch.distributed as dist

_OPS = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}


class GlooComm:
    def __init__(self, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def allreduce(self, a, op="sum"):
        t = torch.from_numpy(a)
        dist.all_reduce(t, op=_OPS[op], group=self.group)
        return a

    def allgather(self, a):
        a = np.ascontiguousarray(a)
        t = torch.from_numpy(a.reshape(-1))
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return np.stack([o.numpy() for o in out])

    def barrier(self):
        dist.barrier(group=self.group)
