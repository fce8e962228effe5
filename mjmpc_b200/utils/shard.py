"""Particle sharding across the GPUs of one box (one process per GPU).

The reference splits particles into contiguous blocks over worker processes
(``mjmpc/envs/vec_env/subproc_vec_env.py:161-168``) and reduces on the parent.  Here rank g
owns particles ``[g*K/N, (g+1)*K/N)``, rolls them out locally with no communication, and the
distribution update exchanges one small partial vector per rank (all-gather), combined in rank
order so every rank ends with bit-identical parameters.  ``torch.distributed`` is the plumbing
(NCCL on GPUs; gloo in the CPU tests of the host-side logic).
"""
from __future__ import annotations

import torch


class ShardContext:
    def __init__(self, rank: int = 0, world_size: int = 1, group=None):
        self.rank = int(rank)
        self.world_size = int(world_size)
        self.group = group

    @classmethod
    def from_env(cls):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return cls(dist.get_rank(), dist.get_world_size())
        return cls()

    def local_range(self, K: int):
        """(first global index, local count) of this rank's contiguous particle block."""
        if K % self.world_size != 0:
            raise AssertionError("Number of particles must be divisible by number of GPUs")
        per = K // self.world_size
        return self.rank * per, per

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        """(world, *t.shape) stack of every rank's tensor, in rank order."""
        if self.world_size == 1:
            return t.reshape((1,) + tuple(t.shape))
        import torch.distributed as dist
        t = t.contiguous()
        flat = torch.empty(self.world_size * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(flat, t.reshape(-1), group=self.group)
        return flat.reshape((self.world_size,) + tuple(t.shape))


SINGLE = ShardContext()
