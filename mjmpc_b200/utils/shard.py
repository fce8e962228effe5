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


class PeerExchange:
    """Symmetric NVLink peer-memory buffer for the fused exchange+combine kernel
    (``mjb_softmax_exchange_combine``): ``[2][world][P]`` doubles + ``[2][world]`` sequence flags on every
    rank, mapped into every peer through ``torch.distributed._symmetric_memory`` (plumbing only -- the
    loads / stores over NVLink are issued by our kernel).  Construction is collective."""

    def __init__(self, shard: "ShardContext", P: int, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.P, self.world, self.rank = int(P), shard.world_size, shard.rank
        n = 2 * self.world * self.P + 2 * self.world
        self.buf = symm_mem.empty(n, dtype=torch.float64, device=device)
        self.buf.zero_()
        group = shard.group if shard.group is not None else dist.group.WORLD
        self.handle = symm_mem.rendezvous(self.buf, group)
        ptrs = self.handle.buffer_ptrs_dev
        self.peer_ptrs_dev = int(ptrs)                       # device array of `world` buffer pointers
        self.seq = 0
        torch.cuda.synchronize(device)
        dist.barrier(group=group)                            # every rank has zeroed its flags before first use

    def next_seq(self) -> int:
        self.seq += 1
        return self.seq


SINGLE = ShardContext()
