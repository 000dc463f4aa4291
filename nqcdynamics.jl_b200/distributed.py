"""Sharding of an ensemble over ranks (one process per GPU) and the single collective of the path.

Trajectories are independent (docs/src/ensemble_simulations.md:25-28), so rank g of G owns the contiguous global
index block [g*T/G, (g+1)*T/G) -- `traj_offset` keys the Philox stream, which makes every trajectory's result
independent of G -- and the only exchange is one all-reduce(sum) of the observable accumulator
(SumReduction / MeanReduction, src/Ensembles/reductions.jl:12-52).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(ntraj: int, world_size: int, rank: int):
    """Contiguous block of global trajectory indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(ntraj), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum(array, group=None):
    """All-reduce a host numpy array or a torch tensor (device accumulators over NCCL, host arrays over gloo)."""
    import torch
    import torch.distributed as dist
    if isinstance(array, np.ndarray):
        t = torch.from_numpy(array)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return array
    dist.all_reduce(array, op=dist.ReduceOp.SUM, group=group)
    return array


class DeviceArray:
    """Expose a raw device pointer (nqcb200_observable_sum_device) to torch via __cuda_array_interface__."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
