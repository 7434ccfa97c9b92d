"""Data parallelism for the hot path: one process per GPU, replicated weights,
one gradient all-reduce per optimizer step over NCCL/NVLink, SyncBN statistics.

The reference is single-GPU (bbhMahoGANy.py:72-74); the only exchange the path needs is the
gradient (and BatchNorm statistic) sum, so this is a thin wrapper over torch.distributed.
`gloo` is accepted so the host-side logic can be tested with world_size 2 on CPU.
"""
import os

import torch
import torch.distributed as dist

from . import nn


class DataParallel:
    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_reduce(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def shard(self, n_global):
        """[lo, hi) of a global batch of n_global samples owned by this rank (even split)."""
        assert n_global % self.world == 0, 'global batch must divide evenly across ranks'
        per = n_global // self.world
        return self.rank * per, (self.rank + 1) * per


def init_data_parallel(backend=None):
    """Join the process group described by RANK/WORLD_SIZE/MASTER_* (torchrun) and make every
    subsequently compiled model data-parallel. No-op for WORLD_SIZE <= 1."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        nn._STATE['dp'] = None
        return None
    if not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group(backend=backend)
    dp = DataParallel()
    nn._STATE['dp'] = dp
    return dp


def shutdown():
    nn._STATE['dp'] = None
    if dist.is_initialized():
        dist.destroy_process_group()


def broadcast_weights(model, src=0):
    """Make every rank start from rank `src`'s weights (Keras would have one initialisation)."""
    dp = nn._STATE['dp']
    if dp is None:
        return
    for p in model.params:
        dist.broadcast(p.data, src=src, group=dp.group)
