"""Distributed runtime: one process per GPU, NCCL/NVLink plumbing via torch.distributed, the
gradient all-reduce by our own fused symmetric-memory kernels.

Horovod surface used by the reference (P1/03:283-322, P2/02:170-211)  ->  here:

    hvd.init / rank / size / local_rank          init / rank / size / local_rank
    hvd.DistributedOptimizer(opt)                DistributedOptimizer(opt)
    hvd.callbacks.BroadcastGlobalVariables...    callbacks.BroadcastGlobalVariablesCallback
    hvd.callbacks.MetricAverageCallback          callbacks.MetricAverageCallback
    hvd.callbacks.LearningRateWarmupCallback     callbacks.LearningRateWarmupCallback
    sparkdl.HorovodRunner(np).run(fn, **kw)      Runner(np).run(fn, **kw)
"""
from .core import (init, shutdown, is_initialized, rank, size, local_rank, barrier, allreduce, broadcast,
                   broadcast_object, allgather_object, device)
from .runner import Runner, HorovodRunner
from .dist_optimizer import DistributedOptimizer
from . import callbacks

__all__ = ["init", "shutdown", "is_initialized", "rank", "size", "local_rank", "barrier", "allreduce", "broadcast",
           "broadcast_object", "allgather_object", "device", "Runner", "HorovodRunner", "DistributedOptimizer",
           "callbacks"]
