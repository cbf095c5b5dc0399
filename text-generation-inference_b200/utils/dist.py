"""Process-group bring-up: one process per GPU, NCCL over NVLink 5 / NVSwitch on GPU, Gloo on CPU (tests).

Mirrors /root/reference/server/text_generation_server/utils/dist.py:21-96 (FakeGroup for world size 1,
high-priority NCCL stream, 60 s timeout).
"""
from __future__ import annotations

import os
from datetime import timedelta

import torch
import torch.distributed

RANK = int(os.getenv("RANK", "0"))


class FakeBarrier:
    def wait(self):
        pass


class FakeGroup:
    """utils/dist.py:21-44: lets the tensor-parallel layers run un-sharded."""

    def __init__(self, rank: int, size: int):
        self._rank = rank
        self._size = size

    def allreduce(self, *args, **kwargs):
        return FakeBarrier()

    def allgather(self, inputs, local_tensor, **kwargs):
        for input_ in inputs:
            input_[0].data = local_tensor[0].data
        return FakeBarrier()

    def barrier(self, *args, **kwargs):
        return FakeBarrier()

    def size(self):
        return self._size

    def rank(self):
        return self._rank


def print_rank_n(*values, rank: int = 0) -> None:
    if RANK == rank:
        print(*values)


def get_torch_dtype(dtype_str: str) -> torch.dtype:
    dt = getattr(torch, dtype_str, None)
    if type(dt) != torch.dtype:
        raise ValueError(f"Unrecognized data type: {dtype_str}")
    return dt


def initialize_torch_distributed(world_size: int, rank: int):
    """utils/dist.py:70-96."""
    if world_size == 1 or os.getenv("DEBUG", None) == "1":
        return FakeGroup(rank, world_size)
    if not torch.distributed.is_initialized():
        if torch.cuda.is_available():
            from torch.distributed import ProcessGroupNCCL
            backend = "nccl"
            options = ProcessGroupNCCL.Options()
            options.is_high_priority_stream = True
            torch.cuda.set_device(int(os.getenv("LOCAL_RANK", rank)) % torch.cuda.device_count())
            kwargs = dict(pg_options=options)
        else:
            backend = "gloo"
            kwargs = {}
        torch.distributed.init_process_group(backend=backend, world_size=world_size, rank=rank,
                                             timeout=timedelta(seconds=60), **kwargs)
    return torch.distributed.group.WORLD
