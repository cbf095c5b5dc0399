"""Process-group bring-up: one process per GPU, NCCL over NVLink 5 / NVSwitch on GPU, Gloo on CPU (tests).

Same behaviour as /root/reference/server/text_generation_server/utils/dist.py:21-96: world size 1 (or DEBUG=1) gets a
`FakeGroup` so the tensor-parallel layers run un-sharded without a backend; otherwise the WORLD group is initialised
from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (set by the launcher, launcher/src/main.rs:679-714) with a 60 s
timeout and, on GPU, NCCL's high-priority stream.
"""
from __future__ import annotations

import os
from datetime import timedelta

import torch
import torch.distributed

RANK = int(os.getenv("RANK", "0"))
_TIMEOUT = timedelta(seconds=60)


class _Completed:
    """What a collective of the fake group returns: a work handle that is already done."""

    def wait(self):
        return None


FakeBarrier = _Completed  # the reference's name for it


class FakeGroup:
    """A process group of one: every collective is the identity (utils/dist.py:21-44)."""

    def __init__(self, rank: int, size: int):
        self._rank, self._size = rank, size

    def rank(self):
        return self._rank

    def size(self):
        return self._size

    def barrier(self, *args, **kwargs):
        return _Completed()

    def allreduce(self, *args, **kwargs):
        return _Completed()

    def allgather(self, inputs, local_tensor, **kwargs):
        for gathered in inputs:
            gathered[0].data = local_tensor[0].data
        return _Completed()


def print_rank_n(*values, rank: int = 0) -> None:
    if RANK == rank:
        print(*values)


def get_torch_dtype(dtype_str: str) -> torch.dtype:
    found = getattr(torch, dtype_str, None)
    if not isinstance(found, torch.dtype):
        raise ValueError(f"Unrecognized data type: {dtype_str}")
    return found


def initialize_torch_distributed(world_size: int, rank: int):
    if world_size == 1 or os.getenv("DEBUG") == "1":
        return FakeGroup(rank, world_size)
    if torch.distributed.is_initialized():
        return torch.distributed.group.WORLD
    extra = {}
    backend = "gloo"
    if torch.cuda.is_available():
        from torch.distributed import ProcessGroupNCCL
        torch.cuda.set_device(int(os.getenv("LOCAL_RANK", rank)) % torch.cuda.device_count())
        nccl = ProcessGroupNCCL.Options()
        nccl.is_high_priority_stream = True
        backend, extra = "nccl", {"pg_options": nccl}
    torch.distributed.init_process_group(backend=backend, world_size=world_size, rank=rank, timeout=_TIMEOUT, **extra)
    return torch.distributed.group.WORLD
