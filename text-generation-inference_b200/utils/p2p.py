"""The tensor-parallel layer boundary over NVLink peer memory (csrc/p2p_allreduce.cu), with NCCL for what it does not take.

The reference calls `torch.distributed.all_reduce(out, group=self.process_group)` after every row-parallel linear and the
vocab-parallel embedding (utils/layers.py:303-306, :343-345) and all-gathers the sharded head's logits (:249-269).
  * `LayerBoundaryAllReduce`: that all-reduce; decode-sized fp16 messages go through the library's one-shot peer-memory kernel,
    everything else (prefill, other dtypes) through NCCL.
  * `FusedBoundary`: the windows the C++ step runtime uses to run the whole sharded decode step without a host-side
    collective: all-reduce + residual + RMSNorm in one kernel (b200_p2p_allreduce_rmsnorm) and the greedy head as a
    (value, index) exchange (b200_p2p_argmax).
B200_P2P_ALLREDUCE=0 keeps every collective on NCCL.
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed

from .. import _lib

P2P_MAX_BYTES = 2 << 20  # messages up to 2 MiB (bs 256 x hidden 4096 fp16) go over peer memory; prefill stays on NCCL


def p2p_requested() -> bool:
    return os.environ.get("B200_P2P_ALLREDUCE", "1") != "0"


def _open_window(process_group, max_bytes: int):
    """b200_p2p_create + IPC handle exchange over the group + b200_p2p_connect -> context handle."""
    lib = _lib.load()
    world, rank = process_group.size(), process_group.rank()
    n = lib.b200_p2p_handle_bytes()
    mine = (ctypes.c_ubyte * n)()
    ctx = ctypes.c_void_p()
    _lib.check(lib.b200_p2p_create(max_bytes, world, rank, ctypes.byref(ctx), mine), "p2p_create")
    local = torch.tensor(list(mine), dtype=torch.uint8, device="cuda")
    gathered = torch.empty(world * n, dtype=torch.uint8, device="cuda")
    torch.distributed.all_gather_into_tensor(gathered, local, group=process_group)
    handles = (ctypes.c_ubyte * (world * n)).from_buffer_copy(bytes(gathered.cpu().tolist()))
    _lib.check(lib.b200_p2p_connect(ctx, handles), "p2p_connect")
    torch.distributed.barrier(group=process_group)  # every window is mapped everywhere before the first use
    return ctx


class FusedBoundary:
    """Windows for the in-step layer boundary.  `norm` serves b200_p2p_allreduce_rmsnorm for up to MAX_ROWS token rows of
    `hidden_size` (decode steps and the prefill chunks of continuous batching; 7 x MAX_ROWS x hidden x 2 bytes per rank);
    `argmax` serves b200_p2p_argmax.  Both are None when the group has one rank or p2p is switched off."""

    MAX_ROWS = 2048

    def __init__(self, process_group, hidden_size: int):
        self.norm = self.argmax = None
        if process_group.size() > 1 and p2p_requested() and torch.cuda.is_available():
            self.norm = _open_window(process_group, self.MAX_ROWS * hidden_size * 2)
            self.argmax = _open_window(process_group, 8192)


class LayerBoundaryAllReduce:
    def __init__(self, process_group, max_bytes: int = P2P_MAX_BYTES):
        self.process_group = process_group
        self.world = process_group.size()
        self._ctx = None
        self._max_bytes = 0
        if self.world > 1 and p2p_requested():
            if torch.cuda.is_available():  # gloo / CPU groups (host-logic tests) stay on torch.distributed
                self._connect(max_bytes)

    def _connect(self, max_bytes: int) -> None:
        self._ctx = _open_window(self.process_group, max_bytes)
        self._max_bytes = _lib.load().b200_p2p_max_bytes(self._ctx)

    @property
    def uses_peer_memory(self) -> bool:
        return self._ctx is not None

    def __call__(self, tensor: torch.Tensor) -> torch.Tensor:
        """In-place sum over the ranks of the group."""
        if self.world == 1:
            return tensor
        nbytes = tensor.numel() * tensor.element_size()
        if (self._ctx is not None and tensor.dtype == torch.float16 and tensor.is_contiguous() and tensor.numel() % 8 == 0
                and tensor.data_ptr() % 16 == 0 and nbytes <= self._max_bytes):
            _lib.check(_lib.load().b200_p2p_allreduce_f16(self._ctx, tensor.data_ptr(), tensor.numel(),
                                                          torch.cuda.current_stream().cuda_stream), "p2p_allreduce")
        else:
            torch.distributed.all_reduce(tensor, group=self.process_group)
        return tensor
