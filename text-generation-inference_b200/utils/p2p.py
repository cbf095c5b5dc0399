"""All-reduce at the tensor-parallel layer boundary: NCCL, or (B200_P2P_ALLREDUCE=1, experimental) the library's one-shot
kernel over NVLink peer memory for decode-sized messages (csrc/p2p_allreduce.cu).

The reference calls `torch.distributed.all_reduce(out, group=self.process_group)` after every row-parallel linear and the
vocab-parallel embedding (utils/layers.py:303-306, :343-345).  `LayerBoundaryAllReduce` keeps that call as the default and
as the path for anything the peer-memory kernel does not take (large prefill messages, dtypes other than fp16).
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed

from .. import _lib

P2P_MAX_BYTES = 2 << 20  # messages up to 2 MiB (bs 256 x hidden 4096 fp16) go over peer memory; prefill stays on NCCL


def p2p_requested() -> bool:
    return os.environ.get("B200_P2P_ALLREDUCE", "0") == "1"


class LayerBoundaryAllReduce:
    def __init__(self, process_group, max_bytes: int = P2P_MAX_BYTES):
        self.process_group = process_group
        self.world = process_group.size()
        self._ctx = None
        self._max_bytes = 0
        if self.world > 1 and p2p_requested():
            if not torch.cuda.is_available():
                raise RuntimeError("B200_P2P_ALLREDUCE=1 needs CUDA devices: the peer-memory all-reduce has no CPU form")
            self._connect(max_bytes)

    def _connect(self, max_bytes: int) -> None:
        lib = _lib.load()
        rank = self.process_group.rank()
        n = lib.b200_p2p_handle_bytes()
        mine = (ctypes.c_ubyte * n)()
        ctx = ctypes.c_void_p()
        _lib.check(lib.b200_p2p_create(max_bytes, self.world, rank, ctypes.byref(ctx), mine), "p2p_create")
        local = torch.tensor(list(mine), dtype=torch.uint8, device="cuda")
        gathered = torch.empty(self.world * n, dtype=torch.uint8, device="cuda")
        torch.distributed.all_gather_into_tensor(gathered, local, group=self.process_group)
        handles = (ctypes.c_ubyte * (self.world * n)).from_buffer_copy(bytes(gathered.cpu().tolist()))
        _lib.check(lib.b200_p2p_connect(ctx, handles), "p2p_connect")
        torch.distributed.barrier(group=self.process_group)  # every window is mapped everywhere before the first use
        self._ctx = ctx
        self._max_bytes = lib.b200_p2p_max_bytes(ctx)

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
