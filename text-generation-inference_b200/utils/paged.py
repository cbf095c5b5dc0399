"""Paged KV cache manager: block-16 pools per layer, per-sequence block lists, device-resident step state.

Stands in for the un-vendored fms-extras `PagedKVCacheManager` that the reference's paged path uses
(/root/reference/server/text_generation_server/models/paged_causal_lm.py:338-353 constructor arguments,
utils/paged.py:92-159 `allocate_tokens`, server.py:233-249 `free_sequences`).  Block bookkeeping is the C++
allocator behind the C ABI (csrc/kv_alloc.cu); this class owns the pools and the per-sequence lists.

HBM layout (DESIGN.md "KV page layout"): one tensor [n_layers, 2, num_blocks, n_kv_heads/tp, 16, head_dim] fp16;
a (block, kv head) tile is 16*d*2 contiguous bytes, its 16-byte chunks XOR-swizzled by (token & 7).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from .. import _lib

BLOCK_SIZE = 16


class OutOfBlocks(RuntimeError):
    """Raised when the pool cannot hold the requested tokens (the server maps it to RESOURCE_EXHAUSTED)."""


@dataclass
class PagedKVState:
    """Device-side view of a batch's KV: what `past_key_values` is in this implementation."""
    sequence_ids: List[int]
    block_table: torch.Tensor   # [B, max_blocks] int32
    context_lens: torch.Tensor  # [B] int32: tokens cached (decode: advanced by the step's bookkeeping kernel)
    slot_mapping: torch.Tensor  # [T] int64 for the tokens of the current step
    max_blocks: int


class PagedKVCacheManager:
    def __init__(self, num_layers: int, num_heads: int, emb_dim: int, kv_heads: int = 0, tensor_parallel_size: int = 1,
                 dtype: torch.dtype = torch.float16, device="cuda", total_num_gpu_blocks: Optional[int] = None,
                 block_size: int = BLOCK_SIZE):
        if block_size != BLOCK_SIZE:
            raise ValueError("the B200 kernels are built for block_size 16 (models/paged_causal_lm.py:308)")
        if dtype != torch.float16:
            raise ValueError("KV cache dtype must be float16")
        self.block_size = block_size
        self.num_layers = num_layers
        self.head_dim = emb_dim // num_heads
        kv_heads = kv_heads or num_heads
        if kv_heads < tensor_parallel_size or kv_heads % tensor_parallel_size:
            # flash_llama_modeling.py:220-222 silently floors; guard it
            raise ValueError(f"num_key_value_heads {kv_heads} not divisible by tensor parallel size {tensor_parallel_size}")
        self.kv_heads = kv_heads // tensor_parallel_size
        self.device = torch.device(device)
        if total_num_gpu_blocks is None:
            free, _ = torch.cuda.mem_get_info(self.device)
            total_num_gpu_blocks = int(free * 0.8) // self.block_bytes()
        self.total_num_gpu_blocks = int(total_num_gpu_blocks)
        # zero-filled: masked tail slots must be finite (attn_decode.cu multiplies them by p = 0)
        self.pool = torch.zeros(num_layers, 2, self.total_num_gpu_blocks, self.kv_heads, block_size, self.head_dim,
                                dtype=dtype, device=self.device)
        self._alloc = _lib.load().b200_kv_alloc_create(self.total_num_gpu_blocks)
        if not self._alloc:
            raise _lib.B200Error("kv_alloc_create failed")
        self._blocks: Dict[int, List[int]] = {}
        self._lens: Dict[int, int] = {}
        self._next_id = 0

    def block_bytes(self) -> int:
        """bytes of one 16-token block across all layers, K and V (get_kv_cache_block_size * n_layers * dtype size)."""
        return self.num_layers * 2 * self.kv_heads * self.block_size * self.head_dim * 2

    @property
    def free_blocks(self) -> int:
        return int(_lib.load().b200_kv_alloc_num_free(self._alloc))

    def __del__(self):
        try:
            if getattr(self, "_alloc", None):
                _lib.load().b200_kv_alloc_destroy(self._alloc)
                self._alloc = None
        except Exception:
            pass

    # -- strides handed to the C step runtime
    @property
    def layer_stride_bytes(self) -> int:
        return self.pool.stride(0) * 2

    @property
    def v_offset_bytes(self) -> int:
        return self.pool.stride(1) * 2

    def layer_pools(self, layer: int):
        return self.pool[layer, 0], self.pool[layer, 1]

    # -- allocation
    def _take(self, n: int) -> List[int]:
        if n == 0:
            return []
        buf = (ctypes.c_int32 * n)()
        st = _lib.load().b200_kv_alloc_take(self._alloc, n, buf)
        if st != 0:
            raise OutOfBlocks(f"KV cache exhausted: need {n} blocks, {self.free_blocks} free")
        return list(buf)

    def _release(self, ids: List[int]) -> None:
        if ids:
            buf = (ctypes.c_int32 * len(ids))(*ids)
            _lib.check(_lib.load().b200_kv_alloc_release(self._alloc, buf, len(ids)), "kv_alloc_release")

    def blocks_needed(self, num_tokens: int) -> int:
        return (num_tokens + self.block_size - 1) // self.block_size

    def allocate_tokens(self, num_tokens_per_sequence: List[int], sequence_ids: Optional[List[int]] = None,
                        reserve_tokens: Optional[List[int]] = None) -> List[int]:
        """Extends existing sequences (or creates new ones when sequence_ids is None) by the given token counts.
        Returns the sequence ids.  `reserve_tokens[i]` additionally pre-books blocks for future tokens so a decode
        loop never has to touch the block table (continuous batching sizes a request by input + max_output)."""
        new = sequence_ids is None
        if new:  # ids are only registered once the blocks are booked: a rejected prefill (OutOfBlocks) leaves no trace
            sequence_ids = list(range(self._next_id, self._next_id + len(num_tokens_per_sequence)))
        need = []
        for i, (sid, n) in enumerate(zip(sequence_ids, num_tokens_per_sequence)):
            have_len, have_blocks = (0, 0) if new else (self._lens[sid], len(self._blocks[sid]))
            target = have_len + n + (reserve_tokens[i] if reserve_tokens else 0)
            need.append(max(0, self.blocks_needed(target) - have_blocks))
        got = self._take(sum(need))  # all or nothing
        if new:
            for sid in sequence_ids:
                self._blocks[sid] = []
                self._lens[sid] = 0
            self._next_id += len(sequence_ids)
        pos = 0
        for sid, n, k in zip(sequence_ids, num_tokens_per_sequence, need):
            self._blocks[sid].extend(got[pos:pos + k])
            pos += k
            self._lens[sid] += n
        return sequence_ids

    def free_sequences(self, sequence_ids: List[int], recursive: bool = False) -> None:
        for sid in sequence_ids:
            blocks = self._blocks.pop(sid, None)
            self._lens.pop(sid, None)
            if blocks:
                self._release(blocks)

    def sequence_length(self, sid: int) -> int:
        return self._lens[sid]

    def sequence_blocks(self, sid: int) -> List[int]:
        return self._blocks[sid]

    def note_decode_step(self, sequence_ids: List[int]) -> bool:
        """Host mirror of the device bookkeeping: one more token per sequence.  Returns True when any sequence
        needed a new block (the caller must then refresh the device block table)."""
        grew = False
        for sid in sequence_ids:
            if self.blocks_needed(self._lens[sid] + 1) > len(self._blocks[sid]):
                self._blocks[sid].extend(self._take(1))
                grew = True
            self._lens[sid] += 1
        return grew

    # -- device tensors
    def block_table_tensor(self, sequence_ids: List[int], min_cols: int = 1) -> torch.Tensor:
        cols = max([len(self._blocks[s]) for s in sequence_ids] + [min_cols])
        bt = torch.zeros(len(sequence_ids), cols, dtype=torch.int32)
        for i, s in enumerate(sequence_ids):
            b = self._blocks[s]
            bt[i, :len(b)] = torch.tensor(b, dtype=torch.int32)
        return bt.to(self.device, non_blocking=True)

    def slot_mapping_for(self, sequence_ids: List[int], starts: List[int], counts: List[int]) -> torch.Tensor:
        """slots of tokens [start, start+count) of each sequence, concatenated (prefill: start 0, count = prompt)."""
        out = []
        for s, st, n in zip(sequence_ids, starts, counts):
            blocks = torch.tensor(self._blocks[s], dtype=torch.int64)
            pos = torch.arange(st, st + n, dtype=torch.int64)
            out.append(blocks[pos // self.block_size] * self.block_size + pos % self.block_size)
        return torch.cat(out).to(self.device, non_blocking=True) if out else torch.empty(0, dtype=torch.int64, device=self.device)
