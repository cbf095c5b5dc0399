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


@dataclass
class PagedAttentionCacheData:
    """What fms-extras' `allocate_tokens` hands to the paged model classes (read at utils/paged.py:92-159, :194-257 and
    paged_llama_modeling.py:388-423 of the reference): per-step index tensors over the block pool.
    Prefill form: `context_lengths` = cumulative sequence lengths [B + 1] (utils/paged.py:148-157).  Generation form: one row per
    query TOKEN - `block_mapping` [T, max_blocks], `context_lengths` [T] = keys that token attends (itself included), so the n
    tokens of a speculative candidate are n rows over the same blocks with contexts L - n + 1 .. L (utils/paged.py:109-134, 230-241)."""
    sequence_ids: List[int]
    slot_mapping: torch.Tensor          # [T] int64 after flattening ([B, n] as allocated)
    block_mapping: torch.Tensor         # int32
    context_lengths: torch.Tensor       # int32
    position_ids: torch.Tensor          # int64
    max_sequence_length: int
    query_length: int
    is_generating: bool
    unflatten_indices: Optional[torch.Tensor] = None
    flatten_indices: Optional[torch.Tensor] = None

    def is_filled(self) -> bool:
        return self.is_generating


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
        # speculative decoding (models/paged_causal_lm.py:481-562): candidate sequences share their parent's blocks
        self._refs: Dict[int, int] = {}     # block id -> sequences holding it (absent = 1)
        self._parent: Dict[int, int] = {}   # child sequence id -> parent sequence id

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
        """drops one reference of every block; blocks nobody holds any more go back to the allocator"""
        free = []
        for b in ids:
            n = self._refs.get(b, 1) - 1
            if n <= 0:
                self._refs.pop(b, None)
                free.append(b)
            else:
                self._refs[b] = n
        if free:
            buf = (ctypes.c_int32 * len(free))(*free)
            _lib.check(_lib.load().b200_kv_alloc_release(self._alloc, buf, len(free)), "kv_alloc_release")

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
        # a sequence that is about to write into a block it shares with its parent / siblings gets its own copy first
        cow = [] if new else [sid for sid, n in zip(sequence_ids, num_tokens_per_sequence)
                              if n > 0 and self._lens[sid] % self.block_size != 0 and self._refs.get(self._blocks[sid][-1], 1) > 1]
        got = self._take(sum(need) + len(cow))  # all or nothing
        if new:
            for sid in sequence_ids:
                self._blocks[sid] = []
                self._lens[sid] = 0
            self._next_id += len(sequence_ids)
        pos = 0
        for sid in cow:
            old, fresh = self._blocks[sid][-1], got[pos]
            pos += 1
            self.pool[:, :, fresh].copy_(self.pool[:, :, old])
            self._blocks[sid][-1] = fresh
            self._release([old])
        for sid, n, k in zip(sequence_ids, num_tokens_per_sequence, need):
            self._blocks[sid].extend(got[pos:pos + k])
            pos += k
            self._lens[sid] += n
        return sequence_ids

    def free_sequences(self, sequence_ids: List[int], recursive: bool = False) -> None:
        """recursive: also the chain of parents a speculative sequence descends from (server.py:249)."""
        for sid in sequence_ids:
            while sid is not None:
                blocks = self._blocks.pop(sid, None)
                self._lens.pop(sid, None)
                if blocks:
                    self._release(blocks)
                sid = self._parent.pop(sid, None) if recursive else None

    def add_child_sequences(self, parent_sequence_id: int, num_children: int) -> List[int]:
        """`num_children` candidate sequences that start as copies of the parent (utils/paged.py:194-203 of the reference): they
        reference the parent's blocks; the first token a child appends to a shared partial block copies that block."""
        blocks, length = self._blocks[parent_sequence_id], self._lens[parent_sequence_id]
        used = blocks[:self.blocks_needed(length)]  # reserved-but-empty tail blocks stay with the parent
        children = list(range(self._next_id, self._next_id + num_children))
        self._next_id += num_children
        for cid in children:
            self._blocks[cid] = list(used)
            self._lens[cid] = length
            self._parent[cid] = parent_sequence_id
            for b in used:
                self._refs[b] = self._refs.get(b, 1) + 1
        return children

    def remove_tokens(self, sequence_id: int, num_tokens: int) -> None:
        """forgets the last `num_tokens` tokens of a sequence (rejected speculative tokens, utils/paged.py:309-315): the
        context shrinks, blocks that no longer hold a token are given back; the slots are simply overwritten later."""
        if num_tokens <= 0:
            return
        new_len = self._lens[sequence_id] - num_tokens
        if new_len < 0:
            raise ValueError(f"remove_tokens: sequence {sequence_id} has {self._lens[sequence_id]} tokens")
        keep = self.blocks_needed(new_len)
        drop = self._blocks[sequence_id][keep:]
        del self._blocks[sequence_id][keep:]
        self._lens[sequence_id] = new_len
        self._release(drop)

    def cache_data(self, sequence_ids: List[int], num_new_tokens: List[int], is_generating: bool) -> PagedAttentionCacheData:
        """Index tensors for the LAST num_new_tokens[i] tokens of every sequence (already allocated with allocate_tokens), in
        the form fms-extras returns them: slot_mapping / position_ids [B, n] padded with -1 / 0 for ragged counts (the callers
        flatten with truncate_and_flatten, utils/paged.py:100-108), block_mapping [B, max_blocks], context_lengths [B]."""
        n_max = max(num_new_tokens) if num_new_tokens else 0
        B = len(sequence_ids)
        slots = torch.full((B, n_max), -1, dtype=torch.int64)
        pos = torch.zeros((B, n_max), dtype=torch.int64)
        for i, (sid, n) in enumerate(zip(sequence_ids, num_new_tokens)):
            L = self._lens[sid]
            p = torch.arange(L - n, L, dtype=torch.int64)
            blocks = torch.tensor(self._blocks[sid], dtype=torch.int64)
            slots[i, :n] = blocks[p // self.block_size] * self.block_size + p % self.block_size
            pos[i, :n] = p
        lens = [self._lens[s] for s in sequence_ids]
        return PagedAttentionCacheData(
            sequence_ids=list(sequence_ids), slot_mapping=slots.to(self.device), block_mapping=self.block_table_tensor(sequence_ids),
            context_lengths=torch.tensor(lens, dtype=torch.int32, device=self.device), position_ids=pos.to(self.device),
            max_sequence_length=max(lens) if lens else 0, query_length=n_max, is_generating=is_generating)

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


# ---------------------------------------------------------------------------------------------------------------------
# Input preparation of the paged calling convention (the pure index arithmetic of utils/paged.py:82-159, 222-257, 259-326 in the
# reference; the speculator model itself, fms-extras' MLPSpeculator, is not part of this library).
# ---------------------------------------------------------------------------------------------------------------------
def prepare_inputs_for_prefill(num_tokens_per_sequence: List[int], kv_cache_manager: PagedKVCacheManager):
    """-> (position_ids [T], cache_data in prefill form): new sequences with their prompt tokens allocated (utils/paged.py:139-159)."""
    sids = kv_cache_manager.allocate_tokens(num_tokens_per_sequence)
    cd = kv_cache_manager.cache_data(sids, num_tokens_per_sequence, is_generating=False)
    keep = cd.slot_mapping.reshape(-1) >= 0  # ragged prompts: drop the padding of the [B, n_max] form
    position_ids = cd.position_ids.reshape(-1)[keep]
    cd.slot_mapping = cd.slot_mapping.reshape(-1)[keep]
    lens = torch.tensor([0] + list(num_tokens_per_sequence), dtype=torch.int32, device=cd.context_lengths.device)
    cd.context_lengths = torch.cumsum(lens, 0, dtype=torch.int32)
    return position_ids, cd


def expand_generation_cache_data(cd: PagedAttentionCacheData):
    """[B, n] allocation form -> one row per query token: token j of the n new tokens of a sequence of length L attends
    L - (n - 1 - j) keys over the sequence's blocks (utils/paged.py:109-134, 230-241).  -> (position_ids [B n], cache_data)."""
    n = cd.query_length
    B = cd.context_lengths.shape[0]
    back = torch.arange(n - 1, -1, -1, dtype=torch.int32, device=cd.context_lengths.device)
    cd.context_lengths = (cd.context_lengths.view(B, 1) - back.view(1, n)).reshape(-1).contiguous()
    cd.block_mapping = cd.block_mapping.repeat_interleave(n, dim=0).contiguous()
    cd.slot_mapping = cd.slot_mapping.reshape(-1)
    return cd.position_ids.reshape(-1), cd


def prepare_inputs_without_speculation(parent_sequence_ids: List[int], kv_cache_manager: PagedKVCacheManager):
    """one more token for every running sequence (utils/paged.py:82-136) -> (position_ids [B], cache_data in generation form)"""
    kv_cache_manager.allocate_tokens([1] * len(parent_sequence_ids), parent_sequence_ids)
    return expand_generation_cache_data(kv_cache_manager.cache_data(parent_sequence_ids, [1] * len(parent_sequence_ids), True))


def prepare_candidates(parent_sequence_ids: List[int], n_candidates: int, n_tokens: int, kv_cache_manager: PagedKVCacheManager):
    """speculative step: `n_candidates` child sequences per parent, `n_tokens` (= 1 + speculated) new tokens each
    (utils/paged.py:185-203).  -> (position_ids [B k n], cache_data in generation form, children per parent)."""
    children, flat = [], []
    for parent in parent_sequence_ids:
        kids = kv_cache_manager.add_child_sequences(parent, n_candidates)
        children.append(kids)
        flat.extend(kids)
    try:
        kv_cache_manager.allocate_tokens([n_tokens] * len(flat), flat)
    except BaseException:
        kv_cache_manager.free_sequences(flat)
        raise
    position_ids, cd = expand_generation_cache_data(kv_cache_manager.cache_data(flat, [n_tokens] * len(flat), True))
    return position_ids, cd, children


def accept_candidates(candidate_inputs: torch.Tensor, next_tokens: torch.Tensor, children: List[List[int]],
                      kv_cache_manager: PagedKVCacheManager):
    """The acceptance rule of utils/paged.py:279-324.  candidate_inputs [B, k, n]: the tokens fed for each candidate (the last
    accepted token followed by n - 1 speculated ones); next_tokens [B, k, n]: the model's greedy choice after each of them.
    A candidate's speculated token j + 1 is correct while it equals the model's choice after token j; the candidate with the
    longest correct prefix wins, the others are freed, the winner forgets its wrong tail.
    -> (surviving sequence ids [B], accepted new tokens per parent: the model's choices along the winner's correct prefix)."""
    B, k, n = candidate_inputs.shape
    agree = (candidate_inputs.roll(-1, 2) == next_tokens).cumprod(2)
    n_correct = agree.sum(2).clamp(0, n - 1)           # [B, k]; clamp: a wrap-around match of the rolled last column does not count
    best = n_correct.argmax(1)
    survivors, accepted = [], []
    for b, kids in enumerate(children):
        w = int(best[b])
        c = int(n_correct[b, w])
        kv_cache_manager.free_sequences(kids[:w] + kids[w + 1:])
        kv_cache_manager.remove_tokens(kids[w], n - c - 1)
        survivors.append(kids[w])
        accepted.append(next_tokens[b, w, :c + 1].tolist())
    return survivors, accepted
