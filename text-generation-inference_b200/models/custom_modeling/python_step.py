"""What the flash families driven op by op from Python (GPT-NeoX, Santacoder, Falcon) share.

1. The attention core (`paged_attention`: fused rotary + KV append, then varlen prefill or paged decode attention, with the
   16-query-heads-per-launch split for multi-query models), the GELU MLP, the row-parallel loader, the layer loop and the
   outer `...ForCausalLM` module (`FlashFamilyForCausalLM`).  The family files keep the reference's class names, constructor
   arguments and checkpoint prefixes and express their graphs with these pieces.
2. Fused greedy decode.  `FlashCausalLM._decode_fused_greedy` drives a model through `make_step` / `run_step`: the step writes
   logits and the arg-max ids into the caller's buffers, the ids chain on the device, and from the third step of a stable
   batch the whole step is replayed as a CUDA graph.  FlashLlama implements the protocol with the C++ step runtime
   (csrc/llama_step.cu); `PythonFusedGreedy` implements it by enqueuing the family's ordinary decode forward, whose torch
   temporaries then come from the graph's private pool.  On by default on a single rank (validated on B200 for GPT-NeoX and
   Santacoder, round 2); sharded models take it with B200_PY_FUSED_STEP=1 only (NCCL inside the captured graph), and
   B200_PY_FUSED_STEP=0 turns it off.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from ...utils import _ops
from ...utils.flash_attn import PagedKVLayer, attention
from ...utils.layers import TensorParallelRowLinear, get_linear
from ...utils.paged import PagedKVState

MAX_GROUP = 16  # query heads that can share one KV head in a single decode-attention launch (csrc/attn_decode.cu)


@dataclass
class PythonStep:
    """What `make_step` hands to FlashCausalLM for these families: the caller's index tensors and output buffers."""
    T: int
    B: int
    max_s: int
    input_ids: torch.Tensor
    position_ids: torch.Tensor
    kv: PagedKVState
    logits: torch.Tensor
    next_ids: Optional[torch.Tensor]
    banned: Optional[torch.Tensor]  # per-row banned id of the arg-max (min_new_tokens EOS mask), set by the caller
    decode_marker: torch.Tensor
    banned_ids: int = 0             # the C struct's field of the same name (FlashCausalLM sets both)


class _NoScratch:
    version = 0  # FlashCausalLM keys its cached step on this; nothing here is ever re-allocated


class PythonFusedGreedy:
    """Mixin for a `...ForCausalLM` module: needs `self.model` (the backbone, called like the reference's forward),
    `self.lm_head` (a TensorParallelHead) and `self.process_group`."""
    scratch = _NoScratch()

    @property
    def fused_greedy_enabled(self) -> bool:
        switch = os.environ.get("B200_PY_FUSED_STEP", "")
        if switch == "0":
            return False
        return switch == "1" or self.process_group.size() == 1

    def make_step(self, *, T: int, B: int, is_prefill: bool, max_s: int, input_ids, position_ids, kv: PagedKVState,
                  cu_seqlens=None, head_rows=None, logits=None, next_ids=None, inputs_embeds=None) -> PythonStep:
        if is_prefill or head_rows is not None or inputs_embeds is not None:
            raise NotImplementedError("the Python fused step is decode-only")
        # any non-None cu_seqlens_q selects the decode branch of the attention modules; the kernels never read it
        marker = torch.zeros(1, dtype=torch.int32, device=input_ids.device)
        return PythonStep(T=T, B=B, max_s=int(max_s), input_ids=input_ids, position_ids=position_ids, kv=kv, logits=logits,
                          next_ids=next_ids, banned=None, decode_marker=marker)

    def run_step(self, s: PythonStep) -> None:
        hidden, _ = self.model(s.input_ids, s.position_ids, None, s.decode_marker, s.max_s, None, s.kv)
        s.logits.copy_(self.lm_head.linear(hidden))  # this rank's vocab rows; FlashCausalLM gathers them when sharded
        if s.next_ids is not None:
            _ops().argmax(s.logits, s.banned, out=s.next_ids)


def row_parallel_linear(config, prefix: str, weights, bias: bool, reduces_itself: bool):
    """A projection whose INPUT dimension is sharded.  Its bias is loaded on the first rank only, so that it is counted once
    in the sum over ranks.  `reduces_itself`: wrap it so that it all-reduces its own output; layers that add the attention
    and MLP branches first and all-reduce the sum once (parallel residual / parallel attention) take the bare linear."""
    weight = weights.get_multi_weights_row(prefix, quantize=config.quantize)
    first_rank_bias = weights.get_tensor(f"{prefix}.bias") if bias and weights.process_group.rank() == 0 else None
    linear = get_linear(weight, first_rank_bias, config.quantize)
    return TensorParallelRowLinear(linear, process_group=weights.process_group) if reduces_itself else linear


class FlashFamilyForCausalLM(PythonFusedGreedy, nn.Module):
    """Outer module of a family driven from Python: a backbone (`self.model`, subclass property) that returns the final hidden
    states, and a vocabulary head (`self.lm_head`).  Carries what FlashCausalLM and the server read on a flash model."""

    def _init_outer(self, config, weights, default_max_positions: int = 2048):
        self.config = config
        self.process_group = weights.process_group
        self.device = torch.device(weights.device)
        self.max_positions = int(getattr(config, "max_position_embeddings", None) or getattr(config, "n_positions", None)
                                 or default_max_positions)

    @property
    def kv_cache_manager(self):
        return self.model.kv_cache_manager

    @kv_cache_manager.setter
    def kv_cache_manager(self, manager):
        self.model.kv_cache_manager = manager

    def forward(self, input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds: Optional[torch.Tensor] = None,
                past_key_values: Optional[PagedKVState] = None, pre_allocate_past_size: Optional[int] = None,
                lm_head_indices: Optional[torch.Tensor] = None):
        hidden, present = self.model(input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds, past_key_values,
                                     pre_allocate_past_size)
        rows = hidden if lm_head_indices is None else hidden[lm_head_indices]  # prefill: only the last token of each prompt
        return self.lm_head(rows), present


def paged_attention(qkv, n_heads: int, n_kv: int, head_size: int, softmax_scale: float, cos_table, sin_table, position_ids,
                    cu_seqlens, max_s, kv: PagedKVState, k_pool, v_pool, cu_seqlens_q, rotary_dim: Optional[int] = None):
    """The attention core every family shares once its fused projection is laid out [q heads | k heads | v heads]:
    one kernel rotates q and k in place (the first `rotary_dim` elements of each head; the whole head by default) and appends
    K / V to the paged pool, then varlen causal attention over this step's tokens (prefill: `cu_seqlens_q is None`) or paged
    attention of one query token per sequence over its cached context, the new token included (decode).
    Returns [T, n_heads, head_size]."""
    d = head_size
    _ops().rope_kv_write_paged(qkv, cos_table, sin_table, position_ids, kv.slot_mapping, k_pool, v_pool, n_heads, n_kv, d,
                               rotary_dim=rotary_dim)
    query = qkv[:, :n_heads * d].unflatten(1, (n_heads, d))
    if cu_seqlens_q is None:
        key = qkv[:, n_heads * d:(n_heads + n_kv) * d].unflatten(1, (n_kv, d))
        value = qkv[:, (n_heads + n_kv) * d:].unflatten(1, (n_kv, d))
        return attention(query, key, value, cu_seqlens, max_s, softmax_scale)
    cache = PagedKVLayer(k_pool, v_pool, kv.block_table, kv.context_lens, int(max_s))
    if n_heads // n_kv <= MAX_GROUP:
        return attention(query, cache, None, cu_seqlens, max_s, softmax_scale, cu_seqlens_q, 1, False)
    if n_kv != 1:
        raise NotImplementedError(f"{n_heads // n_kv} query heads per KV head with {n_kv} KV heads: more than {MAX_GROUP} per "
                                  "launch is only served for a single shared KV head")
    out = torch.empty(qkv.shape[0], n_heads, d, dtype=qkv.dtype, device=qkv.device)
    for first in range(0, n_heads, MAX_GROUP):  # multi-query: any slice of the query heads shares the one KV head
        last = min(n_heads, first + MAX_GROUP)
        attention(query[:, first:last], cache, None, cu_seqlens, max_s, softmax_scale, cu_seqlens_q, 1, False, out=out[:, first:last])
    return out


def gelu_mlp(up, down, hidden_states, approximate_tanh: bool):
    """down(GELU(up(x))): the two-projection MLP of the GELU families; `approximate_tanh` for gelu_fast / gelu_pytorch_tanh."""
    return down(_ops().gelu(up(hidden_states), approximate_tanh))


def gelu_is_tanh(activation: str, where: str) -> bool:
    if "gelu" not in activation:
        raise NotImplementedError(f"activation {activation!r}: only the GELU variants of {where} are built")
    return activation in ("gelu_fast", "gelu_pytorch_tanh")


def run_layers(layers, manager, hidden_states, *per_step):
    """hidden, residual through the layer stack; every layer gets its own (k_pool, v_pool) of the paged cache."""
    residual = None
    for index, layer in enumerate(layers):
        k_pool, v_pool = manager.layer_pools(index)
        hidden_states, residual = layer(hidden_states, residual, *per_step, k_pool, v_pool)
    return hidden_states, residual
