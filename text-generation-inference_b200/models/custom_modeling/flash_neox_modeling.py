"""Flash GPT-NeoX on the B200 kernels: LayerNorm family, partial rotary, parallel residual with ONE all-reduce per layer.

Mirrors /root/reference/server/text_generation_server/models/custom_modeling/flash_neox_modeling.py: `load_row` (:40-54),
`load_qkv` (:57-80, the head-interleaved [h, 3, d] checkpoint layout re-laid-out to [3, h, d] at load), `FlashNeoxAttention`
(:83-170), `FlashMLP` (:173-199), `FlashNeoXLayer` (:202-281), `FlashGPTNeoXModel` (:291-389), `FlashGPTNeoXForCausalLM`
(:392-429) — same class names, constructor arguments `(config, weights)` and `forward` signature.  Differences, as for
FlashLlama: `past_key_values` is the batch's `PagedKVState` (block-table KV pool written by the fused RoPE + KV-write
kernel, read by the paged decode-attention kernel) instead of a contiguous tensor, and every op goes through the C ABI
(ops.py): `b200_layernorm_residual`, `b200_gemm_f16` / `b200_gemm_w4a16` (bias in the epilogue), `b200_rope_kv_write_paged_ex`
(partial rotation), `b200_attn_prefill_varlen` / `b200_attn_decode_paged`, `b200_gelu`.  This family runs op by op from
Python (no C++ step runtime / CUDA graph yet); the two residual adds of the parallel-residual branch are torch fp16 adds.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed
from torch import nn

from ...utils import _ops
from ...utils.flash_attn import PagedKVLayer, attention
from ...utils.layers import (FastLayerNorm, PositionRotaryEmbedding, TensorParallelColumnLinear, TensorParallelEmbedding,
                             TensorParallelHead, TensorParallelRowLinear, get_linear)
from ...utils.paged import PagedKVCacheManager, PagedKVState
from .python_step import PythonFusedGreedy


def load_row(config, prefix: str, weights, bias: bool):
    """flash_neox_modeling.py:40-54: bias lives on rank 0 only; with the parallel residual the all-reduce is done once per
    layer by the layer itself, so the bare linear is returned."""
    weight = weights.get_multi_weights_row(prefix, quantize=config.quantize)
    b = weights.get_tensor(f"{prefix}.bias") if bias and weights.process_group.rank() == 0 else None
    linear = get_linear(weight, b, config.quantize)
    if config.use_parallel_residual:
        return linear
    return TensorParallelRowLinear(linear, process_group=weights.process_group)


def load_qkv(config, prefix: str, weights, num_heads, head_size, hidden_size):
    """flash_neox_modeling.py:57-80: rows come head-interleaved [h, 3, d]; they are re-laid-out to [3, h, d] so that the
    product is [q heads | k heads | v heads] (what the RoPE / KV-write and attention kernels expect)."""
    weight = weights.get_multi_weights_col([prefix], quantize=config.quantize, dim=0)
    if isinstance(weight, torch.Tensor):  # only on non quantized versions
        weight = weight.view(num_heads, 3, head_size, hidden_size).permute(1, 0, 2, 3).reshape(-1, hidden_size)
    bias = weights.get_sharded(f"{prefix}.bias", dim=0)
    bias = bias.view(num_heads, 3, head_size).permute(1, 0, 2).reshape(-1)
    linear = get_linear(weight, bias.contiguous(), config.quantize)
    if config.use_parallel_residual:
        return linear
    return TensorParallelColumnLinear(linear)


class FlashNeoxAttention(nn.Module):
    def __init__(self, config, prefix, weights):
        super().__init__()
        num_heads = config.num_attention_heads
        self.hidden_size = config.hidden_size
        self.head_size = self.hidden_size // num_heads
        if num_heads % weights.process_group.size() != 0:
            raise ValueError(f"`num_heads` must be divisible by `num_shards` (got `num_heads`: {num_heads} "
                             f"and `num_shards`: {weights.process_group.size()}")
        self.num_heads = num_heads // weights.process_group.size()
        self.rotary_emb = PositionRotaryEmbedding.load(prefix=f"{prefix}.rotary_emb", weights=weights)
        self.rotary_dim = 2 * self.rotary_emb.inv_freq.shape[0]  # utils/layers.py:467-469: rotary_dim = cos.shape[-1] * 2
        self.softmax_scale = self.head_size ** (-0.5)
        self.query_key_value = load_qkv(config, prefix=f"{prefix}.query_key_value", weights=weights, num_heads=self.num_heads,
                                        head_size=self.head_size, hidden_size=self.hidden_size)
        self.dense = load_row(config, prefix=f"{prefix}.dense", weights=weights, bias=True)

    def forward(self, hidden_states, cos_table, sin_table, position_ids, cu_seqlens, max_s, kv: PagedKVState, k_pool, v_pool,
                cu_seqlens_q):
        qkv = self.query_key_value(hidden_states)  # [T, 3 * h * d] = [q | k | v]
        # in-place rotary on q and k (:127-129) + KV append (:137 / :150), one kernel
        _ops().rope_kv_write_paged(qkv, cos_table, sin_table, position_ids, kv.slot_mapping, k_pool, v_pool, self.num_heads,
                                   self.num_heads, self.head_size, rotary_dim=self.rotary_dim)
        q3 = qkv.view(-1, 3, self.num_heads, self.head_size)
        if cu_seqlens_q is None:  # prefill (:133-147)
            attn_output = attention(q3[:, 0], q3[:, 1], q3[:, 2], cu_seqlens, max_s, self.softmax_scale)
        else:  # decode (:149-166): one query token per sequence over the paged cache incl. the token just written
            layer = PagedKVLayer(k_pool, v_pool, kv.block_table, kv.context_lens, int(max_s))
            attn_output = attention(q3[:, 0], layer, None, cu_seqlens, max_s, self.softmax_scale, cu_seqlens_q, 1, False)
        return self.dense(attn_output.reshape(-1, self.num_heads * self.head_size))


class FlashMLP(nn.Module):
    def __init__(self, config, prefix, weights):
        super().__init__()
        act = config.hidden_act
        if "gelu" not in act:
            raise NotImplementedError(f"hidden_act {act!r}: only the GELU variants of flash_neox_modeling.py:186-196 are built")
        self.approximate_tanh = act in ["gelu_fast", "gelu_pytorch_tanh"]
        self.dense_h_to_4h = TensorParallelColumnLinear.load(config, prefix=f"{prefix}.dense_h_to_4h", weights=weights, bias=True)
        self.dense_4h_to_h = load_row(config, prefix=f"{prefix}.dense_4h_to_h", weights=weights, bias=True)

    def forward(self, hidden_states):
        hidden_states = self.dense_h_to_4h(hidden_states)
        hidden_states = _ops().gelu(hidden_states, self.approximate_tanh)
        return self.dense_4h_to_h(hidden_states)


class FlashNeoXLayer(nn.Module):
    def __init__(self, layer_id, config, weights):
        super().__init__()
        prefix = f"gpt_neox.layers.{layer_id}"
        self.use_parallel_residual = config.use_parallel_residual
        self.input_layernorm = FastLayerNorm.load(prefix=f"{prefix}.input_layernorm", weights=weights, eps=config.layer_norm_eps)
        self.post_attention_layernorm = FastLayerNorm.load(prefix=f"{prefix}.post_attention_layernorm", weights=weights,
                                                           eps=config.layer_norm_eps)
        self.attention = FlashNeoxAttention(config, prefix=f"{prefix}.attention", weights=weights)
        self.mlp = FlashMLP(config, prefix=f"{prefix}.mlp", weights=weights)
        self.process_group = weights.process_group

    def forward(self, hidden_states, residual, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q):
        if self.use_parallel_residual:  # :232-255
            ln1_hidden_states, _ = self.input_layernorm(hidden_states)
            attn_output = self.attention(ln1_hidden_states, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q)
            ln2_hidden_states, _ = self.post_attention_layernorm(hidden_states)
            mlp_output = self.mlp(ln2_hidden_states)
            intermediate = mlp_output + attn_output
            if self.process_group.size() > 1:
                torch.distributed.all_reduce(intermediate, group=self.process_group)
            return intermediate + hidden_states, None
        hidden_states, residual = self.input_layernorm(hidden_states, residual)  # :257-281
        hidden_states = self.attention(hidden_states, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q)
        hidden_states, residual = self.post_attention_layernorm(hidden_states, residual)
        return self.mlp(hidden_states), residual


class FlashGPTNeoXModel(nn.Module):
    def __init__(self, config, weights):
        super().__init__()
        self.config = config
        self.embed_in = TensorParallelEmbedding(prefix="gpt_neox.embed_in", weights=weights)
        self.layers = nn.ModuleList([FlashNeoXLayer(layer_id, config, weights) for layer_id in range(config.num_hidden_layers)])
        self.final_layer_norm = FastLayerNorm.load(prefix="gpt_neox.final_layer_norm", weights=weights, eps=config.layer_norm_eps)
        self.head_size = self.layers[0].attention.head_size
        self.num_heads = self.layers[0].attention.num_heads
        self.num_key_value_heads = self.num_heads
        self.kv_cache_manager: Optional[PagedKVCacheManager] = None

    def forward(self, input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds: Optional[torch.Tensor] = None,
                past_key_values: Optional[PagedKVState] = None, pre_allocate_past_size: Optional[int] = None):
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if past_key_values is None:
            raise ValueError("past_key_values must be the batch's PagedKVState (allocate it with kv_cache_manager)")
        hidden_states = inputs_embeds if inputs_embeds is not None else self.embed_in(input_ids)
        # fp16 cos / sin tables cached by position (utils/layers.py:436-464); the kernel gathers rows by position_ids
        rot = self.layers[0].attention.rotary_emb
        cos, sin = rot.tables(max(int(max_s), 1), hidden_states.dtype, hidden_states.device)
        residual = None
        mgr = self.kv_cache_manager
        for i, layer in enumerate(self.layers):
            k_pool, v_pool = mgr.layer_pools(i)
            hidden_states, residual = layer(hidden_states, residual, cos, sin, position_ids, cu_seqlens, max_s, past_key_values,
                                            k_pool, v_pool, cu_seqlens_q)
        hidden_states, _ = self.final_layer_norm(hidden_states, residual)
        return hidden_states, past_key_values


class FlashGPTNeoXForCausalLM(PythonFusedGreedy, nn.Module):
    def __init__(self, config, weights):
        super().__init__()
        self.config = config
        self.process_group = weights.process_group
        self.device = torch.device(weights.device)
        self.gpt_neox = FlashGPTNeoXModel(config, weights)
        self.embed_out = TensorParallelHead.load(config, prefix="embed_out", weights=weights)
        self.max_positions = int(getattr(config, "max_position_embeddings", 2048) or 2048)

    # the attributes FlashCausalLM / the server read on a flash model
    @property
    def model(self):
        return self.gpt_neox

    @property
    def lm_head(self):
        return self.embed_out

    @property
    def kv_cache_manager(self):
        return self.gpt_neox.kv_cache_manager

    @kv_cache_manager.setter
    def kv_cache_manager(self, mgr):
        self.gpt_neox.kv_cache_manager = mgr

    def get_input_embeddings(self) -> nn.Module:
        return self.gpt_neox.embed_in

    # fused greedy decode (make_step / run_step): python_step.PythonFusedGreedy, off by default

    def forward(self, input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds: Optional[torch.Tensor] = None,
                past_key_values: Optional[PagedKVState] = None, pre_allocate_past_size: Optional[int] = None,
                lm_head_indices: Optional[torch.Tensor] = None):
        hidden_states, present = self.gpt_neox(input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds,
                                               past_key_values, pre_allocate_past_size)
        if lm_head_indices is not None:
            hidden_states = hidden_states[lm_head_indices]
        logits = self.embed_out(hidden_states)
        return logits, present
