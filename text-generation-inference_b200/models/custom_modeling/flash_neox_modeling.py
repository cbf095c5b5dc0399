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

from ...utils.layers import (FastLayerNorm, PositionRotaryEmbedding, TensorParallelColumnLinear, TensorParallelEmbedding,
                             TensorParallelHead, get_linear)
from ...utils.paged import PagedKVCacheManager, PagedKVState
from .python_step import FlashFamilyForCausalLM, gelu_is_tanh, gelu_mlp, paged_attention, row_parallel_linear, run_layers


def load_row(config, prefix: str, weights, bias: bool):
    """flash_neox_modeling.py:40-54; with the parallel residual the layer all-reduces attention + MLP once itself."""
    return row_parallel_linear(config, prefix, weights, bias, reduces_itself=not config.use_parallel_residual)


def load_qkv(config, prefix: str, weights, num_heads, head_size, hidden_size):
    """flash_neox_modeling.py:57-80: checkpoint rows come head-interleaved [h, 3, d]; fp16 weights are re-laid-out to [3, h, d] so
    that the product is [q heads | k heads | v heads] (what the RoPE / KV-write and attention kernels expect).  GPTQ tensors are
    left as they are, as in the reference."""
    def by_kind(t, *tail):
        return t.view(num_heads, 3, head_size, *tail).transpose(0, 1).reshape(3 * num_heads * head_size, *tail).contiguous()
    weight = weights.get_multi_weights_col([prefix], quantize=config.quantize, dim=0)
    if torch.is_tensor(weight):
        weight = by_kind(weight, hidden_size)
    linear = get_linear(weight, by_kind(weights.get_sharded(f"{prefix}.bias", dim=0)), config.quantize)
    return linear if config.use_parallel_residual else TensorParallelColumnLinear(linear)


class FlashNeoxAttention(nn.Module):
    def __init__(self, config, prefix, weights):
        super().__init__()
        world = weights.process_group.size()
        total_heads = config.num_attention_heads
        if total_heads % world != 0:
            raise ValueError(f"`num_heads` must be divisible by `num_shards` (got `num_heads`: {total_heads} and `num_shards`: {world}")
        self.hidden_size = config.hidden_size
        self.num_heads = total_heads // world
        self.head_size = self.hidden_size // total_heads
        self.softmax_scale = self.head_size ** (-0.5)
        self.rotary_emb = PositionRotaryEmbedding.load(prefix=f"{prefix}.rotary_emb", weights=weights)
        self.rotary_dim = 2 * self.rotary_emb.inv_freq.shape[0]  # utils/layers.py:467-469: rotary_dim = cos.shape[-1] * 2
        self.query_key_value = load_qkv(config, f"{prefix}.query_key_value", weights, self.num_heads, self.head_size, self.hidden_size)
        self.dense = load_row(config, f"{prefix}.dense", weights, bias=True)

    def forward(self, hidden_states, cos, sin, position_ids, cu_seqlens, max_s, kv: PagedKVState, cu_seqlens_q, k_pool, v_pool):
        # rotary on q and k (:127-129), KV placement (:137 / :150) and attention (:133-166) on the [q | k | v] product
        out = paged_attention(self.query_key_value(hidden_states), self.num_heads, self.num_heads, self.head_size, self.softmax_scale,
                              cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q, rotary_dim=self.rotary_dim)
        return self.dense(out.reshape(-1, self.num_heads * self.head_size))


class FlashMLP(nn.Module):
    def __init__(self, config, prefix, weights):
        super().__init__()
        self.approximate_tanh = gelu_is_tanh(config.hidden_act, "flash_neox_modeling.py:186-196")
        self.dense_h_to_4h = TensorParallelColumnLinear.load(config, prefix=f"{prefix}.dense_h_to_4h", weights=weights, bias=True)
        self.dense_4h_to_h = load_row(config, f"{prefix}.dense_4h_to_h", weights, bias=True)

    def forward(self, hidden_states):
        return gelu_mlp(self.dense_h_to_4h, self.dense_4h_to_h, hidden_states, self.approximate_tanh)


class FlashNeoXLayer(nn.Module):
    def __init__(self, layer_id, config, weights):
        super().__init__()
        prefix = f"gpt_neox.layers.{layer_id}"
        self.use_parallel_residual = config.use_parallel_residual
        self.process_group = weights.process_group
        for name in ("input_layernorm", "post_attention_layernorm"):
            setattr(self, name, FastLayerNorm.load(prefix=f"{prefix}.{name}", weights=weights, eps=config.layer_norm_eps))
        self.attention = FlashNeoxAttention(config, prefix=f"{prefix}.attention", weights=weights)
        self.mlp = FlashMLP(config, prefix=f"{prefix}.mlp", weights=weights)

    def forward(self, hidden_states, residual, *attention_args):
        if not self.use_parallel_residual:  # :257-281: x -> ln -> attention -> ln -> mlp, residual carried by the fused norms
            normed, residual = self.input_layernorm(hidden_states, residual)
            normed, residual = self.post_attention_layernorm(self.attention(normed, *attention_args), residual)
            return self.mlp(normed), residual
        # :232-255: both branches read the layer input; one all-reduce for their sum; fp16 adds as in the reference
        branches = self.mlp(self.post_attention_layernorm(hidden_states)[0])
        branches = branches + self.attention(self.input_layernorm(hidden_states)[0], *attention_args)
        if self.process_group.size() > 1:
            torch.distributed.all_reduce(branches, group=self.process_group)
        return branches + hidden_states, None


class FlashGPTNeoXModel(nn.Module):
    def __init__(self, config, weights):
        super().__init__()
        self.config = config
        self.embed_in = TensorParallelEmbedding(prefix="gpt_neox.embed_in", weights=weights)
        self.layers = nn.ModuleList(FlashNeoXLayer(n, config, weights) for n in range(config.num_hidden_layers))
        self.final_layer_norm = FastLayerNorm.load(prefix="gpt_neox.final_layer_norm", weights=weights, eps=config.layer_norm_eps)
        first = self.layers[0].attention
        self.head_size, self.num_heads, self.num_key_value_heads = first.head_size, first.num_heads, first.num_heads
        self.kv_cache_manager: Optional[PagedKVCacheManager] = None

    def forward(self, input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds: Optional[torch.Tensor] = None,
                past_key_values: Optional[PagedKVState] = None, pre_allocate_past_size: Optional[int] = None):
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if past_key_values is None:
            raise ValueError("past_key_values must be the batch's PagedKVState (allocate it with kv_cache_manager)")
        hidden_states = self.embed_in(input_ids) if inputs_embeds is None else inputs_embeds
        # fp16 cos / sin tables cached by position (utils/layers.py:436-464); the kernel gathers rows by position_ids
        cos, sin = self.layers[0].attention.rotary_emb.tables(max(int(max_s), 1), hidden_states.dtype, hidden_states.device)
        hidden_states, residual = run_layers(self.layers, self.kv_cache_manager, hidden_states, cos, sin, position_ids, cu_seqlens,
                                             max_s, past_key_values, cu_seqlens_q)
        return self.final_layer_norm(hidden_states, residual)[0], past_key_values


class FlashGPTNeoXForCausalLM(FlashFamilyForCausalLM):
    def __init__(self, config, weights):
        super().__init__()
        self._init_outer(config, weights)
        self.gpt_neox = FlashGPTNeoXModel(config, weights)
        self.embed_out = TensorParallelHead.load(config, prefix="embed_out", weights=weights)

    @property
    def model(self):
        return self.gpt_neox

    @property
    def lm_head(self):
        return self.embed_out

    def get_input_embeddings(self) -> nn.Module:
        return self.gpt_neox.embed_in
