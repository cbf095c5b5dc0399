"""Flash Santacoder / StarCoder (gpt_bigcode) on the B200 kernels: multi-query attention — ONE key / value head shared by
every query head and replicated on every tensor-parallel rank —, learned position embeddings, LayerNorm, GELU, tied head.

Mirrors /root/reference/server/text_generation_server/models/custom_modeling/flash_santacoder_modeling.py: `load_multi_mqa`
(:19-159: this rank's block of query rows followed by the shared 2 * head_size key / value rows, for fp16 and GPTQ
checkpoints), `load_col` / `load_row` (:162-190), `FlashMQAttention` (:193-268), `MLP` (:271-300), `Block` (:303-347),
`FlashSantacoderModel` (:350-432), `FlashSantacoderForCausalLM` (:435-478) — same class names, constructor arguments and
`forward` signature.  Differences, as for the other flash families here: `past_key_values` is the batch's `PagedKVState`
(a block-table KV pool with a single head per block) instead of a contiguous `[layers, tokens, 2, 1, d]` tensor, and every op
goes through the C ABI.  There is no rotation in this family: the fused RoPE + KV-append kernel is called with a 2-wide
identity rotation (cos = 1, sin = 0, exact in fp16), so it only appends K / V to the pool.  The decode-attention kernel
shares a KV head between at most 16 query heads per launch; wider models (StarCoder: 48 heads) are served 16 heads at a time.

EXPERIMENTAL: composed of GPU-validated kernels and pinned on CPU against the reference's own module graph through
oracle/santacoder.py, but this file itself has not run on a GPU yet (tests/test_gpu_santacoder.py is opt-in).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed
from torch import nn

from ...utils import _ops
from ...utils.flash_attn import PagedKVLayer, attention
from ...utils.layers import (FastLayerNorm, TensorParallelColumnLinear, TensorParallelEmbedding, TensorParallelHead,
                             TensorParallelRowLinear, get_linear)
from ...utils.paged import PagedKVCacheManager, PagedKVState
from .python_step import PythonFusedGreedy

MAX_GROUP = 16  # query heads per KV head in one decode-attention launch (csrc/attn_decode.cu)


def _q_block_then_kv(view, dim: int, kv_width: int, world: int, rank: int):
    """This rank's share of the leading (query) part of `dim`, followed by the trailing `kv_width` entries every rank keeps."""
    size = view.get_shape()[dim]
    q_total = size - kv_width
    assert q_total % world == 0, f"{q_total} query columns do not split over {world} shards"
    block = q_total // world
    lo, hi = rank * block, (rank + 1) * block
    if dim == 0:
        return torch.cat([view[lo:hi], view[size - kv_width:]], dim=0)
    return torch.cat([view[:, lo:hi], view[:, size - kv_width:]], dim=1)


def load_multi_mqa(config, prefix: str, weights, bias: bool, head_size, num_heads, hidden_size):
    """The fused `c_attn` projection of this rank: [num_heads * head_size query rows | head_size key rows | head_size value rows]."""
    world, rank = weights.process_group.size(), weights.process_group.rank()
    if not any("c_attn" in k for k in weights.routing.keys()):
        raise NotImplementedError("checkpoints with separate q_attn / kv_attn tensors are not supported")
    kv = 2 * head_size
    if config.quantize == "gptq":
        if config.transpose:
            raise NotImplementedError("Gptq loading with santacoder is not implemented")  # as the reference (:85-86)
        # GPTQ tensors are [in, out]: the output dimension is dim 1; qzeros packs 8 outputs per int32
        assert kv % 8 == 0
        qweight = _q_block_then_kv(weights._get_slice(f"{prefix}.c_attn.qweight"), 1, kv, world, rank).to(weights.device)
        scales = _q_block_then_kv(weights._get_slice(f"{prefix}.c_attn.scales"), 1, kv, world, rank).to(weights.device)
        qzeros = _q_block_then_kv(weights._get_slice(f"{prefix}.c_attn.qzeros"), 1, kv // 8, world, rank).to(weights.device)
        g_idx = weights.get_tensor(f"{prefix}.c_attn.g_idx")
        bits, groupsize = weights._get_gptq_params()
        weight = (qweight, qzeros, scales, g_idx, bits, groupsize, True)
    else:
        view = weights._get_slice(f"{prefix}.c_attn.weight")
        if config.transpose:  # GPT2-architecture checkpoints store Conv1D weights [in, out]
            weight = _q_block_then_kv(view, 1, kv, world, rank).T
        else:
            weight = _q_block_then_kv(view, 0, kv, world, rank)
        weight = weight.to(dtype=weights.dtype).to(device=weights.device).contiguous()
        assert list(weight.shape) == [(num_heads + 2) * head_size, hidden_size], \
            f"{list(weight.shape)} != {[(num_heads + 2) * head_size, hidden_size]}"
    b = None
    if bias:
        b = _q_block_then_kv(weights._get_slice(f"{prefix}.c_attn.bias"), 0, kv, world, rank)
        b = b.to(dtype=weights.dtype).to(device=weights.device)
        assert list(b.shape) == [(num_heads + 2) * head_size]
    return TensorParallelColumnLinear(get_linear(weight, b, config.quantize))


def load_col(config, prefix: str, weights, bias: bool):
    if config.transpose:
        weight = weights.get_sharded(f"{prefix}.weight", dim=1).T.contiguous()
    else:
        weight = weights.get_multi_weights_col([prefix], quantize=config.quantize, dim=0)
    b = weights.get_sharded(f"{prefix}.bias", dim=0) if bias else None
    return TensorParallelColumnLinear(get_linear(weight, b, config.quantize))


def load_row(config, prefix: str, weights, bias: bool):
    if config.transpose:
        weight = weights.get_sharded(f"{prefix}.weight", dim=0).T.contiguous()
    else:
        weight = weights.get_multi_weights_row(prefix, quantize=config.quantize)
    # the bias is added once: on the first rank, before the all-reduce (:183-187)
    b = weights.get_tensor(f"{prefix}.bias") if bias and weights.process_group.rank() == 0 else None
    return TensorParallelRowLinear(get_linear(weight, b, config.quantize), process_group=weights.process_group)


class FlashMQAttention(nn.Module):
    def __init__(self, prefix, config, weights):
        super().__init__()
        num_heads = config.num_attention_heads
        self.hidden_size = config.hidden_size
        self.head_size = self.hidden_size // num_heads
        if num_heads % weights.process_group.size() != 0:
            raise ValueError(f"`num_heads` must be divisible by `num_shards` (got `num_heads`: {num_heads} "
                             f"and `num_shards`: {weights.process_group.size()}")
        self.num_heads = num_heads // weights.process_group.size()
        self.softmax_scale = self.head_size ** (-0.5)
        self.c_attn = load_multi_mqa(config, prefix=prefix, weights=weights, bias=True, head_size=self.head_size,
                                     hidden_size=self.hidden_size, num_heads=self.num_heads)
        self.c_proj = load_row(config, prefix=f"{prefix}.c_proj", weights=weights, bias=True)

    def forward(self, hidden_states, identity_cos, identity_sin, position_ids, cu_seqlens, max_s, kv: PagedKVState, k_pool, v_pool,
                cu_seqlens_q):
        h, d = self.num_heads, self.head_size
        qkv = self.c_attn(hidden_states)  # [T, (h + 2) * d] = [q heads | k | v]
        # KV append (:236 / :250): the RoPE + KV-write kernel with an identity rotation of the first two elements
        _ops().rope_kv_write_paged(qkv, identity_cos, identity_sin, position_ids, kv.slot_mapping, k_pool, v_pool, h, 1, d, rotary_dim=2)
        query = qkv[:, :h * d].unflatten(1, (h, d))
        if cu_seqlens_q is None:  # prefill (:233-246)
            key = qkv[:, h * d:(h + 1) * d].unflatten(1, (1, d))
            value = qkv[:, (h + 1) * d:].unflatten(1, (1, d))
            attn_output = attention(query, key, value, cu_seqlens, max_s, self.softmax_scale)
        else:  # decode (:248-265): one query token per sequence over the paged cache incl. the token just written
            layer = PagedKVLayer(k_pool, v_pool, kv.block_table, kv.context_lens, int(max_s))
            attn_output = torch.empty(qkv.shape[0], h, d, dtype=qkv.dtype, device=qkv.device)
            for g0 in range(0, h, MAX_GROUP):
                g1 = min(h, g0 + MAX_GROUP)
                attention(query[:, g0:g1], layer, None, cu_seqlens, max_s, self.softmax_scale, cu_seqlens_q, 1, False,
                          out=attn_output[:, g0:g1])
        return self.c_proj(attn_output.reshape(-1, h * d))


class MLP(nn.Module):
    def __init__(self, prefix, config, weights):
        super().__init__()
        act = config.activation_function
        if "gelu" not in act:
            raise NotImplementedError(f"activation_function {act!r}: only the GELU variants of flash_santacoder_modeling.py:259-270 are built")
        self.approximate_tanh = act in ["gelu_fast", "gelu_pytorch_tanh"]
        self.c_fc = load_col(config, prefix=f"{prefix}.c_fc", weights=weights, bias=True)
        self.c_proj = load_row(config, prefix=f"{prefix}.c_proj", weights=weights, bias=True)

    def forward(self, hidden_states):
        hidden_states = self.c_fc(hidden_states)
        hidden_states = _ops().gelu(hidden_states, self.approximate_tanh)
        return self.c_proj(hidden_states)


class Block(nn.Module):
    def __init__(self, layer_id, config, weights):
        super().__init__()
        prefix = f"transformer.h.{layer_id}"
        self.ln_1 = FastLayerNorm.load(prefix=f"{prefix}.ln_1", weights=weights, eps=config.layer_norm_epsilon)
        self.ln_2 = FastLayerNorm.load(prefix=f"{prefix}.ln_2", weights=weights, eps=config.layer_norm_epsilon)
        self.attn = FlashMQAttention(prefix=f"{prefix}.attn", config=config, weights=weights)
        self.mlp = MLP(prefix=f"{prefix}.mlp", config=config, weights=weights)

    def forward(self, hidden_states, residual, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q):
        hidden_states, residual = self.ln_1(hidden_states, residual)
        hidden_states = self.attn(hidden_states, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q)
        hidden_states, residual = self.ln_2(hidden_states, residual)
        return self.mlp(hidden_states), residual


class FlashSantacoderModel(nn.Module):
    def __init__(self, config, weights):
        super().__init__()
        self.config = config
        self.process_group = weights.process_group
        self.wte = TensorParallelEmbedding(prefix="transformer.wte", weights=weights, reduce=False)
        self.wpe = TensorParallelEmbedding(prefix="transformer.wpe", weights=weights, reduce=False)
        self.h = nn.ModuleList([Block(layer_id, config, weights) for layer_id in range(config.num_hidden_layers)])
        self.ln_f = FastLayerNorm.load(prefix="transformer.ln_f", weights=weights, eps=config.layer_norm_epsilon)
        self.head_size = self.h[0].attn.head_size
        self.num_heads = self.h[0].attn.num_heads
        self.num_key_value_heads = 1
        self.kv_cache_manager: Optional[PagedKVCacheManager] = None
        self._identity = None

    def _identity_rotation(self, n_positions: int, dtype, device):
        """[n_positions, 1] cos = 1 / sin = 0 tables: x1 * 1 - x2 * 0 and x1 * 0 + x2 * 1 are exact in fp16."""
        if self._identity is None or self._identity[0].shape[0] < n_positions or self._identity[0].device != device:
            self._identity = (torch.ones(n_positions, 1, dtype=dtype, device=device), torch.zeros(n_positions, 1, dtype=dtype, device=device))
        return self._identity

    def forward(self, input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds: Optional[torch.Tensor] = None,
                past_key_values: Optional[PagedKVState] = None, pre_allocate_past_size: Optional[int] = None):
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if past_key_values is None:
            raise ValueError("past_key_values must be the batch's PagedKVState (allocate it with kv_cache_manager)")
        if inputs_embeds is not None:
            hidden_states = inputs_embeds + self.wpe(position_ids)
        else:
            hidden_states = self.wte(input_ids) + self.wpe(position_ids)
        if self.process_group.size() > 1:  # both lookups are rank-partial: one all-reduce for the sum (:388-389)
            torch.distributed.all_reduce(hidden_states, group=self.process_group)
        n_positions = max(int(max_s), int(getattr(self.config, "n_positions", 0) or getattr(self.config, "max_position_embeddings", 0) or 0), 1)
        cos, sin = self._identity_rotation(n_positions, hidden_states.dtype, hidden_states.device)
        residual = None
        mgr = self.kv_cache_manager
        for i, layer in enumerate(self.h):
            k_pool, v_pool = mgr.layer_pools(i)
            hidden_states, residual = layer(hidden_states, residual, cos, sin, position_ids, cu_seqlens, max_s, past_key_values,
                                            k_pool, v_pool, cu_seqlens_q)
        hidden_states, _ = self.ln_f(hidden_states, residual)
        return hidden_states, past_key_values


class FlashSantacoderForCausalLM(PythonFusedGreedy, nn.Module):
    def __init__(self, config, weights):
        super().__init__()
        self.config = config
        self.process_group = weights.process_group
        self.device = torch.device(weights.device)
        self.transformer = FlashSantacoderModel(config, weights)
        self.lm_head = TensorParallelHead.load(config, prefix="transformer.wte", weights=weights)  # tied to the embedding (:447-449)
        self.max_positions = int(getattr(config, "n_positions", 0) or getattr(config, "max_position_embeddings", 2048) or 2048)

    @staticmethod
    def kv_cache_layout(config, world: int):
        """(KV heads in the whole model, ranks they are split over): one head, kept whole on every rank (:214-224)"""
        return 1, 1

    # the attributes FlashCausalLM / the server read on a flash model
    @property
    def model(self):
        return self.transformer

    @property
    def kv_cache_manager(self):
        return self.transformer.kv_cache_manager

    @kv_cache_manager.setter
    def kv_cache_manager(self, mgr):
        self.transformer.kv_cache_manager = mgr

    def get_input_embeddings(self) -> nn.Module:
        return self.transformer.wte

    def forward(self, input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds: Optional[torch.Tensor] = None,
                past_key_values: Optional[PagedKVState] = None, pre_allocate_past_size: Optional[int] = None,
                lm_head_indices: Optional[torch.Tensor] = None):
        hidden_states, present = self.transformer(input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds,
                                                  past_key_values, pre_allocate_past_size)
        if lm_head_indices is not None:
            hidden_states = hidden_states[lm_head_indices]
        logits = self.lm_head(hidden_states)
        return logits, present
