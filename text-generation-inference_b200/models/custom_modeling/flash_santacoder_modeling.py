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

from ...utils.layers import (FastLayerNorm, TensorParallelColumnLinear, TensorParallelEmbedding, TensorParallelHead,
                             TensorParallelRowLinear, get_linear)
from ...utils.paged import PagedKVCacheManager, PagedKVState
from .python_step import FlashFamilyForCausalLM, gelu_is_tanh, gelu_mlp, paged_attention, row_parallel_linear, run_layers


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
    """:173-190: always all-reduces its own output; GPT2-style checkpoints keep the weight transposed"""
    if not config.transpose:
        return row_parallel_linear(config, prefix, weights, bias, reduces_itself=True)
    weight = weights.get_sharded(f"{prefix}.weight", dim=0).T.contiguous()
    b = weights.get_tensor(f"{prefix}.bias") if bias and weights.process_group.rank() == 0 else None
    return TensorParallelRowLinear(get_linear(weight, b, config.quantize), process_group=weights.process_group)


class FlashMQAttention(nn.Module):
    def __init__(self, prefix, config, weights):
        super().__init__()
        world = weights.process_group.size()
        total_heads = config.num_attention_heads
        if total_heads % world != 0:
            raise ValueError(f"`num_heads` must be divisible by `num_shards` (got `num_heads`: {total_heads} and `num_shards`: {world}")
        self.hidden_size = config.hidden_size
        self.num_heads = total_heads // world
        self.head_size = self.hidden_size // total_heads
        self.softmax_scale = self.head_size ** (-0.5)
        self.c_attn = load_multi_mqa(config, prefix=prefix, weights=weights, bias=True, head_size=self.head_size,
                                     hidden_size=self.hidden_size, num_heads=self.num_heads)
        self.c_proj = load_row(config, prefix=f"{prefix}.c_proj", weights=weights, bias=True)

    def forward(self, hidden_states, identity_cos, identity_sin, position_ids, cu_seqlens, max_s, kv: PagedKVState, cu_seqlens_q,
                k_pool, v_pool):
        # c_attn gives [q heads | k | v] (:214-224).  KV placement (:236 / :250) is the RoPE + KV-write kernel with a 2-wide
        # identity rotation; the single KV head is shared by every query head, 16 of them per decode launch
        out = paged_attention(self.c_attn(hidden_states), self.num_heads, 1, self.head_size, self.softmax_scale, identity_cos,
                              identity_sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q, rotary_dim=2)
        return self.c_proj(out.reshape(-1, self.num_heads * self.head_size))


class MLP(nn.Module):
    def __init__(self, prefix, config, weights):
        super().__init__()
        self.approximate_tanh = gelu_is_tanh(config.activation_function, "flash_santacoder_modeling.py:259-270")
        self.c_fc = load_col(config, prefix=f"{prefix}.c_fc", weights=weights, bias=True)
        self.c_proj = load_row(config, prefix=f"{prefix}.c_proj", weights=weights, bias=True)

    def forward(self, hidden_states):
        return gelu_mlp(self.c_fc, self.c_proj, hidden_states, self.approximate_tanh)


class Block(nn.Module):
    def __init__(self, layer_id, config, weights):
        super().__init__()
        prefix = f"transformer.h.{layer_id}"
        for name in ("ln_1", "ln_2"):
            setattr(self, name, FastLayerNorm.load(prefix=f"{prefix}.{name}", weights=weights, eps=config.layer_norm_epsilon))
        self.attn = FlashMQAttention(prefix=f"{prefix}.attn", config=config, weights=weights)
        self.mlp = MLP(prefix=f"{prefix}.mlp", config=config, weights=weights)

    def forward(self, hidden_states, residual, *attention_args):
        normed, residual = self.ln_1(hidden_states, residual)
        normed, residual = self.ln_2(self.attn(normed, *attention_args), residual)
        return self.mlp(normed), residual


class FlashSantacoderModel(nn.Module):
    def __init__(self, config, weights):
        super().__init__()
        self.config = config
        self.process_group = weights.process_group
        # both tables are vocab- / position-sharded without their own all-reduce: the sum is reduced once (:388-389)
        self.wte = TensorParallelEmbedding(prefix="transformer.wte", weights=weights, reduce=False)
        self.wpe = TensorParallelEmbedding(prefix="transformer.wpe", weights=weights, reduce=False)
        self.h = nn.ModuleList(Block(n, config, weights) for n in range(config.num_hidden_layers))
        self.ln_f = FastLayerNorm.load(prefix="transformer.ln_f", weights=weights, eps=config.layer_norm_epsilon)
        self.head_size, self.num_heads, self.num_key_value_heads = self.h[0].attn.head_size, self.h[0].attn.num_heads, 1
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
        tokens = self.wte(input_ids) if inputs_embeds is None else inputs_embeds
        hidden_states = tokens + self.wpe(position_ids)
        if self.process_group.size() > 1:
            torch.distributed.all_reduce(hidden_states, group=self.process_group)
        table_rows = max(int(max_s), int(getattr(self.config, "n_positions", 0) or getattr(self.config, "max_position_embeddings", 0) or 0), 1)
        cos, sin = self._identity_rotation(table_rows, hidden_states.dtype, hidden_states.device)
        hidden_states, residual = run_layers(self.h, self.kv_cache_manager, hidden_states, cos, sin, position_ids, cu_seqlens, max_s,
                                             past_key_values, cu_seqlens_q)
        return self.ln_f(hidden_states, residual)[0], past_key_values


class FlashSantacoderForCausalLM(FlashFamilyForCausalLM):
    def __init__(self, config, weights):
        super().__init__()
        self._init_outer(config, weights)
        self.transformer = FlashSantacoderModel(config, weights)
        self.lm_head = TensorParallelHead.load(config, prefix="transformer.wte", weights=weights)  # tied to the embedding (:447-449)

    @staticmethod
    def kv_cache_layout(config, world: int):
        """(KV heads in the whole model, ranks they are split over): one head, kept whole on every rank (:214-224)"""
        return 1, 1

    @property
    def model(self):
        return self.transformer

    def get_input_embeddings(self) -> nn.Module:
        return self.transformer.wte
