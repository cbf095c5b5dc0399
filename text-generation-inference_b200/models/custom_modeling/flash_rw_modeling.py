"""Flash Falcon / RefinedWeb on the B200 kernels: multi-query (Falcon-7B) and grouped-query (Falcon-40B) attention, full-head
rotary embedding, LayerNorm, exact GELU, attention and MLP in parallel off one or two LayerNorms.

Mirrors /root/reference/server/text_generation_server/models/custom_modeling/flash_rw_modeling.py: `load_row` (:20-33),
`RWConfig` (:36-118), `FlashRWAttention` (:121-214), `FlashRWLargeAttention` (:217-323), `FlashMLP` (:326-345), `FlashRWLayer`
(:348-433), `FlashRWLargeLayer` (:436-492), `FlashRWModel` (:499-603), `FlashRWForCausalLM` (:606-648) — same class names,
constructor arguments and `forward` signature.  Differences, as for the other flash families here: `past_key_values` is the
batch's `PagedKVState`, every op goes through the C ABI, and the grouped fused projection of the large form
([q heads of a group | k | v] per KV group, :259-266) is re-laid-out once at load to the [q heads | k heads | v heads] order
the RoPE / KV-write and attention kernels expect (like GPT-NeoX's [h, 3, d] rows).  The decode-attention kernel shares a KV
head between at most 16 query heads per launch: Falcon-40B has exactly 16 per group; Falcon-7B's 71 heads on one KV head
are served 16 at a time.  Tensor parallelism: over KV groups for the large form (:246-254); the small form is single-rank
here (the reference shards its fused [q | k | v] rows evenly, which only partitions heads correctly for world size 1).

EXPERIMENTAL: composed of GPU-validated kernels and pinned on CPU against the reference's own module graph through
oracle/falcon.py, but this file itself has not run on a GPU yet (tests/test_gpu_falcon.py is opt-in).
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

MAX_GROUP = 16  # query heads per KV head in one decode-attention launch (csrc/attn_decode.cu)
RW_MODEL_TYPES = ("falcon", "RefinedWeb", "RefinedWebModel")


class RWConfig:
    """The fields of the reference's RWConfig (:36-118) this family reads, from either spelling: the original RefinedWeb
    checkpoints (`n_head`, `n_head_kv`, `n_layer`) or transformers' FalconConfig (`num_attention_heads`, `num_kv_heads`)."""

    def __init__(self, model_type="RefinedWeb", vocab_size=250880, hidden_size=64, num_hidden_layers=None, num_attention_heads=None,
                 layer_norm_epsilon=1e-5, num_kv_heads=None, multi_query=False, alibi=False, new_decoder_architecture=None,
                 bias=False, parallel_attn=False, quantize=None, **kwargs):
        if alibi:
            raise NotImplementedError("alibi is not supported by this version of the model")
        self.model_type = model_type
        self.vocab_size = vocab_size
        self.hidden_size = kwargs.pop("n_embed", None) or hidden_size
        self.n_layer = num_hidden_layers if num_hidden_layers is not None else kwargs.pop("n_layer", 2)
        self.n_head = num_attention_heads if num_attention_heads is not None else kwargs.pop("n_head", 8)
        self.layer_norm_epsilon = layer_norm_epsilon
        self.bias, self.parallel_attn, self.multi_query, self.quantize = bias, parallel_attn, multi_query, quantize
        if num_kv_heads is not None:
            self.n_head_kv = num_kv_heads
        else:
            self.n_head_kv = kwargs.pop("n_head_kv", None) or (1 if multi_query else self.n_head)
        self.new_decoder_architecture = (model_type == "RefinedWeb") if new_decoder_architecture is None else new_decoder_architecture
        for k, v in kwargs.items():
            setattr(self, k, v)

    num_hidden_layers = property(lambda self: self.n_layer)
    num_attention_heads = property(lambda self: self.n_head)

    @classmethod
    def of(cls, config) -> "RWConfig":
        """Normalises whatever config object the engine loaded."""
        if isinstance(config, cls):
            return config
        get = lambda *names, default=None: next((getattr(config, n) for n in names if getattr(config, n, None) is not None), default)  # noqa: E731
        model_type = get("model_type", default="falcon")
        large = get("new_decoder_architecture", default=(model_type == "RefinedWeb"))
        heads = get("n_head", "num_attention_heads")
        kv = get("n_head_kv", "num_kv_heads")
        if not large and get("multi_query", default=False) and not hasattr(config, "n_head_kv"):
            kv = 1  # transformers' FalconConfig keeps num_kv_heads = num_attention_heads for the multi-query 7B form
        out = cls(model_type=model_type, vocab_size=get("vocab_size"), hidden_size=get("hidden_size"),
                  num_hidden_layers=get("n_layer", "num_hidden_layers"), num_attention_heads=heads,
                  layer_norm_epsilon=get("layer_norm_epsilon", default=1e-5), num_kv_heads=kv if kv is not None else heads,
                  alibi=bool(get("alibi", default=False)), new_decoder_architecture=bool(large), bias=bool(get("bias", default=False)),
                  parallel_attn=bool(get("parallel_attn", default=False)), quantize=getattr(config, "quantize", None))
        for name in ("eos_token_id", "pad_token_id", "bos_token_id", "max_position_embeddings"):
            if hasattr(config, name):
                setattr(out, name, getattr(config, name))
        return out


def load_row(config, prefix: str, weights, bias: bool):
    """:20-33: bias on rank 0 only; with parallel_attn the layer all-reduces attention + MLP once, so the bare linear is returned"""
    weight = weights.get_multi_weights_row(prefix, quantize=config.quantize)
    b = weights.get_tensor(f"{prefix}.bias") if bias and weights.process_group.rank() == 0 else None
    linear = get_linear(weight, b, config.quantize)
    if config.parallel_attn:
        return linear
    return TensorParallelRowLinear(linear, process_group=weights.process_group)


def load_grouped_qkv(config, prefix: str, weights, groups: int, heads_per_group: int, head_size: int, hidden_size: int):
    """Large form: checkpoint rows are [kv group][q heads of the group | k | v][d]; this rank's groups are a contiguous row
    block.  Re-laid-out to [q heads | k heads | v heads]."""
    weight = weights.get_multi_weights_col([prefix], quantize=config.quantize, dim=0)
    if not isinstance(weight, torch.Tensor):
        raise NotImplementedError("GPTQ checkpoints of the grouped Falcon projection are not supported")

    def regroup(t, *tail):
        t = t.view(groups, heads_per_group + 2, head_size, *tail)
        q, k, v = t[:, :heads_per_group], t[:, heads_per_group], t[:, heads_per_group + 1]
        return torch.cat([q.reshape(-1, *tail), k.reshape(-1, *tail), v.reshape(-1, *tail)], dim=0).contiguous()
    b = regroup(weights.get_sharded(f"{prefix}.bias", dim=0)) if config.bias else None
    return TensorParallelColumnLinear(get_linear(regroup(weight, hidden_size), b, config.quantize))


def _paged_attention(module, qkv, n_heads, n_kv, cos_table, sin_table, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q):
    """Shared by both attention forms once the projection is [q heads | k heads | v heads]: in-place rotary on q and k + KV
    append (one kernel), then varlen prefill attention or paged decode attention."""
    d = module.head_size
    _ops().rope_kv_write_paged(qkv, cos_table, sin_table, position_ids, kv.slot_mapping, k_pool, v_pool, n_heads, n_kv, d)
    query = qkv[:, :n_heads * d].unflatten(1, (n_heads, d))
    if cu_seqlens_q is None:
        key = qkv[:, n_heads * d:(n_heads + n_kv) * d].unflatten(1, (n_kv, d))
        value = qkv[:, (n_heads + n_kv) * d:].unflatten(1, (n_kv, d))
        return attention(query, key, value, cu_seqlens, max_s, module.softmax_scale)
    layer = PagedKVLayer(k_pool, v_pool, kv.block_table, kv.context_lens, int(max_s))
    if n_heads // n_kv <= MAX_GROUP:
        return attention(query, layer, None, cu_seqlens, max_s, module.softmax_scale, cu_seqlens_q, 1, False)
    if n_kv != 1:
        raise NotImplementedError(f"{n_heads // n_kv} query heads per KV head with {n_kv} KV heads: more than {MAX_GROUP} per launch "
                                  "is only served for a single shared KV head")
    out = torch.empty(qkv.shape[0], n_heads, d, dtype=qkv.dtype, device=qkv.device)
    for g0 in range(0, n_heads, MAX_GROUP):
        g1 = min(n_heads, g0 + MAX_GROUP)
        attention(query[:, g0:g1], layer, None, cu_seqlens, max_s, module.softmax_scale, cu_seqlens_q, 1, False, out=out[:, g0:g1])
    return out


class FlashRWAttention(nn.Module):
    def __init__(self, config, prefix, weights):
        super().__init__()
        self.num_heads = config.n_head
        self.num_heads_kv = config.n_head_kv
        self.hidden_size = config.hidden_size
        self.head_size = self.hidden_size // self.num_heads
        self.rotary_emb = PositionRotaryEmbedding.static(dim=self.head_size, base=10000.0, device=weights.device)
        self.softmax_scale = self.head_size ** (-0.5)
        if weights.process_group.size() != 1:
            raise NotImplementedError("the multi-query Falcon form runs on one rank (its fused [q | k | v] rows do not shard evenly by head)")
        self.query_key_value = TensorParallelColumnLinear.load(config, prefix=f"{prefix}.query_key_value", weights=weights, bias=config.bias)
        self.dense = load_row(config, prefix=f"{prefix}.dense", weights=weights, bias=config.bias)

    def forward(self, hidden_states, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q):
        qkv = self.query_key_value(hidden_states)  # [T, (h + 2 kv) d] = [q | k | v] already (:156-163)
        out = _paged_attention(self, qkv, self.num_heads, self.num_heads_kv, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool,
                               cu_seqlens_q)
        return self.dense(out.reshape(-1, self.num_heads * self.head_size))


class FlashRWLargeAttention(nn.Module):
    def __init__(self, config, prefix, weights):
        super().__init__()
        hidden_size, num_heads, num_groups = config.hidden_size, config.n_head, config.n_head_kv
        self.hidden_size = hidden_size
        self.head_size = hidden_size // num_heads
        self.num_heads = num_heads // num_groups  # query heads per KV group
        self.rotary_emb = PositionRotaryEmbedding.static(self.head_size, base=10000.0, device=weights.device)
        self.softmax_scale = self.head_size ** (-0.5)
        world = weights.process_group.size()
        if world > num_groups:
            raise NotImplementedError("Tensor Parallelism is not implemented for world_size > n groups")
        if num_groups % world != 0:
            raise NotImplementedError(f"Tensor Parallelism is not implemented for {num_groups} not divisible by {world}")
        self.num_groups = num_groups // world
        self.query_key_value = load_grouped_qkv(config, f"{prefix}.query_key_value", weights, self.num_groups, self.num_heads,
                                                self.head_size, hidden_size)
        self.dense = load_row(config, prefix=f"{prefix}.dense", weights=weights, bias=config.bias)

    def forward(self, hidden_states, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q):
        qkv = self.query_key_value(hidden_states)
        heads = self.num_groups * self.num_heads
        out = _paged_attention(self, qkv, heads, self.num_groups, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q)
        return self.dense(out.reshape(-1, heads * self.head_size))


class FlashMLP(nn.Module):
    def __init__(self, config, prefix, weights):
        super().__init__()
        self.dense_h_to_4h = TensorParallelColumnLinear.load(config, prefix=f"{prefix}.dense_h_to_4h", weights=weights, bias=config.bias)
        self.dense_4h_to_h = load_row(config, prefix=f"{prefix}.dense_4h_to_h", weights=weights, bias=config.bias)

    def forward(self, hidden_states):
        hidden_states = self.dense_h_to_4h(hidden_states)
        hidden_states = _ops().gelu(hidden_states, False)  # torch.nn.functional.gelu: the exact form (:331)
        return self.dense_4h_to_h(hidden_states)


class FlashRWLayer(nn.Module):
    def __init__(self, layer_id, config, weights):
        super().__init__()
        self.parallel_attn = config.parallel_attn
        prefix = f"transformer.h.{layer_id}"
        self.input_layernorm = FastLayerNorm.load(prefix=f"{prefix}.input_layernorm", weights=weights, eps=config.layer_norm_epsilon)
        self.self_attention = FlashRWAttention(config, prefix=f"{prefix}.self_attention", weights=weights)
        self.post_attention_layernorm = None if self.parallel_attn else FastLayerNorm.load(
            prefix=f"{prefix}.post_attention_layernorm", weights=weights, eps=config.layer_norm_epsilon)
        self.mlp = FlashMLP(config, prefix=f"{prefix}.mlp", weights=weights)
        self.process_group = weights.process_group

    def forward(self, hidden_states, residual, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q):
        if self.parallel_attn:  # :394-414
            ln_hidden_states, residual = self.input_layernorm(hidden_states, residual)
            attn_output = self.self_attention(ln_hidden_states, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q)
            intermediate = self.mlp(ln_hidden_states) + attn_output
            if self.process_group.size() > 1:
                torch.distributed.all_reduce(intermediate, group=self.process_group)
            return intermediate, residual
        hidden_states, residual = self.input_layernorm(hidden_states, residual)  # :415-433
        hidden_states = self.self_attention(hidden_states, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q)
        hidden_states, residual = self.post_attention_layernorm(hidden_states, residual)
        return self.mlp(hidden_states), residual


class FlashRWLargeLayer(nn.Module):
    def __init__(self, layer_id, config, weights):
        super().__init__()
        prefix = f"transformer.h.{layer_id}"
        self.ln_attn = FastLayerNorm.load(prefix=f"{prefix}.ln_attn", weights=weights, eps=config.layer_norm_epsilon)
        self.ln_mlp = FastLayerNorm.load(prefix=f"{prefix}.ln_mlp", weights=weights, eps=config.layer_norm_epsilon)
        self.self_attention = FlashRWLargeAttention(config, prefix=f"{prefix}.self_attention", weights=weights)
        assert config.parallel_attn, "This version doesn't support non parallel_attn"
        self.mlp = FlashMLP(config, prefix=f"{prefix}.mlp", weights=weights)
        self.process_group = weights.process_group

    def forward(self, hidden_states, residual, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q):
        ln_attn, residual = self.ln_attn(hidden_states, residual)
        ln_mlp, _ = self.ln_mlp(residual)
        attn_output = self.self_attention(ln_attn, cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q)
        intermediate = attn_output + self.mlp(ln_mlp)
        if self.process_group.size() > 1:
            torch.distributed.all_reduce(intermediate, group=self.process_group)
        return intermediate, residual


class FlashRWModel(nn.Module):
    def __init__(self, config, weights):
        super().__init__()
        self.config = config
        self.word_embeddings = TensorParallelEmbedding(prefix="transformer.word_embeddings", weights=weights)
        layer_class = FlashRWLargeLayer if config.new_decoder_architecture else FlashRWLayer
        self.h = nn.ModuleList([layer_class(layer_id, config, weights) for layer_id in range(config.num_hidden_layers)])
        self.ln_f = FastLayerNorm.load(prefix="transformer.ln_f", weights=weights, eps=config.layer_norm_epsilon)
        attn = self.h[0].self_attention
        self.head_size = attn.head_size
        if config.new_decoder_architecture:
            self.num_key_value_heads, self.num_heads = attn.num_groups, attn.num_groups * attn.num_heads
        else:
            self.num_key_value_heads, self.num_heads = attn.num_heads_kv, attn.num_heads
        self.kv_cache_manager: Optional[PagedKVCacheManager] = None

    def forward(self, input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds: Optional[torch.Tensor] = None,
                past_key_values: Optional[PagedKVState] = None, pre_allocate_past_size: Optional[int] = None):
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if past_key_values is None:
            raise ValueError("past_key_values must be the batch's PagedKVState (allocate it with kv_cache_manager)")
        hidden_states = inputs_embeds if inputs_embeds is not None else self.word_embeddings(input_ids)
        # fp16 cos / sin tables cached by position (utils/layers.py:436-464); the kernel gathers rows by position_ids
        cos, sin = self.h[0].self_attention.rotary_emb.tables(max(int(max_s), 1), hidden_states.dtype, hidden_states.device)
        residual = None
        mgr = self.kv_cache_manager
        for i, layer in enumerate(self.h):
            k_pool, v_pool = mgr.layer_pools(i)
            hidden_states, residual = layer(hidden_states, residual, cos, sin, position_ids, cu_seqlens, max_s, past_key_values,
                                            k_pool, v_pool, cu_seqlens_q)
        hidden_states, _ = self.ln_f(hidden_states, residual)
        return hidden_states, past_key_values


class FlashRWForCausalLM(PythonFusedGreedy, nn.Module):
    def __init__(self, config, weights):
        super().__init__()
        config = RWConfig.of(config)
        self.config = config
        self.process_group = weights.process_group
        self.device = torch.device(weights.device)
        self.transformer = FlashRWModel(config, weights)
        self.lm_head = TensorParallelHead.load(config, prefix="lm_head", weights=weights)
        self.max_positions = int(getattr(config, "max_position_embeddings", 2048) or 2048)

    @staticmethod
    def kv_cache_layout(config, world: int):
        """(KV heads in the whole model, ranks they are split over): KV groups shard with the ranks in the large form; the
        multi-query form keeps its head(s) whole on its single rank."""
        config = RWConfig.of(config)
        return (config.n_head_kv, world) if config.new_decoder_architecture else (config.n_head_kv, 1)

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
        return self.transformer.word_embeddings

    def forward(self, input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds: Optional[torch.Tensor] = None,
                past_key_values: Optional[PagedKVState] = None, pre_allocate_past_size: Optional[int] = None,
                lm_head_indices: Optional[torch.Tensor] = None):
        hidden_states, present = self.transformer(input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds,
                                                  past_key_values, pre_allocate_past_size)
        if lm_head_indices is not None:
            hidden_states = hidden_states[lm_head_indices]
        logits = self.lm_head(hidden_states)
        return logits, present
