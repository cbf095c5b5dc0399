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

from ...utils.layers import (FastLayerNorm, PositionRotaryEmbedding, TensorParallelColumnLinear, TensorParallelEmbedding,
                             TensorParallelHead, get_linear)
from ...utils.paged import PagedKVCacheManager, PagedKVState
from .python_step import FlashFamilyForCausalLM, gelu_mlp, paged_attention, row_parallel_linear, run_layers

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
    """:20-33; with parallel_attn the layer all-reduces attention + MLP once itself"""
    return row_parallel_linear(config, prefix, weights, bias, reduces_itself=not config.parallel_attn)


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


class FlashRWAttention(nn.Module):
    def __init__(self, config, prefix, weights):
        super().__init__()
        if weights.process_group.size() != 1:
            raise NotImplementedError("the multi-query Falcon form runs on one rank (its fused [q | k | v] rows do not shard evenly by head)")
        self.num_heads, self.num_heads_kv, self.hidden_size = config.n_head, config.n_head_kv, config.hidden_size
        self.head_size = self.hidden_size // self.num_heads
        self.softmax_scale = self.head_size ** (-0.5)
        self.rotary_emb = PositionRotaryEmbedding.static(dim=self.head_size, base=10000.0, device=weights.device)
        self.query_key_value = TensorParallelColumnLinear.load(config, prefix=f"{prefix}.query_key_value", weights=weights, bias=config.bias)
        self.dense = load_row(config, prefix=f"{prefix}.dense", weights=weights, bias=config.bias)

    def forward(self, hidden_states, cos, sin, position_ids, cu_seqlens, max_s, kv, cu_seqlens_q, k_pool, v_pool):
        # the projection is [q | k | v] already (:156-163)
        out = paged_attention(self.query_key_value(hidden_states), self.num_heads, self.num_heads_kv, self.head_size, self.softmax_scale,
                              cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q)
        return self.dense(out.reshape(-1, self.num_heads * self.head_size))


class FlashRWLargeAttention(nn.Module):
    def __init__(self, config, prefix, weights):
        super().__init__()
        world, groups = weights.process_group.size(), config.n_head_kv
        if world > groups:
            raise NotImplementedError("Tensor Parallelism is not implemented for world_size > n groups")
        if groups % world != 0:
            raise NotImplementedError(f"Tensor Parallelism is not implemented for {groups} not divisible by {world}")
        self.hidden_size = config.hidden_size
        self.head_size = self.hidden_size // config.n_head
        self.num_heads = config.n_head // groups  # query heads per KV group
        self.num_groups = groups // world          # this rank's KV groups
        self.softmax_scale = self.head_size ** (-0.5)
        self.rotary_emb = PositionRotaryEmbedding.static(self.head_size, base=10000.0, device=weights.device)
        self.query_key_value = load_grouped_qkv(config, f"{prefix}.query_key_value", weights, self.num_groups, self.num_heads,
                                                self.head_size, self.hidden_size)
        self.dense = load_row(config, prefix=f"{prefix}.dense", weights=weights, bias=config.bias)

    def forward(self, hidden_states, cos, sin, position_ids, cu_seqlens, max_s, kv, cu_seqlens_q, k_pool, v_pool):
        heads = self.num_groups * self.num_heads
        out = paged_attention(self.query_key_value(hidden_states), heads, self.num_groups, self.head_size, self.softmax_scale,
                              cos, sin, position_ids, cu_seqlens, max_s, kv, k_pool, v_pool, cu_seqlens_q)
        return self.dense(out.reshape(-1, heads * self.head_size))


class FlashMLP(nn.Module):
    def __init__(self, config, prefix, weights):
        super().__init__()
        self.dense_h_to_4h = TensorParallelColumnLinear.load(config, prefix=f"{prefix}.dense_h_to_4h", weights=weights, bias=config.bias)
        self.dense_4h_to_h = load_row(config, prefix=f"{prefix}.dense_4h_to_h", weights=weights, bias=config.bias)

    def forward(self, hidden_states):
        return gelu_mlp(self.dense_h_to_4h, self.dense_4h_to_h, hidden_states, False)  # torch.nn.functional.gelu: exact form (:331)


def _reduced(branches, process_group):
    if process_group.size() > 1:
        torch.distributed.all_reduce(branches, group=process_group)
    return branches


class FlashRWLayer(nn.Module):
    def __init__(self, layer_id, config, weights):
        super().__init__()
        prefix = f"transformer.h.{layer_id}"
        self.parallel_attn = config.parallel_attn
        self.process_group = weights.process_group
        self.input_layernorm = FastLayerNorm.load(prefix=f"{prefix}.input_layernorm", weights=weights, eps=config.layer_norm_epsilon)
        self.post_attention_layernorm = None if self.parallel_attn else FastLayerNorm.load(
            prefix=f"{prefix}.post_attention_layernorm", weights=weights, eps=config.layer_norm_epsilon)
        self.self_attention = FlashRWAttention(config, prefix=f"{prefix}.self_attention", weights=weights)
        self.mlp = FlashMLP(config, prefix=f"{prefix}.mlp", weights=weights)

    def forward(self, hidden_states, residual, *attention_args):
        normed, residual = self.input_layernorm(hidden_states, residual)
        if self.parallel_attn:  # :394-414: one LayerNorm feeds both branches; their fp16 sum is all-reduced once
            return _reduced(self.mlp(normed) + self.self_attention(normed, *attention_args), self.process_group), residual
        normed, residual = self.post_attention_layernorm(self.self_attention(normed, *attention_args), residual)  # :415-433
        return self.mlp(normed), residual


class FlashRWLargeLayer(nn.Module):
    def __init__(self, layer_id, config, weights):
        super().__init__()
        assert config.parallel_attn, "This version doesn't support non parallel_attn"
        prefix = f"transformer.h.{layer_id}"
        self.process_group = weights.process_group
        for name in ("ln_attn", "ln_mlp"):
            setattr(self, name, FastLayerNorm.load(prefix=f"{prefix}.{name}", weights=weights, eps=config.layer_norm_epsilon))
        self.self_attention = FlashRWLargeAttention(config, prefix=f"{prefix}.self_attention", weights=weights)
        self.mlp = FlashMLP(config, prefix=f"{prefix}.mlp", weights=weights)

    def forward(self, hidden_states, residual, *attention_args):
        for_attention, residual = self.ln_attn(hidden_states, residual)  # the residual add happens here ...
        for_mlp, _ = self.ln_mlp(residual)                               # ... so the MLP's norm reads the updated stream (:469-470)
        return _reduced(self.self_attention(for_attention, *attention_args) + self.mlp(for_mlp), self.process_group), residual


class FlashRWModel(nn.Module):
    def __init__(self, config, weights):
        super().__init__()
        self.config = config
        self.word_embeddings = TensorParallelEmbedding(prefix="transformer.word_embeddings", weights=weights)
        layer_class = FlashRWLargeLayer if config.new_decoder_architecture else FlashRWLayer
        self.h = nn.ModuleList(layer_class(n, config, weights) for n in range(config.num_hidden_layers))
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
        hidden_states = self.word_embeddings(input_ids) if inputs_embeds is None else inputs_embeds
        # fp16 cos / sin tables cached by position (utils/layers.py:436-464); the kernel gathers rows by position_ids
        cos, sin = self.h[0].self_attention.rotary_emb.tables(max(int(max_s), 1), hidden_states.dtype, hidden_states.device)
        hidden_states, residual = run_layers(self.h, self.kv_cache_manager, hidden_states, cos, sin, position_ids, cu_seqlens, max_s,
                                             past_key_values, cu_seqlens_q)
        return self.ln_f(hidden_states, residual)[0], past_key_values


class FlashRWForCausalLM(FlashFamilyForCausalLM):
    def __init__(self, config, weights):
        super().__init__()
        config = RWConfig.of(config)
        self._init_outer(config, weights)
        self.transformer = FlashRWModel(config, weights)
        self.lm_head = TensorParallelHead.load(config, prefix="lm_head", weights=weights)

    @staticmethod
    def kv_cache_layout(config, world: int):
        """(KV heads in the whole model, ranks they are split over): KV groups shard with the ranks in the large form; the
        multi-query form keeps its head(s) whole on its single rank."""
        config = RWConfig.of(config)
        return (config.n_head_kv, world) if config.new_decoder_architecture else (config.n_head_kv, 1)

    @property
    def model(self):
        return self.transformer

    def get_input_embeddings(self) -> nn.Module:
        return self.transformer.word_embeddings
