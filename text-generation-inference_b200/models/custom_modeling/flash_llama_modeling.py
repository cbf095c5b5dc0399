"""FlashLlama on the B200 step runtime.

Mirrors /root/reference/server/text_generation_server/models/custom_modeling/flash_llama_modeling.py: same module
tree and constructor contract `FlashLlamaForCausalLM(config, weights)` (:499-512), same `forward` positional
signature (:514-540) and op order (:240-297, 332-335, 356-389, 425-497).  Differences, all storage-only:
  * `past_key_values` is a `PagedKVState` over the model's block pool (utils/paged.py) instead of a contiguous
    [n_layers, slots, 2, h_kv, d] tensor (:447-459): concatenate/prune never move KV, and there is no per-step
    re-pack (flash_causal_lm.py:439-447 in the reference);
  * the layer loop runs inside the C++ step runtime (csrc/llama_step.cu), one FFI call per step (per half-layer when
    tensor-parallel, with the NCCL all-reduce of utils/layers.py:318-322 in between);
  * prefill projects only `lm_head_indices` rows (the reference materialises logits for every prompt token, :539).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch
import torch.distributed
from torch import nn

from ... import _lib
from ...utils.layers import (
    FastLinear,
    LinearScalingPositionRotaryEmbedding,
    PositionRotaryEmbedding,
    TensorParallelColumnLinear,
    TensorParallelEmbedding,
    TensorParallelHead,
    TensorParallelRowLinear,
)
from ...utils.gptq.exllamav2 import Ex4bitLinearV2
from ...utils.p2p import FusedBoundary, LayerBoundaryAllReduce
from ...utils.paged import PagedKVCacheManager, PagedKVState


class LlamaRMSNorm(nn.Module):
    def __init__(self, prefix, weights, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(weights.get_tensor(f"{prefix}.weight").contiguous(), requires_grad=False)
        self.variance_epsilon = eps

    def forward(self, hidden_states, residual=None):
        from ... import ops
        return ops.rmsnorm_residual(hidden_states, residual, self.weight, self.variance_epsilon)


def _load_qkv(config, prefix: str, weights):
    """fused q/k/v column-parallel linear: each of q, k, v sharded separately then concatenated
    (flash_llama_modeling.py:155-180, 229-238)."""
    return TensorParallelColumnLinear.load_multi(
        config, prefixes=[f"{prefix}.q_proj", f"{prefix}.k_proj", f"{prefix}.v_proj"], dim=0, weights=weights,
        bias=getattr(config, "attention_bias", False))


class FlashLlamaAttention(nn.Module):
    def __init__(self, prefix: str, config, weights):
        super().__init__()
        self.num_heads = config.num_attention_heads
        self.hidden_size = config.hidden_size
        self.head_size = self.hidden_size // self.num_heads
        rope_scaling = getattr(config, "rope_scaling", None)
        if rope_scaling and "type" in rope_scaling:
            if rope_scaling["type"] != "linear":
                raise ValueError(f"rope_scaling of type {rope_scaling['type']} is not supported with FLASH_ATTENTION=True")
            self.rotary_emb = LinearScalingPositionRotaryEmbedding.static(
                dim=self.head_size, base=config.rope_theta, scaling_factor=rope_scaling.get("factor", 1.0), device=weights.device)
        else:
            self.rotary_emb = PositionRotaryEmbedding.static(dim=self.head_size, base=config.rope_theta, device=weights.device)
        self.softmax_scale = self.head_size ** -0.5
        tp = weights.process_group.size()
        if self.num_heads % tp != 0:
            raise ValueError(f"`num_heads` must be divisible by `num_shards` (got `num_heads`: {self.num_heads} and `num_shards`: {tp}")
        if config.num_key_value_heads % tp != 0:
            raise ValueError(f"`num_key_value_heads` ({config.num_key_value_heads}) must be divisible by `num_shards` ({tp})")
        self.num_heads = self.num_heads // tp
        self.num_key_value_heads = config.num_key_value_heads // tp
        self.query_key_value = _load_qkv(config, prefix, weights)
        self.o_proj = TensorParallelRowLinear.load(config, prefix=f"{prefix}.o_proj", weights=weights,
                                                   bias=getattr(config, "attention_bias", False))


class LlamaMLP(nn.Module):
    def __init__(self, prefix, config, weights):
        super().__init__()
        if getattr(config, "hidden_act", "silu") != "silu":
            raise NotImplementedError("the fused MLP kernel implements SiLU (Llama) only")
        self.gate_up_proj = TensorParallelColumnLinear.load_multi(
            config, prefixes=[f"{prefix}.gate_proj", f"{prefix}.up_proj"], weights=weights, dim=0,
            bias=getattr(config, "mlp_bias", False))
        self.down_proj = TensorParallelRowLinear.load(config, prefix=f"{prefix}.down_proj", weights=weights,
                                                      bias=getattr(config, "mlp_bias", False))
        self.intermediate_size = config.intermediate_size // weights.process_group.size()


class FlashLlamaLayer(nn.Module):
    def __init__(self, layer_id, config, weights):
        super().__init__()
        prefix = f"model.layers.{layer_id}"
        self.self_attn = FlashLlamaAttention(prefix=f"{prefix}.self_attn", config=config, weights=weights)
        self.mlp = LlamaMLP(prefix=f"{prefix}.mlp", config=config, weights=weights)
        self.input_layernorm = LlamaRMSNorm(prefix=f"{prefix}.input_layernorm", weights=weights, eps=config.rms_norm_eps)
        self.post_attention_layernorm = LlamaRMSNorm(prefix=f"{prefix}.post_attention_layernorm", weights=weights,
                                                     eps=config.rms_norm_eps)


class FlashLlamaModel(nn.Module):
    def __init__(self, config, weights):
        super().__init__()
        self.config = config
        process_group = weights.process_group
        self.tp_rank = process_group.rank()
        self.tp_world_size = process_group.size()
        self.embed_tokens = TensorParallelEmbedding(prefix="model.embed_tokens", weights=weights)
        self.layers = nn.ModuleList([FlashLlamaLayer(i, config, weights) for i in range(config.num_hidden_layers)])
        self.norm = LlamaRMSNorm(prefix="model.norm", weights=weights, eps=config.rms_norm_eps)
        self.head_size = self.layers[0].self_attn.head_size
        self.num_heads = self.layers[0].self_attn.num_heads
        self.num_key_value_heads = self.layers[0].self_attn.num_key_value_heads


def _fill_linear(dst: _lib.B200Linear, lin, gate_up: bool = False) -> None:
    dst.layout = 0
    if isinstance(lin, Ex4bitLinearV2):
        lin.post_init(layout=1 if gate_up else 0)  # gate|up record order: SiLU * up fuses into the GEMM epilogue
        dst.weight = None
        dst.qweight = lin.q_handle.data_ptr()
        dst.perm = lin.q_perm.data_ptr() if lin.q_perm is not None else None
        dst.N, dst.K, dst.groupsize = lin.outfeatures, lin.infeatures, lin.group_size
        dst.layout = lin.pack_layout
    elif isinstance(lin, FastLinear):
        dst.weight = lin.weight.data_ptr()
        dst.qweight = dst.perm = None
        dst.N, dst.K, dst.groupsize = lin.weight.shape[0], lin.weight.shape[1], 0
    else:
        raise TypeError(type(lin))
    dst.bias = lin.bias.data_ptr() if lin.bias is not None else None


class StepScratch:
    """Grow-only activation scratch for one step of T tokens (owned by the torch caching allocator)."""

    def __init__(self, model: "FlashLlamaForCausalLM"):
        self.m = model
        self.cap = 0
        self.version = 0  # bumped whenever a buffer is re-allocated: cached step structs / CUDA graphs must be rebuilt
        self.bufs = {}
        self.attn_ws = None
        self.gemm_ws = None

    def ensure(self, T: int, B: int, max_ctx: int):
        m, dev = self.m, self.m.device
        if T > self.cap:
            cap = max(T, 16)
            H, d = m.config.hidden_size, m.model.head_size
            nqkv = (m.model.num_heads + 2 * m.model.num_key_value_heads) * d
            I = m.model.layers[0].mlp.intermediate_size
            f16 = dict(dtype=torch.float16, device=dev)
            self.bufs = dict(hidden=torch.empty(cap, H, **f16), residual=torch.empty(cap, H, **f16), normed=torch.empty(cap, H, **f16),
                             qkv=torch.empty(cap, nqkv, **f16), attn_out=torch.empty(cap, m.model.num_heads * d, **f16),
                             gate_up=torch.empty(cap, 2 * I, **f16), act=torch.empty(cap, I, **f16),
                             head_in=torch.empty(cap, H, **f16))
            if m.has_act_order:  # gathered activations of act-order GPTQ linears
                self.bufs["perm_x"] = torch.empty(cap, max(K for _, K in m.linear_shapes), **f16)
            self.cap = cap
            self.version += 1
            lib = _lib.load()
            need = 0
            for (N, K) in m.linear_shapes:
                need = max(need, lib.b200_gemm_workspace_bytes_max(N, K))
            if self.gemm_ws is None or self.gemm_ws.numel() < need:
                self.gemm_ws = torch.zeros(need, dtype=torch.uint8, device=dev)
                self.version += 1
        need = _lib.load().b200_attn_decode_workspace_bytes(max(B, 1), m.model.num_heads, m.model.head_size, max(max_ctx, 1))
        if self.attn_ws is None or self.attn_ws.numel() < need:
            self.attn_ws = torch.empty(need, dtype=torch.uint8, device=dev)
            self.version += 1


class FlashLlamaForCausalLM(nn.Module):
    def __init__(self, config, weights):
        super().__init__()
        self.config = config
        self.process_group = weights.process_group
        self.device = torch.device(weights.device)
        self.model = FlashLlamaModel(config, weights)
        self.lm_head = TensorParallelHead.load(config, prefix="lm_head", weights=weights)
        self.max_positions = int(getattr(config, "max_position_embeddings", 4096) or 4096)
        self.kv_cache_manager: Optional[PagedKVCacheManager] = None
        self._cw = None
        self.scratch = StepScratch(self)
        # the layer boundary: inside the C++ step over NVLink peer memory at decode sizes (FusedBoundary), NCCL / the one-shot
        # all-reduce between the half-layer block calls otherwise (prefill); B200_P2P_ALLREDUCE=0 keeps everything on NCCL
        self._all_reduce = LayerBoundaryAllReduce(self.process_group)
        self._boundary = FusedBoundary(self.process_group, config.hidden_size)
        self.defer_splitk = __import__("os").environ.get("B200_DEFER_SPLITK", "1") != "0"
        self.prefill_paged = __import__("os").environ.get("B200_PREFILL_PAGED", "1") != "0"

    def get_input_embeddings(self) -> nn.Module:
        return self.model.embed_tokens

    def get_kv_cache_block_size(self, block_size: int) -> int:
        """elements of one block for one layer, K and V (paged_llama_modeling.py:437-441)."""
        return block_size * self.model.num_key_value_heads * self.model.head_size * 2

    # ---------------------------------------------------------------------------- C structs
    def _rope_tables(self, max_s: int):
        rot = self.model.layers[0].self_attn.rotary_emb
        n = max(self.max_positions, max_s)
        if n > rot._seq_len_cached:
            n = max(n, 2 * rot._seq_len_cached)  # grow geometrically: pointers in the C struct must stay stable
            self._cw = None
            # cached step structs and captured CUDA graphs hold the old tables' pointers: FlashCausalLM keys them on this version
            self.scratch.version += 1
        return rot.tables(n, torch.float16, self.device)

    def c_weights(self, max_s: int = 0) -> _lib.B200LlamaWeights:
        cos, sin = self._rope_tables(max_s)
        if self._cw is not None:
            return self._cw
        m = self.model
        n = len(m.layers)
        arr = (_lib.B200LlamaLayer * n)()
        shapes = []
        self.has_act_order = False
        for i, layer in enumerate(m.layers):
            arr[i].input_ln = layer.input_layernorm.weight.data_ptr()
            arr[i].post_ln = layer.post_attention_layernorm.weight.data_ptr()
            for name, lin in (("qkv", layer.self_attn.query_key_value.linear), ("o", layer.self_attn.o_proj.linear),
                              ("gate_up", layer.mlp.gate_up_proj.linear), ("down", layer.mlp.down_proj.linear)):
                dst = getattr(arr[i], name)
                _fill_linear(dst, lin, gate_up=(name == "gate_up"))
                self.has_act_order |= bool(dst.perm)
                shapes.append((dst.N, dst.K))
        head = self.lm_head.linear
        assert isinstance(head, FastLinear), "GPTQ never quantizes the head (utils/layers.py:236-237)"
        shapes.append((head.weight.shape[0], head.weight.shape[1]))
        self.linear_shapes = sorted(set(shapes))
        w = _lib.B200LlamaWeights()
        w.n_layers, w.hidden_size = n, self.config.hidden_size
        w.n_heads, w.n_kv_heads, w.head_dim = m.num_heads, m.num_key_value_heads, m.head_size
        w.tp_size, w.tp_rank = m.tp_world_size, m.tp_rank
        w.rms_eps, w.softmax_scale = float(self.config.rms_norm_eps), float(m.layers[0].self_attn.softmax_scale)
        w.layers = ctypes.cast(arr, ctypes.POINTER(_lib.B200LlamaLayer))
        w.embed = m.embed_tokens.weight.data_ptr()
        w.vocab_start, w.vocab_rows = m.embed_tokens.min_id, m.embed_tokens.weight.shape[0]
        w.final_norm = m.norm.weight.data_ptr()
        w.lm_head, w.vocab_rows_head = head.weight.data_ptr(), head.weight.shape[0]
        w.rope_cos, w.rope_sin = cos.data_ptr(), sin.data_ptr()
        self._cw_keep = (arr, cos, sin)
        self._cw = w
        return w

    def make_step(self, *, T: int, B: int, is_prefill: bool, max_s: int, input_ids, position_ids, kv: PagedKVState,
                  cu_seqlens=None, head_rows=None, logits=None, next_ids=None, inputs_embeds=None) -> _lib.B200LlamaStep:
        """Fills a B200LlamaStep over the model's scratch (sized for T) and the caller's index tensors."""
        self.c_weights(max_s)
        sc = self.scratch
        sc.ensure(T, B, max_s)
        mgr = self.kv_cache_manager
        s = _lib.B200LlamaStep()
        s.T, s.B, s.is_prefill, s.max_s = T, B, int(is_prefill), int(max_s)
        s.input_ids = input_ids.data_ptr() if input_ids is not None else None
        s.position_ids = position_ids.data_ptr()
        s.slot_mapping = kv.slot_mapping.data_ptr()
        s.cu_seqlens = cu_seqlens.data_ptr() if cu_seqlens is not None else None
        s.block_table, s.block_table_stride = kv.block_table.data_ptr(), kv.block_table.stride(0)
        s.context_lens = kv.context_lens.data_ptr()
        s.kv_pool, s.kv_layer_stride_bytes, s.kv_v_offset_bytes = mgr.pool.data_ptr(), mgr.layer_stride_bytes, mgr.v_offset_bytes
        for k in ("hidden", "residual", "normed", "qkv", "attn_out", "gate_up", "act", "head_in"):
            setattr(s, k, sc.bufs[k].data_ptr())
        s.perm_x = sc.bufs["perm_x"].data_ptr() if "perm_x" in sc.bufs else None
        if inputs_embeds is not None:
            sc.bufs["hidden"][:T].copy_(inputs_embeds)
        s.attn_ws, s.attn_ws_bytes = sc.attn_ws.data_ptr(), sc.attn_ws.numel()
        s.gemm_ws = sc.gemm_ws.data_ptr()
        if head_rows is not None:
            s.head_rows, s.n_head_rows = head_rows.data_ptr(), head_rows.shape[0]
        s.logits = logits.data_ptr()
        s.next_ids = next_ids.data_ptr() if next_ids is not None else None
        s.defer_splitk = int(self.defer_splitk)
        s.p2p_norm, s.p2p_argmax = self._boundary.norm, self._boundary.argmax
        # prefill attends through the block pool (tcgen05 + TMA, csrc/attn_prefill_paged.cu): the step's K / V are appended first,
        # so the same kernel serves prompts that follow a cached context.  B200_PREFILL_PAGED=0: the varlen HMMA kernel on fresh q/k/v.
        if is_prefill and self.prefill_paged and kv.block_table is not None and kv.block_table.numel() > 0:
            s.kv_num_blocks = mgr.pool.shape[2]
            s.max_q = int(max_s)
        return s

    @property
    def greedy_ids_in_step(self) -> bool:
        """the step can produce the greedy ids itself: single rank, or a sharded head with the (value, index) exchange"""
        return self.model.tp_world_size == 1 or (self._boundary.argmax is not None and self.lm_head.should_gather)

    def run_step(self, s: _lib.B200LlamaStep, embed: bool = True) -> None:
        """Enqueues the step on the current stream: one C call single-rank; per half-layer + NCCL when sharded."""
        lib = _lib.load()
        w = self._cw
        st = torch.cuda.current_stream().cuda_stream
        tp = self.model.tp_world_size
        if embed and (tp == 1 or (s.p2p_norm and s.T <= FusedBoundary.MAX_ROWS)):
            # one C call for the whole step; when sharded the layer boundary runs inside it over NVLink peer memory
            _lib.check(lib.b200_llama_step(ctypes.byref(w), ctypes.byref(s), st), "llama_step")
            return
        hidden = self.scratch.bufs["hidden"][:s.T]
        if embed:
            _lib.check(lib.b200_llama_embed(ctypes.byref(w), ctypes.byref(s), st), "llama_embed")
            if tp > 1:
                self._all_reduce(hidden)
        for l in range(w.n_layers):
            _lib.check(lib.b200_llama_attn_block(ctypes.byref(w), ctypes.byref(s), l, st), "llama_attn_block")
            if tp > 1:
                self._all_reduce(hidden)
            _lib.check(lib.b200_llama_mlp_block(ctypes.byref(w), ctypes.byref(s), l, st), "llama_mlp_block")
            if tp > 1:
                self._all_reduce(hidden)
        _lib.check(lib.b200_llama_head(ctypes.byref(w), ctypes.byref(s), st), "llama_head")

    # ---------------------------------------------------------------------------- reference-shaped forward
    def forward(self, input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds: Optional[torch.Tensor] = None,
                past_key_values: Optional[PagedKVState] = None, pre_allocate_past_size: Optional[int] = None,
                lm_head_indices: Optional[torch.Tensor] = None):
        """-> (logits [rows, V] fp16, present).  Prefill when cu_seqlens_q is None (slots for every token in
        past_key_values.slot_mapping); decode otherwise (one token per sequence, context_lens already include it)."""
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if past_key_values is None:
            raise ValueError("past_key_values must be the batch's PagedKVState (allocate it with kv_cache_manager)")
        T = position_ids.shape[0]
        B = past_key_values.context_lens.shape[0]
        is_prefill = cu_seqlens_q is None
        rows = lm_head_indices.shape[0] if lm_head_indices is not None else T
        V_local = self.lm_head.linear.weight.shape[0]
        logits = torch.empty(rows, V_local, dtype=torch.float16, device=self.device)
        s = self.make_step(T=T, B=B, is_prefill=is_prefill, max_s=max_s, input_ids=input_ids, position_ids=position_ids,
                           kv=past_key_values, cu_seqlens=cu_seqlens, head_rows=lm_head_indices, logits=logits,
                           inputs_embeds=inputs_embeds)
        self.run_step(s, embed=inputs_embeds is None)
        return self._gather_logits(logits), past_key_values

    def _gather_logits(self, logits: torch.Tensor) -> torch.Tensor:
        """vocab-sharded head: all-gather of this rank's [rows, V / tp] (utils/layers.py:249-269)"""
        if not self.lm_head.should_gather:
            return logits
        world = self.process_group.size()
        rows, V_local = logits.shape
        gathered = logits.new_empty(world, rows, V_local)
        torch.distributed.all_gather_into_tensor(gathered, logits, group=self.process_group)
        return gathered.permute(1, 0, 2).reshape(rows, world * V_local)
