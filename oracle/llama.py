"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement (PyTorch, fp32 arithmetic with fp16 rounding exactly where the reference rounds)
of the reference's FlashLlama decode/prefill hot path.  Each function cites the reference lines
it follows (paths relative to /root/reference/server/text_generation_server/).

The GPU kernels that the reference calls here are third-party and un-vendored (flash-attn 2.5.6
`flash_attn_2_cuda`, `dropout_layer_norm`, `rotary_emb`), and no reference test pins
their results: for those ops the status is "PARITY UNPINNED" — the arithmetic below restates their
published behaviour (fp32 statistics / fp32 softmax, fp16 I/O) and is anchored on
  (a) the reference's own call sites, and
  (b) an independent implementation: transformers.LlamaForCausalLM eager fp32 (tests/test_oracle.py).
What IS pinned by running reference code in this container (tests/golden/make_golden.py):
rotary cos/sin tables (utils/layers.py:436-464), the eager H>8192 RMSNorm branch
(flash_llama_modeling.py:114-129), TP shard slicing (utils/weights.py:79-201), GPTQ pack layout.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import gptq as ogptq

F16 = torch.float16


# --------------------------------------------------------------------------------------
# elementwise / normalisation
# --------------------------------------------------------------------------------------
def rmsnorm_residual(h: torch.Tensor, residual: Optional[torch.Tensor], gamma: torch.Tensor,
                     eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """flash_llama_modeling.py:113-152 (fused branch, H <= 8192): `dropout_add_ln_fwd(h, residual,
    gamma, ..., eps, 1.0, 0, None, False, True)` -> (normed fp16, residual_out fp16).
    x32 = h + residual in fp32; residual_out = fp16(x32); normed = fp16(x32 * rsqrt(mean(x32^2)+eps) * gamma).
    residual None -> residual_out = h (:149-150)."""
    x = h.float()
    if residual is not None:
        x = x + residual.float()
        res_out = x.to(F16)
    else:
        res_out = h
    var = x.pow(2).mean(-1, keepdim=True)
    normed = (x * torch.rsqrt(var + eps) * gamma.float()).to(F16)
    return normed, res_out


def rope_tables(head_dim: int, theta: float, max_s: int, scaling_factor: float = 1.0):
    """utils/layers.py:419-425 (inv_freq fp32), :436-451 (tables computed fp32, cast to model dtype)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    t = torch.arange(max_s, dtype=torch.float32)
    if scaling_factor != 1.0:
        t = t / scaling_factor
    freqs = torch.outer(t, inv_freq)
    return torch.cos(freqs).to(F16), torch.sin(freqs).to(F16)


def apply_rotary(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """utils/layers.py:466-472 -> rotary_emb.apply_rotary(x1, x2, cos, sin, x1, x2, False):
    half-split (NeoX style) pairs (x[j], x[j + d/2]); fp32 math, fp16 store.
    x [T, h, d]; cos/sin [T, d/2] fp16."""
    rd = cos.shape[-1]
    x1 = x[..., :rd].float()
    x2 = x[..., rd:2 * rd].float()
    c = cos.float()[:, None, :]
    s = sin.float()[:, None, :]
    o1 = (x1 * c - x2 * s).to(F16)
    o2 = (x1 * s + x2 * c).to(F16)
    out = x.clone()
    out[..., :rd] = o1
    out[..., rd:2 * rd] = o2
    return out


def silu_mul(gate_up: torch.Tensor, inter: int) -> torch.Tensor:
    """flash_llama_modeling.py:332-335: view(-1, 2, I); act(gu[:,0]) * gu[:,1]; eager fp16 ops, i.e.
    silu computed in fp32 and rounded to fp16, then an fp16 multiply."""
    gu = gate_up.view(-1, 2, inter)
    g = gu[:, 0].float()
    act = (g * torch.sigmoid(g)).to(F16)  # torch's fp16 SiLU: fp32 opmath, one rounding
    return (act.float() * gu[:, 1].float()).to(F16)


# --------------------------------------------------------------------------------------
# attention  (op boundary: utils/flash_attn.py:43-127)
# --------------------------------------------------------------------------------------
def attention_prefill(q, k, v, cu_seqlens, softmax_scale: float) -> torch.Tensor:
    """Varlen causal attention, flash_llama_modeling.py:271-278.
    q [T,h,d], k/v [T,h_kv,d] fp16 -> [T,h,d] fp16; GQA head i -> kv head i // (h/h_kv)."""
    T, h, d = q.shape
    h_kv = k.shape[1]
    rep = h // h_kv
    out = torch.empty_like(q)
    cu = [int(c) for c in cu_seqlens]
    for b in range(len(cu) - 1):
        s, e = cu[b], cu[b + 1]
        L = e - s
        if L == 0:
            continue
        qq = q[s:e].float().transpose(0, 1)  # [h, L, d]
        kk = k[s:e].float().transpose(0, 1).repeat_interleave(rep, dim=0)
        vv = v[s:e].float().transpose(0, 1).repeat_interleave(rep, dim=0)
        sc = torch.matmul(qq, kk.transpose(1, 2)) * softmax_scale
        mask = torch.ones(L, L, dtype=torch.bool).tril()
        sc = sc.masked_fill(~mask, float("-inf"))
        p = torch.softmax(sc, dim=-1)
        out[s:e] = torch.matmul(p, vv).transpose(0, 1).to(F16)
    return out


def attention_decode(q, k_list, v_list, softmax_scale: float) -> torch.Tensor:
    """Decode attention: one query token per sequence over its L cached keys *including* the token
    just written (flash_llama_modeling.py:282 precedes :285), causal=False (:293-295).
    q [B,h,d]; k_list[b], v_list[b]: [L_b, h_kv, d]."""
    B, h, d = q.shape
    out = torch.empty_like(q)
    for b in range(B):
        k = k_list[b].float()
        v = v_list[b].float()
        rep = h // k.shape[1]
        kk = k.transpose(0, 1).repeat_interleave(rep, dim=0)  # [h, L, d]
        vv = v.transpose(0, 1).repeat_interleave(rep, dim=0)
        sc = torch.einsum("hd,hld->hl", q[b].float(), kk) * softmax_scale
        p = torch.softmax(sc, dim=-1)
        out[b] = torch.einsum("hl,hld->hd", p, vv).to(F16)
    return out


# --------------------------------------------------------------------------------------
# linears
# --------------------------------------------------------------------------------------
@dataclass
class Linear:
    """fp16 `FastLinear` (utils/layers.py:88-111) or GPTQ `Ex4bitLinearV2` (utils/gptq/exllamav2.py:100-144)."""
    weight: Optional[torch.Tensor] = None  # [N, K] fp16
    qweight: Optional[torch.Tensor] = None
    qzeros: Optional[torch.Tensor] = None
    scales: Optional[torch.Tensor] = None
    g_idx: Optional[torch.Tensor] = None
    groupsize: int = 128
    bias: Optional[torch.Tensor] = None
    _wdq: Optional[torch.Tensor] = field(default=None, repr=False)

    def dense_kn(self) -> torch.Tensor:
        """fp32 copy of the (de-quantised) weight as [K, N]."""
        if self._wdq is None:
            if self.weight is not None:
                self._wdq = self.weight.to(F16).float().t().contiguous()
            else:
                self._wdq = ogptq.dequantize(self.qweight, self.qzeros, self.scales, self.g_idx,
                                             self.groupsize).float()
        return self._wdq

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        y = (x.to(F16).float() @ self.dense_kn()).to(F16)
        if self.bias is not None:
            y = y + self.bias.to(F16)
        return y


# --------------------------------------------------------------------------------------
# model
# --------------------------------------------------------------------------------------
@dataclass
class LlamaConfig:
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    vocab_size: int
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    rope_scaling_factor: float = 1.0

    @property
    def head_dim(self):
        return self.hidden_size // self.num_attention_heads


@dataclass
class LlamaLayer:
    input_ln: torch.Tensor
    post_ln: torch.Tensor
    qkv: Linear
    o: Linear
    gate_up: Linear
    down: Linear


@dataclass
class LlamaShard:
    """One tensor-parallel rank's weights (world size 1 == the whole model)."""
    cfg: LlamaConfig
    embed: torch.Tensor  # [V, H] (rank-local rows when TP; see tp_embed)
    layers: List[LlamaLayer]
    norm: torch.Tensor
    lm_head: torch.Tensor  # [V(/tp), H]
    tp: int = 1
    rank: int = 0
    vocab_start: int = 0

    @property
    def n_heads(self):
        return self.cfg.num_attention_heads // self.tp

    @property
    def n_kv(self):
        return self.cfg.num_key_value_heads // self.tp


class LlamaOracle:
    """Restates FlashLlamaForCausalLM.forward (flash_llama_modeling.py:425-540) over a ragged batch.
    KV is kept per sequence as python lists of [L, h_kv, d] tensors per layer per rank — the
    reference's contiguous `past_key_values [n_layers, slots, 2, h_kv, d]` (:447-459) and any paged
    layout are both just storage for this.
    Tensor parallel: `shards` holds tp ranks; all-reduce (utils/layers.py:318-322, 355-356) is an
    fp16 sum over ranks in rank order; lm_head all-gather (:249-269) is a concat."""

    def __init__(self, shards: List[LlamaShard]):
        self.shards = shards
        self.cfg = shards[0].cfg
        self.tp = len(shards)

    # -- collectives ---------------------------------------------------------------
    @staticmethod
    def _all_reduce(parts: List[torch.Tensor]) -> torch.Tensor:
        if len(parts) == 1:
            return parts[0]
        acc = parts[0].float()
        for p in parts[1:]:
            acc = acc + p.float()
        return acc.to(F16)

    def _embed(self, input_ids: torch.Tensor) -> torch.Tensor:
        """TensorParallelEmbedding.forward, utils/layers.py:346-357 (out-of-shard ids -> null row)."""
        parts = []
        for sh in self.shards:
            V_loc = sh.embed.shape[0]
            local = input_ids - sh.vocab_start
            ok = (local >= 0) & (local < V_loc)
            e = torch.zeros(input_ids.shape[0], sh.embed.shape[1], dtype=F16)
            e[ok] = sh.embed[local[ok]].to(F16)
            parts.append(e)
        return self._all_reduce(parts)

    # -- forward -------------------------------------------------------------------
    def forward(self, input_ids: torch.Tensor, position_ids: torch.Tensor, cu_seqlens: List[int],
                kv: Optional[list], prefill: bool, all_logits: bool = False):
        """prefill: tokens ragged by cu_seqlens; returns logits for last tokens [B,V] (or all T).
        decode: one token per sequence; `kv[rank][layer][b]` = (k [L,h_kv,d], v) gets the new token appended.
        Returns (logits fp16, kv)."""
        cfg = self.cfg
        d = cfg.head_dim
        B = len(cu_seqlens) - 1
        max_pos = int(position_ids.max()) + 1
        cos_t, sin_t = rope_tables(d, cfg.rope_theta, max_pos, cfg.rope_scaling_factor)
        cos = cos_t[position_ids]
        sin = sin_t[position_ids]
        scale = d ** -0.5
        hidden = self._embed(input_ids)
        residual = None
        if kv is None:
            kv = [[[None] * B for _ in range(cfg.num_hidden_layers)] for _ in range(self.tp)]
        for li in range(cfg.num_hidden_layers):
            normed, residual = rmsnorm_residual(hidden, residual, self.shards[0].layers[li].input_ln,
                                                cfg.rms_norm_eps)
            attn_parts = []
            for r, sh in enumerate(self.shards):
                L = sh.layers[li]
                qkv = L.qkv(normed)
                q, k, v = qkv.split([sh.n_heads * d, sh.n_kv * d, sh.n_kv * d], dim=1)
                q = apply_rotary(q.reshape(-1, sh.n_heads, d), cos, sin)
                k = apply_rotary(k.reshape(-1, sh.n_kv, d), cos, sin)
                v = v.reshape(-1, sh.n_kv, d)
                if prefill:
                    for b in range(B):
                        s, e = cu_seqlens[b], cu_seqlens[b + 1]
                        kv[r][li][b] = (k[s:e].clone(), v[s:e].clone())
                    a = attention_prefill(q, k, v, cu_seqlens, scale)
                else:
                    ks, vs = [], []
                    for b in range(B):
                        pk, pv = kv[r][li][b]
                        nk = torch.cat([pk, k[b:b + 1]], 0)
                        nv = torch.cat([pv, v[b:b + 1]], 0)
                        kv[r][li][b] = (nk, nv)
                        ks.append(nk)
                        vs.append(nv)
                    a = attention_decode(q, ks, vs, scale)
                attn_parts.append(L.o(a.reshape(-1, sh.n_heads * d)))
            attn_out = self._all_reduce(attn_parts)
            normed2, residual = rmsnorm_residual(attn_out, residual, self.shards[0].layers[li].post_ln,
                                                 cfg.rms_norm_eps)
            mlp_parts = []
            for sh in self.shards:
                L = sh.layers[li]
                gu = L.gate_up(normed2)
                mlp_parts.append(L.down(silu_mul(gu, cfg.intermediate_size // self.tp)))
            hidden = self._all_reduce(mlp_parts)
        hidden, _ = rmsnorm_residual(hidden, residual, self.shards[0].norm, cfg.rms_norm_eps)
        if prefill and not all_logits:
            last = torch.tensor([c - 1 for c in cu_seqlens[1:]], dtype=torch.long)
            hidden = hidden[last]
        logits = torch.cat([(hidden.float() @ sh.lm_head.to(F16).float().t()).to(F16) for sh in self.shards], dim=1)
        return logits, kv

    # -- greedy generate (Greedy chooser, utils/tokens.py:44-46) ----------------------
    def generate_greedy(self, prompts: List[List[int]], n_new: int, banned_token: Optional[int] = None,
                        forced: Optional[torch.Tensor] = None, keep_logits: bool = True, on_step=None):
        """Returns (tokens [B, n_new], logits list per step). Mirrors FlashCausalLM.generate_token's
        loop (models/flash_causal_lm.py:405-460): prefill then n_new-1 decode steps.
        banned_token: min_new_tokens EOS mask, `scores[idx, eos] = -inf` before the arg-max (utils/tokens.py:244-246).
        forced [B, n_new] (optional): teacher forcing - the returned tokens are still this oracle's own choices, but step t + 1 is
        fed forced[:, t], so a long trajectory produced elsewhere can be checked token by token without diverging after a tie.
        on_step(step, logits) (optional) is called with every step's logits (use it with keep_logits=False on long runs)."""
        def choose(lg):
            lg = lg.float().clone()
            if banned_token is not None:
                lg[:, banned_token] = float("-inf")
            return lg.argmax(-1)

        cu = [0]
        for p in prompts:
            cu.append(cu[-1] + len(p))
        ids = torch.tensor([t for p in prompts for t in p], dtype=torch.long)
        pos = torch.cat([torch.arange(len(p)) for p in prompts])
        logits, kv = self.forward(ids, pos, cu, None, prefill=True)
        lens = [len(p) for p in prompts]
        toks, all_logits = [], ([logits] if keep_logits else [])
        if on_step:
            on_step(0, logits)
        nxt = choose(logits)
        toks.append(nxt)
        for step in range(1, n_new):
            pos = torch.tensor(lens, dtype=torch.long)
            fed = nxt if forced is None else forced[:, step - 1].to(torch.long)
            logits, kv = self.forward(fed, pos, list(range(len(prompts) + 1)), kv, prefill=False)
            lens = [l + 1 for l in lens]
            nxt = choose(logits)
            toks.append(nxt)
            if keep_logits:
                all_logits.append(logits)
            if on_step:
                on_step(step, logits)
        return torch.stack(toks, 1), all_logits


# --------------------------------------------------------------------------------------
# synthetic checkpoints and TP slicing (utils/weights.py:79-201)
# --------------------------------------------------------------------------------------
def make_state_dict(cfg: LlamaConfig, seed: int = 1234, quantize: Optional[str] = None,
                    groupsize: int = 128, std: float = 0.02, act_order: bool = False) -> Dict[str, torch.Tensor]:
    """HF-named tensors. quantize='gptq' -> reference-format int4 tensors.
    act_order: every linear gets a shuffled g_idx (rows assigned to groups at random, `groupsize` rows each); projections
    that the loader fuses (q/k/v, gate/up) share one, as utils/weights.py:131-135 requires."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape):
        return (torch.randn(*shape, generator=g) * std).to(F16)

    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    d = cfg.head_dim
    sd: Dict[str, torch.Tensor] = {"model.embed_tokens.weight": rnd(V, H), "lm_head.weight": rnd(V, H),
                                   "model.norm.weight": (1.0 + 0.1 * torch.randn(H, generator=g)).to(F16)}
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}"
        sd[f"{p}.input_layernorm.weight"] = (1.0 + 0.1 * torch.randn(H, generator=g)).to(F16)
        sd[f"{p}.post_attention_layernorm.weight"] = (1.0 + 0.1 * torch.randn(H, generator=g)).to(F16)
        shapes = {"self_attn.q_proj": (cfg.num_attention_heads * d, H), "self_attn.k_proj": (cfg.num_key_value_heads * d, H),
                  "self_attn.v_proj": (cfg.num_key_value_heads * d, H), "self_attn.o_proj": (H, cfg.num_attention_heads * d),
                  "mlp.gate_proj": (I, H), "mlp.up_proj": (I, H), "mlp.down_proj": (H, I)}
        shuffled = {}
        for name, (n, k) in shapes.items():
            w = rnd(n, k)
            if quantize == "gptq":
                qw, qz, sc, gi = ogptq.quantize_rtn(w, groupsize)
                if act_order:
                    fused = {"self_attn.q_proj": "qkv", "self_attn.k_proj": "qkv", "self_attn.v_proj": "qkv",
                             "mlp.gate_proj": "gate_up", "mlp.up_proj": "gate_up"}.get(name, name)
                    if fused not in shuffled:
                        shuffled[fused] = (torch.randperm(k, generator=g) // groupsize).to(torch.int32)
                    gi = shuffled[fused].clone()  # safetensors refuses tensors that share storage
                sd[f"{p}.{name}.qweight"], sd[f"{p}.{name}.qzeros"] = qw, qz
                sd[f"{p}.{name}.scales"], sd[f"{p}.{name}.g_idx"] = sc, gi
            else:
                sd[f"{p}.{name}.weight"] = w
    if quantize == "gptq":
        sd["gptq_bits"] = torch.tensor(4)
        sd["gptq_groupsize"] = torch.tensor(groupsize)
    return sd


def _shard(t: torch.Tensor, dim: int, rank: int, tp: int) -> torch.Tensor:
    """Weights.get_partial_sharded, utils/weights.py:79-101: block = size // world; [rank*block, (rank+1)*block)."""
    block = t.shape[dim] // tp
    return t.narrow(dim, rank * block, block).contiguous()


def _col_linear(sd, prefixes: List[str], rank: int, tp: int, groupsize: int) -> Linear:
    """get_multi_weights_col, utils/weights.py:115-142: every prefix sharded separately, then concatenated."""
    if f"{prefixes[0]}.qweight" in sd:
        return Linear(qweight=torch.cat([_shard(sd[f"{p}.qweight"], 1, rank, tp) for p in prefixes], 1),
                      qzeros=torch.cat([_shard(sd[f"{p}.qzeros"], 1, rank, tp) for p in prefixes], 1),
                      scales=torch.cat([_shard(sd[f"{p}.scales"], 1, rank, tp) for p in prefixes], 1),
                      g_idx=sd[f"{prefixes[0]}.g_idx"], groupsize=groupsize)
    return Linear(weight=torch.cat([_shard(sd[f"{p}.weight"], 0, rank, tp) for p in prefixes], 0))


def _row_linear(sd, prefix: str, rank: int, tp: int, groupsize: int) -> Linear:
    """get_multi_weights_row, utils/weights.py:144-201: qweight dim 0; scales/qzeros dim 0 when groupsize >= 0;
    g_idx kept only at world size 1 (:182-186) -> trivial groups of the local rows."""
    if f"{prefix}.qweight" in sd:
        qweight = _shard(sd[f"{prefix}.qweight"], 0, rank, tp)
        if groupsize >= 0:
            qzeros = _shard(sd[f"{prefix}.qzeros"], 0, rank, tp)
            scales = _shard(sd[f"{prefix}.scales"], 0, rank, tp)
        else:
            qzeros, scales = sd[f"{prefix}.qzeros"], sd[f"{prefix}.scales"]
        g_idx = sd[f"{prefix}.g_idx"] if tp == 1 else None
        return Linear(qweight=qweight, qzeros=qzeros, scales=scales, g_idx=g_idx, groupsize=groupsize)
    return Linear(weight=_shard(sd[f"{prefix}.weight"], 1, rank, tp))


def build_shards(cfg: LlamaConfig, sd: Dict[str, torch.Tensor], tp: int = 1) -> List[LlamaShard]:
    groupsize = int(sd["gptq_groupsize"]) if "gptq_groupsize" in sd else 128
    shards = []
    for rank in range(tp):
        layers = []
        for i in range(cfg.num_hidden_layers):
            p = f"model.layers.{i}"
            layers.append(LlamaLayer(
                input_ln=sd[f"{p}.input_layernorm.weight"], post_ln=sd[f"{p}.post_attention_layernorm.weight"],
                qkv=_col_linear(sd, [f"{p}.self_attn.q_proj", f"{p}.self_attn.k_proj", f"{p}.self_attn.v_proj"], rank, tp, groupsize),
                o=_row_linear(sd, f"{p}.self_attn.o_proj", rank, tp, groupsize),
                gate_up=_col_linear(sd, [f"{p}.mlp.gate_proj", f"{p}.mlp.up_proj"], rank, tp, groupsize),
                down=_row_linear(sd, f"{p}.mlp.down_proj", rank, tp, groupsize)))
        # TensorParallelEmbedding (utils/layers.py:328-338) rows sharded; TensorParallelHead (:223-231)
        # sharded on dim 0 when V % tp == 0 else replicated.
        emb = _shard(sd["model.embed_tokens.weight"], 0, rank, tp)
        V = cfg.vocab_size
        if V % tp == 0:
            head = _shard(sd["lm_head.weight"], 0, rank, tp)
        else:
            head = sd["lm_head.weight"] if rank == 0 else sd["lm_head.weight"][:0]
        shards.append(LlamaShard(cfg=cfg, embed=emb, layers=layers, norm=sd["model.norm.weight"], lm_head=head,
                                 tp=tp, rank=rank, vocab_start=rank * (V // tp)))
    return shards
