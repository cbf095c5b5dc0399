"""CPU oracle (test infrastructure only) for the flash Falcon / RefinedWeb graph (multi-query and grouped-query attention).

Restates /root/reference/server/text_generation_server/models/custom_modeling/flash_rw_modeling.py with the fp16 rounding
points of its fused ops, in its three layer forms:
  * "RefinedWebModel" with parallel_attn (Falcon-7B, FlashRWLayer :339-356): ONE LayerNorm feeds attention and MLP, their
    outputs are added in fp16 and re-enter the residual stream through the next layer's fused residual add;
  * "RefinedWebModel" without parallel_attn (:357-374): sequential, two LayerNorms;
  * "RefinedWeb" / new_decoder_architecture (Falcon-40B, FlashRWLargeLayer :376-424): `ln_attn` (with the residual add) and
    `ln_mlp` (of the updated residual) feed attention and MLP.
Fused projection layouts: the first two keep [q heads | k heads | v heads] (:156-163); the large form interleaves per KV group
[q heads of the group | k | v] (:233-240).  Rotary embedding: full head, half-split pairs, base 10000, fp16 tables
(:126-128, utils/layers.py:436-472).  GELU is the exact (erf) form (:286).
Pinned (tests/test_oracle_falcon.py) against the reference's OWN FlashRWForCausalLM executed on CPU
(tests/golden/flash_rw_ref.npz, written by tests/golden/make_golden.py with the CUDA extensions shimmed by the oracle's
restatements) and against an independent implementation, transformers' FalconForCausalLM (eager, fp32, CPU).  The arithmetic
of the un-vendored CUDA extensions has no reference test or golden vector: parity unpinned at that level.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import llama as oll
from .neox import gelu, layernorm_residual, linear

F16 = torch.float16


@dataclass
class FalconConfig:
    hidden_size: int
    num_hidden_layers: int
    n_head: int
    n_head_kv: int
    vocab_size: int
    new_decoder_architecture: bool = False
    parallel_attn: bool = True
    bias: bool = False
    layer_norm_epsilon: float = 1e-5

    @property
    def head_dim(self):
        return self.hidden_size // self.n_head


def make_state_dict(cfg: FalconConfig, seed: int = 1234, std: float = 0.02) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, s=std):
        return (torch.randn(*shape, generator=g) * s).to(F16)

    def ln(name):
        sd[f"{name}.weight"] = (1.0 + 0.1 * torch.randn(H, generator=g)).to(F16)
        sd[f"{name}.bias"] = rnd(H, s=0.05)

    def lin(name, n_out, n_in):
        sd[f"{name}.weight"] = rnd(n_out, n_in)
        if cfg.bias:
            sd[f"{name}.bias"] = rnd(n_out, s=0.05)

    H, V, d = cfg.hidden_size, cfg.vocab_size, cfg.head_dim
    sd = {"transformer.word_embeddings.weight": rnd(V, H), "lm_head.weight": rnd(V, H)}
    ln("transformer.ln_f")
    for i in range(cfg.num_hidden_layers):
        p = f"transformer.h.{i}"
        if cfg.new_decoder_architecture:
            ln(f"{p}.ln_attn")
            ln(f"{p}.ln_mlp")
        else:
            ln(f"{p}.input_layernorm")
            if not cfg.parallel_attn:
                ln(f"{p}.post_attention_layernorm")
        lin(f"{p}.self_attention.query_key_value", (cfg.n_head + 2 * cfg.n_head_kv) * d, H)
        lin(f"{p}.self_attention.dense", H, H)
        lin(f"{p}.mlp.dense_h_to_4h", 4 * H, H)
        lin(f"{p}.mlp.dense_4h_to_h", H, 4 * H)
    return sd


def split_qkv(cfg: FalconConfig, qkv: torch.Tensor):
    """fused projection output [T, (h + 2 kv) d] -> q [T, h, d], k [T, kv, d], v [T, kv, d] for the layout of the architecture"""
    h, kv, d = cfg.n_head, cfg.n_head_kv, cfg.head_dim
    if cfg.new_decoder_architecture:
        grouped = qkv.view(-1, kv, h // kv + 2, d)
        return grouped[:, :, :-2].reshape(-1, h, d), grouped[:, :, -2], grouped[:, :, -1]
    q, rest = qkv[:, :h * d], qkv[:, h * d:].view(-1, 2, kv, d)
    return q.reshape(-1, h, d), rest[:, 0], rest[:, 1]


class FalconOracle:
    """Single-rank restatement (tensor-parallel sums are associative re-groupings of the same products)."""

    def __init__(self, cfg: FalconConfig, sd: Dict[str, torch.Tensor]):
        self.cfg, self.sd = cfg, sd
        self.kv: Optional[List[List[Dict[str, torch.Tensor]]]] = None

    def _rope(self, max_s: int):
        d = self.cfg.head_dim
        inv_freq = 1.0 / (10000.0 ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
        freqs = torch.outer(torch.arange(max_s, dtype=torch.float32), inv_freq)
        return torch.cos(freqs).to(F16), torch.sin(freqs).to(F16)

    def _lin(self, x, name):
        return linear(x, self.sd[f"{name}.weight"], self.sd.get(f"{name}.bias"))

    def forward(self, input_ids, position_ids, cu_seqlens: List[int], decode: bool) -> torch.Tensor:
        cfg, sd = self.cfg, self.sd
        h, d = cfg.n_head, cfg.head_dim
        B = len(cu_seqlens) - 1
        cos_t, sin_t = self._rope(int(position_ids.max().item()) + 1)
        cos, sin = cos_t[position_ids], sin_t[position_ids]
        if not decode:
            self.kv = [[{"k": None, "v": None} for _ in range(B)] for _ in range(cfg.num_hidden_layers)]
        hidden = sd["transformer.word_embeddings.weight"][input_ids]
        residual = None
        scale = d ** -0.5
        eps = cfg.layer_norm_epsilon
        for i in range(cfg.num_hidden_layers):
            p = f"transformer.h.{i}"

            def attn(x):
                q, k, v = split_qkv(cfg, self._lin(x, f"{p}.self_attention.query_key_value"))
                q, k = oll.apply_rotary(q, cos, sin), oll.apply_rotary(k, cos, sin)
                if not decode:
                    for bi in range(B):
                        s, e = cu_seqlens[bi], cu_seqlens[bi + 1]
                        self.kv[i][bi]["k"], self.kv[i][bi]["v"] = k[s:e].clone(), v[s:e].clone()
                    o = oll.attention_prefill(q, k, v, cu_seqlens, scale)
                else:
                    for bi in range(B):
                        self.kv[i][bi]["k"] = torch.cat([self.kv[i][bi]["k"], k[bi:bi + 1]])
                        self.kv[i][bi]["v"] = torch.cat([self.kv[i][bi]["v"], v[bi:bi + 1]])
                    o = oll.attention_decode(q, [c["k"] for c in self.kv[i]], [c["v"] for c in self.kv[i]], scale)
                return self._lin(o.reshape(-1, h * d), f"{p}.self_attention.dense")

            def mlp(x):
                return self._lin(gelu(self._lin(x, f"{p}.mlp.dense_h_to_4h"), False), f"{p}.mlp.dense_4h_to_h")

            def norm(name, x, res):
                return layernorm_residual(x, res, sd[f"{p}.{name}.weight"], sd[f"{p}.{name}.bias"], eps)

            if cfg.new_decoder_architecture:
                ln_attn, residual = norm("ln_attn", hidden, residual)
                ln_mlp, _ = norm("ln_mlp", residual, None)
                hidden = (attn(ln_attn).float() + mlp(ln_mlp).float()).to(F16)  # fp16 add
            elif cfg.parallel_attn:
                x, residual = norm("input_layernorm", hidden, residual)
                hidden = (mlp(x).float() + attn(x).float()).to(F16)              # fp16 add
            else:
                x, residual = norm("input_layernorm", hidden, residual)
                x, residual = norm("post_attention_layernorm", attn(x), residual)
                hidden = mlp(x)
        out, _ = layernorm_residual(hidden, residual, sd["transformer.ln_f.weight"], sd["transformer.ln_f.bias"], eps)
        return linear(out, sd["lm_head.weight"], None)

    def generate_greedy(self, prompts: List[List[int]], n_new: int):
        """-> (tokens [B, n_new], [logits of the last prompt token / of every decode step])"""
        lens = [len(p) for p in prompts]
        cu = [0]
        for L in lens:
            cu.append(cu[-1] + L)
        ids = torch.tensor([t for p in prompts for t in p])
        pos = torch.cat([torch.arange(L) for L in lens])
        logits = self.forward(ids, pos, cu, decode=False)
        last = torch.tensor(cu[1:]) - 1
        step_logits = [logits[last]]
        toks = [step_logits[0].float().argmax(-1)]
        cur = list(lens)
        B = len(prompts)
        for _ in range(1, n_new):
            lg = self.forward(toks[-1], torch.tensor(cur), list(range(B + 1)), decode=True)
            cur = [c + 1 for c in cur]
            step_logits.append(lg)
            toks.append(lg.float().argmax(-1))
        return torch.stack(toks, 1), step_logits
