"""CPU oracle (test infrastructure only) for the flash Santacoder / StarCoder (gpt_bigcode, multi-query attention) graph.

Restates /root/reference/server/text_generation_server/models/custom_modeling/flash_santacoder_modeling.py with the fp16
rounding points of its fused ops: token + learned position embedding added in fp16 (:383-386), FastLayerNorm = residual add
+ LayerNorm with fp32 statistics (utils/layers.py:360-392), the fused `c_attn` projection [q heads | k | v] with ONE key /
value head shared by every query head (:214-224), varlen causal / decode attention (utils/flash_attn.py:43-127), GELU
(tanh form for `gelu_pytorch_tanh` / `gelu_fast`, :259-270) in fp32 rounded once, bias adds inside the linears, sequential
residual (:311-326), final LayerNorm and the head tied to `transformer.wte` (:447-449).
Pinned (tests/test_oracle_santacoder.py) against an independent implementation, transformers' GPTBigCodeForCausalLM (eager,
fp32, CPU) on the same weights.  The arithmetic of the un-vendored CUDA extensions has no reference test or golden vector:
parity unpinned at that level, as for the Llama oracle.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import llama as oll
from .neox import gelu, layernorm_residual, linear

F16 = torch.float16


@dataclass
class SantacoderConfig:
    hidden_size: int
    n_inner: int
    num_hidden_layers: int
    num_attention_heads: int
    vocab_size: int
    n_positions: int = 256
    layer_norm_epsilon: float = 1e-5
    activation_function: str = "gelu_pytorch_tanh"

    @property
    def head_dim(self):
        return self.hidden_size // self.num_attention_heads


def make_state_dict(cfg: SantacoderConfig, seed: int = 1234, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """HF-named gpt_bigcode tensors: `c_attn` rows are [h * d query rows | d key rows | d value rows] (not transposed)."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, s=std):
        return (torch.randn(*shape, generator=g) * s).to(F16)

    def ln(name):
        sd[f"{name}.weight"] = (1.0 + 0.1 * torch.randn(H, generator=g)).to(F16)
        sd[f"{name}.bias"] = rnd(H, s=0.05)

    H, I, V, d = cfg.hidden_size, cfg.n_inner, cfg.vocab_size, cfg.head_dim
    sd = {"transformer.wte.weight": rnd(V, H), "transformer.wpe.weight": rnd(cfg.n_positions, H)}
    ln("transformer.ln_f")
    for i in range(cfg.num_hidden_layers):
        p = f"transformer.h.{i}"
        ln(f"{p}.ln_1")
        ln(f"{p}.ln_2")
        sd[f"{p}.attn.c_attn.weight"] = rnd(H + 2 * d, H)
        sd[f"{p}.attn.c_attn.bias"] = rnd(H + 2 * d, s=0.05)
        sd[f"{p}.attn.c_proj.weight"] = rnd(H, H)
        sd[f"{p}.attn.c_proj.bias"] = rnd(H, s=0.05)
        sd[f"{p}.mlp.c_fc.weight"] = rnd(I, H)
        sd[f"{p}.mlp.c_fc.bias"] = rnd(I, s=0.05)
        sd[f"{p}.mlp.c_proj.weight"] = rnd(H, I)
        sd[f"{p}.mlp.c_proj.bias"] = rnd(H, s=0.05)
    return sd


class SantacoderOracle:
    """Single-rank restatement (tensor-parallel sums are associative re-groupings of the same products)."""

    def __init__(self, cfg: SantacoderConfig, sd: Dict[str, torch.Tensor]):
        self.cfg, self.sd = cfg, sd
        self.kv: Optional[List[List[Dict[str, torch.Tensor]]]] = None

    def forward(self, input_ids, position_ids, cu_seqlens: List[int], decode: bool) -> torch.Tensor:
        """Prefill (decode=False): ragged tokens, fills the per-sequence KV.  Decode: one token per sequence.
        Returns logits [T, V] fp16."""
        cfg, sd = self.cfg, self.sd
        h, d, H = cfg.num_attention_heads, cfg.head_dim, cfg.hidden_size
        B = len(cu_seqlens) - 1
        if not decode:
            self.kv = [[{"k": None, "v": None} for _ in range(B)] for _ in range(cfg.num_hidden_layers)]
        hidden = (sd["transformer.wte.weight"][input_ids].float() + sd["transformer.wpe.weight"][position_ids].float()).to(F16)
        residual = None
        scale = d ** -0.5
        tanh = cfg.activation_function in ("gelu_fast", "gelu_pytorch_tanh")
        for i in range(cfg.num_hidden_layers):
            p = f"transformer.h.{i}"
            x, residual = layernorm_residual(hidden, residual, sd[f"{p}.ln_1.weight"], sd[f"{p}.ln_1.bias"], cfg.layer_norm_epsilon)
            qkv = linear(x, sd[f"{p}.attn.c_attn.weight"], sd[f"{p}.attn.c_attn.bias"])
            q = qkv[:, :h * d].reshape(-1, h, d)
            k = qkv[:, h * d:h * d + d].reshape(-1, 1, d)
            v = qkv[:, h * d + d:].reshape(-1, 1, d)
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
            x = linear(o.reshape(-1, h * d), sd[f"{p}.attn.c_proj.weight"], sd[f"{p}.attn.c_proj.bias"])
            x, residual = layernorm_residual(x, residual, sd[f"{p}.ln_2.weight"], sd[f"{p}.ln_2.bias"], cfg.layer_norm_epsilon)
            x = linear(x, sd[f"{p}.mlp.c_fc.weight"], sd[f"{p}.mlp.c_fc.bias"])
            hidden = linear(gelu(x, tanh), sd[f"{p}.mlp.c_proj.weight"], sd[f"{p}.mlp.c_proj.bias"])
        out, _ = layernorm_residual(hidden, residual, sd["transformer.ln_f.weight"], sd["transformer.ln_f.bias"], cfg.layer_norm_epsilon)
        return linear(out, sd["transformer.wte.weight"], None)

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
