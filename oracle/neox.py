"""CPU oracle (test infrastructure only) for the flash GPT-NeoX graph.

Restates /root/reference/server/text_generation_server/models/custom_modeling/flash_neox_modeling.py with the fp16
rounding points of its fused ops: FastLayerNorm = residual add + LayerNorm with fp32 statistics (utils/layers.py:360-392 ->
dropout_layer_norm), QKV re-layout [h, 3, d] -> [3, h, d] (:57-80), partial half-split rotary with fp16 tables
(utils/layers.py:436-472; rotary_dim = cos.shape[-1] * 2), varlen causal / decode attention (utils/flash_attn.py:43-127),
GELU in fp32 rounded once, bias adds inside the linears, parallel residual (:232-255) or sequential (:257-281).
Pinned: cross-checked against an independent implementation, transformers' GPTNeoXForCausalLM (eager, fp32, CPU), in
tests/test_oracle_neox.py.  The arithmetic of the un-vendored CUDA extensions (flash-attn 2.5.6 layer_norm / rotary /
attention) has no reference test or golden vector: parity unpinned at that level, as for the Llama oracle.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import llama as oll

F16 = torch.float16


@dataclass
class NeoXConfig:
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    vocab_size: int
    rotary_pct: float = 0.25
    rotary_emb_base: float = 10000.0
    layer_norm_eps: float = 1e-5
    use_parallel_residual: bool = True
    hidden_act: str = "gelu"

    @property
    def head_dim(self):
        return self.hidden_size // self.num_attention_heads

    @property
    def rotary_dim(self):
        return int(self.head_dim * self.rotary_pct)


def layernorm_residual(h, residual, gamma, beta, eps):
    """FastLayerNorm.forward (utils/layers.py:360-392): x = h + residual in fp32; residual_out = fp16(x);
    normed = fp16(LN(x) * gamma + beta).  residual None -> residual_out is h."""
    x = h.float() + (residual.float() if residual is not None else 0.0)
    res_out = x.to(F16) if residual is not None else h
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    y = (x - mean) * torch.rsqrt(var + eps) * gamma.float() + beta.float()
    return y.to(F16), res_out


def gelu(x, approximate_tanh: bool):
    return torch.nn.functional.gelu(x.float(), approximate="tanh" if approximate_tanh else "none").to(F16)


def linear(x, w, b):
    """fp16 linear with fp32 accumulation, bias added before the single rounding (GEMM epilogue)."""
    y = x.float() @ w.float().t()
    if b is not None:
        y = y + b.float()
    return y.to(F16)


def make_state_dict(cfg: NeoXConfig, seed: int = 1234, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """HF-named GPT-NeoX tensors, QKV rows head-interleaved [h, 3, d] as in the checkpoints."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, s=std):
        return (torch.randn(*shape, generator=g) * s).to(F16)

    H, I, V, d = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.head_dim
    sd = {"gpt_neox.embed_in.weight": rnd(V, H), "embed_out.weight": rnd(V, H),
          "gpt_neox.final_layer_norm.weight": (1.0 + 0.1 * torch.randn(H, generator=g)).to(F16),
          "gpt_neox.final_layer_norm.bias": rnd(H, s=0.05)}
    rd = cfg.rotary_dim
    inv_freq = 1.0 / (cfg.rotary_emb_base ** (torch.arange(0, rd, 2, dtype=torch.float32) / rd))
    for i in range(cfg.num_hidden_layers):
        p = f"gpt_neox.layers.{i}"
        for ln in ("input_layernorm", "post_attention_layernorm"):
            sd[f"{p}.{ln}.weight"] = (1.0 + 0.1 * torch.randn(H, generator=g)).to(F16)
            sd[f"{p}.{ln}.bias"] = rnd(H, s=0.05)
        sd[f"{p}.attention.query_key_value.weight"] = rnd(3 * H, H)
        sd[f"{p}.attention.query_key_value.bias"] = rnd(3 * H, s=0.05)
        sd[f"{p}.attention.dense.weight"] = rnd(H, H)
        sd[f"{p}.attention.dense.bias"] = rnd(H, s=0.05)
        sd[f"{p}.attention.rotary_emb.inv_freq"] = inv_freq.clone()
        sd[f"{p}.mlp.dense_h_to_4h.weight"] = rnd(I, H)
        sd[f"{p}.mlp.dense_h_to_4h.bias"] = rnd(I, s=0.05)
        sd[f"{p}.mlp.dense_4h_to_h.weight"] = rnd(H, I)
        sd[f"{p}.mlp.dense_4h_to_h.bias"] = rnd(H, s=0.05)
    return sd


class NeoXOracle:
    """Single-rank restatement (tensor-parallel sums are associative re-groupings of the same products)."""

    def __init__(self, cfg: NeoXConfig, sd: Dict[str, torch.Tensor]):
        self.cfg, self.sd = cfg, sd
        self.kv: Optional[List[List[Dict[str, torch.Tensor]]]] = None

    def _rope(self, max_s: int):
        inv_freq = self.sd["gpt_neox.layers.0.attention.rotary_emb.inv_freq"].float()
        t = torch.arange(max_s, dtype=torch.float32)
        freqs = torch.outer(t, inv_freq)
        return torch.cos(freqs).to(F16), torch.sin(freqs).to(F16)

    def forward(self, input_ids, position_ids, cu_seqlens: List[int], decode: bool) -> torch.Tensor:
        """Prefill (decode=False): ragged tokens, fills the per-sequence KV.  Decode: one token per sequence.
        Returns the final-layer-norm output rows [T, H] projected to logits [T, V] fp16."""
        cfg, sd = self.cfg, self.sd
        h, d, H = cfg.num_attention_heads, cfg.head_dim, cfg.hidden_size
        B = len(cu_seqlens) - 1
        cos_t, sin_t = self._rope(int(position_ids.max().item()) + 1)
        cos, sin = cos_t[position_ids], sin_t[position_ids]
        if not decode:
            self.kv = [[{"k": None, "v": None} for _ in range(B)] for _ in range(cfg.num_hidden_layers)]
        hidden = sd["gpt_neox.embed_in.weight"][input_ids]
        residual = None
        scale = d ** -0.5
        tanh = cfg.hidden_act in ("gelu_fast", "gelu_pytorch_tanh")
        for i in range(cfg.num_hidden_layers):
            p = f"gpt_neox.layers.{i}"

            def attn(x):
                w = sd[f"{p}.attention.query_key_value.weight"].view(h, 3, d, H).permute(1, 0, 2, 3).reshape(-1, H)
                b = sd[f"{p}.attention.query_key_value.bias"].view(h, 3, d).permute(1, 0, 2).reshape(-1)
                qkv = linear(x, w, b).view(-1, 3, h, d)
                q = oll.apply_rotary(qkv[:, 0], cos, sin)
                k = oll.apply_rotary(qkv[:, 1], cos, sin)
                v = qkv[:, 2]
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
                return linear(o.reshape(-1, h * d), sd[f"{p}.attention.dense.weight"], sd[f"{p}.attention.dense.bias"])

            def mlp(x):
                y = linear(x, sd[f"{p}.mlp.dense_h_to_4h.weight"], sd[f"{p}.mlp.dense_h_to_4h.bias"])
                return linear(gelu(y, tanh), sd[f"{p}.mlp.dense_4h_to_h.weight"], sd[f"{p}.mlp.dense_4h_to_h.bias"])

            ln1 = (sd[f"{p}.input_layernorm.weight"], sd[f"{p}.input_layernorm.bias"])
            ln2 = (sd[f"{p}.post_attention_layernorm.weight"], sd[f"{p}.post_attention_layernorm.bias"])
            if cfg.use_parallel_residual:
                a = attn(layernorm_residual(hidden, None, *ln1, cfg.layer_norm_eps)[0])
                m = mlp(layernorm_residual(hidden, None, *ln2, cfg.layer_norm_eps)[0])
                inter = (m.float() + a.float()).to(F16)            # fp16 add
                hidden = (inter.float() + hidden.float()).to(F16)  # fp16 add
                residual = None
            else:
                x, residual = layernorm_residual(hidden, residual, *ln1, cfg.layer_norm_eps)
                x = attn(x)
                x, residual = layernorm_residual(x, residual, *ln2, cfg.layer_norm_eps)
                hidden = mlp(x)
        out, _ = layernorm_residual(hidden, residual, sd["gpt_neox.final_layer_norm.weight"], sd["gpt_neox.final_layer_norm.bias"],
                                    cfg.layer_norm_eps)
        return linear(out, sd["embed_out.weight"], None)

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
