"""Synthetic checkpoints and tokenizer for benchmarks and tests: seeded random weights of a named
architecture generated directly in device memory, served through the `Weights` interface so the model code loads them
exactly like a safetensors checkpoint (same tensor names — SURVEY.md Appendix D — and the same TP slicing rules).
There is no network and no real checkpoint in the image; `data: synthetic` in bench lines refers to this.
"""
from __future__ import annotations

import types
import zlib
from typing import Dict, Optional, Tuple

import torch

from .weights import Weights

ARCHS = {
    # name: (hidden, intermediate, layers, heads, kv_heads, vocab)
    "llama-2-7b": (4096, 11008, 32, 32, 32, 32000),
    "tinyllama-1.1b": (2048, 5632, 22, 32, 4, 32000),
    "llama-3-8b": (4096, 14336, 32, 32, 8, 128256),
    "llama-3-70b": (8192, 28672, 80, 64, 8, 128256),
    "tiny-test": (256, 512, 2, 4, 2, 512),
}
NEOX_ARCHS = {
    # name: (hidden, intermediate, layers, heads, vocab, rotary_pct)
    "pythia-12b": (5120, 20480, 36, 40, 50688, 0.25),  # head size 128 (gpt-neox-20b's 96 is not a kernel instantiation)
    "tiny-neox": (256, 1024, 2, 4, 512, 0.25),
}


def llama_config(name: str, quantize: Optional[str] = None, max_position_embeddings: int = 4096, num_layers: Optional[int] = None):
    H, I, L, h, kv, V = ARCHS[name]
    cfg = types.SimpleNamespace(
        model_type="llama", hidden_size=H, intermediate_size=I, num_hidden_layers=num_layers or L, num_attention_heads=h,
        num_key_value_heads=kv, vocab_size=V, rms_norm_eps=1e-5, rope_theta=10000.0, rope_scaling=None, hidden_act="silu",
        attention_bias=False, mlp_bias=False, max_position_embeddings=max_position_embeddings, tie_word_embeddings=False,
        eos_token_id=2, pad_token_id=0, bos_token_id=1, quantize=quantize, name=name)
    cfg.to_dict = lambda: {k: v for k, v in vars(cfg).items() if not callable(v)}
    return cfg


def neox_config(name: str, max_position_embeddings: int = 2048, num_layers: Optional[int] = None):
    """GPT-NeoX (the second tensor-parallel flash family, flash_neox_modeling.py): parallel residual, partial rotary, biases."""
    H, I, L, h, V, pct = NEOX_ARCHS[name]
    cfg = types.SimpleNamespace(
        model_type="gpt_neox", hidden_size=H, intermediate_size=I, num_hidden_layers=num_layers or L, num_attention_heads=h,
        num_key_value_heads=h, vocab_size=V, rotary_pct=pct, rotary_emb_base=10000.0, layer_norm_eps=1e-5,
        use_parallel_residual=True, hidden_act="gelu_fast", max_position_embeddings=max_position_embeddings,
        tie_word_embeddings=False, eos_token_id=2, pad_token_id=0, bos_token_id=1, quantize=None, name=name)
    cfg.to_dict = lambda: {k: v for k, v in vars(cfg).items() if not callable(v)}
    return cfg


def model_config(name: str, **kw):
    """config of a named synthetic architecture of either family"""
    if name in NEOX_ARCHS:
        kw.pop("quantize", None)
        return neox_config(name, **kw)
    return llama_config(name, **kw)


class SyntheticWeights(Weights):
    """Generates every requested tensor on the device from a per-name seed; ranks slice identical full tensors."""

    def __init__(self, config, device, dtype, process_group, quantize: Optional[str] = None, groupsize: int = 128, seed: int = 1234):
        self.cfg = config
        self.device = device
        self.dtype = dtype
        self.process_group = process_group
        self.quantize = quantize
        self.gptq_bits, self.gptq_groupsize = 4, groupsize
        self.seed = seed
        self.aliases = {}

    # -- shapes ----------------------------------------------------------------------------------
    def _linear_shape(self, name: str) -> Tuple[int, int]:
        c = self.cfg
        d = c.hidden_size // c.num_attention_heads
        if getattr(c, "model_type", "llama") == "gpt_neox":
            table = {"query_key_value": (3 * c.hidden_size, c.hidden_size), "attention.dense": (c.hidden_size, c.hidden_size),
                     "dense_h_to_4h": (c.intermediate_size, c.hidden_size), "dense_4h_to_h": (c.hidden_size, c.intermediate_size)}
            for k, v in table.items():
                if f".{k}." in name:
                    return v
            raise RuntimeError(f"weight {name} does not exist")
        table = {"q_proj": (c.num_attention_heads * d, c.hidden_size), "k_proj": (c.num_key_value_heads * d, c.hidden_size),
                 "v_proj": (c.num_key_value_heads * d, c.hidden_size), "o_proj": (c.hidden_size, c.num_attention_heads * d),
                 "gate_proj": (c.intermediate_size, c.hidden_size), "up_proj": (c.intermediate_size, c.hidden_size),
                 "down_proj": (c.hidden_size, c.intermediate_size)}
        for k, v in table.items():
            if f".{k}." in name:
                return v
        raise RuntimeError(f"weight {name} does not exist")

    def get_shape(self, tensor_name: str):
        c = self.cfg
        if tensor_name in ("model.embed_tokens.weight", "lm_head.weight", "gpt_neox.embed_in.weight", "embed_out.weight"):
            return [c.vocab_size, c.hidden_size]
        if self._is_norm(tensor_name):
            return [c.hidden_size]
        if tensor_name.endswith("rotary_emb.inv_freq"):
            return [int(c.hidden_size // c.num_attention_heads * c.rotary_pct) // 2]
        if tensor_name.endswith(".bias"):
            return [self._linear_shape(tensor_name)[0]]
        n, k = self._linear_shape(tensor_name)
        g = self.gptq_groupsize if self.gptq_groupsize > 0 else k
        if tensor_name.endswith(".weight"):
            return [n, k]
        if tensor_name.endswith(".qweight"):
            return [k // 8, n]
        if tensor_name.endswith(".qzeros"):
            return [k // g, n // 8]
        if tensor_name.endswith(".scales"):
            return [k // g, n]
        if tensor_name.endswith(".g_idx"):
            return [k]
        raise RuntimeError(f"weight {tensor_name} does not exist")

    @staticmethod
    def _is_norm(tensor_name: str) -> bool:
        stem = tensor_name.rsplit(".", 1)[0]
        return stem.endswith(("layernorm", "layer_norm")) or tensor_name == "model.norm.weight"

    # -- generation ------------------------------------------------------------------------------
    def _full(self, tensor_name: str) -> torch.Tensor:
        if tensor_name == "gptq_bits":
            return torch.tensor(self.gptq_bits)
        if tensor_name == "gptq_groupsize":
            return torch.tensor(self.gptq_groupsize)
        shape = self.get_shape(tensor_name)
        gen = torch.Generator(device=self.device).manual_seed(self.seed * 1000003 + zlib.crc32(tensor_name.encode()))
        if self._is_norm(tensor_name):
            fill = torch.zeros if tensor_name.endswith(".bias") else torch.ones
            return fill(shape, dtype=self.dtype, device=self.device)
        if tensor_name.endswith("rotary_emb.inv_freq"):
            rot = 2 * shape[0]
            return 1.0 / (self.cfg.rotary_emb_base ** (torch.arange(0, rot, 2, device=self.device, dtype=torch.float32) / rot))
        if tensor_name.endswith((".qweight", ".qzeros")):
            return torch.randint(-2 ** 31, 2 ** 31 - 1, shape, generator=gen, device=self.device, dtype=torch.int32)
        if tensor_name.endswith(".scales"):
            # W = s * (q - z): (q - z) has std ~6.5, so s ~ 0.02 / 6.5 keeps W ~ N(0, 0.02^2)-like
            return (torch.rand(shape, generator=gen, device=self.device) * 0.004 + 0.001).to(torch.float16)
        if tensor_name.endswith(".g_idx"):
            g = self.gptq_groupsize if self.gptq_groupsize > 0 else shape[0]
            return (torch.arange(shape[0], device=self.device) // g).to(torch.int32)
        t = torch.empty(shape, dtype=self.dtype, device=self.device)
        t.normal_(0.0, 0.02, generator=gen)
        return t

    def get_tensor(self, tensor_name: str):
        return self._full(tensor_name)

    def get_partial_sharded(self, tensor_name: str, dim: int):
        full = self._full(tensor_name)
        world, rank = self.process_group.size(), self.process_group.rank()
        block = full.shape[dim] // world
        return full.narrow(dim, rank * block, block).contiguous()

    def get_sharded(self, tensor_name: str, dim: int):
        size = self.get_shape(tensor_name)[dim]
        world = self.process_group.size()
        assert size % world == 0, f"The choosen size {size} is not compatible with sharding on {world} shards"
        return self.get_partial_sharded(tensor_name, dim)

    def _get_gptq_params(self):
        return self.gptq_bits, self.gptq_groupsize

    def _set_gptq_params(self, model_config, model_path):
        pass


def make_tokenizer(vocab_size: int):
    """4-entry WordLevel vocabulary (<pad>, <s>, </s>, test) padded to the model's vocab so `"test " * N` is N tokens
    — the reference's own synthetic-load recipe (utils/memory_characterizer.py:219-240, utils/warmup.py:14-26)."""
    from tokenizers import Tokenizer
    from tokenizers.models import WordLevel
    from tokenizers.pre_tokenizers import WhitespaceSplit
    from transformers import PreTrainedTokenizerFast

    vocab: Dict[str, int] = {"<pad>": 0, "<s>": 1, "</s>": 2, "test": 3}
    for i in range(4, vocab_size):
        vocab[f"<tok{i}>"] = i
    tok = Tokenizer(WordLevel(vocab, unk_token="<pad>"))
    tok.pre_tokenizer = WhitespaceSplit()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, pad_token="<pad>", bos_token="<s>", eos_token="</s>",
                                   padding_side="left", truncation_side="left")
    return fast
