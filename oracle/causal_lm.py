"""ORACLE / CPU BASELINE (test + measurement infrastructure only — never imported by the product path).

Restates the reference's own CPU path for this workload: `CausalLM.generate_token` on the `hf_transformers` engine
(/root/reference/server/text_generation_server/models/causal_lm.py:548-739, inference_engine/hf_transformers.py:11-78):
a left-padded rectangular batch, HF `AutoModelForCausalLM.forward` in fp32 with a growing KV cache
(`past_key_values`, causal_lm.py:612-634), greedy `argmax` on the last position (utils/tokens.py:44-46), wall clock
around `model.forward` exactly as causal_lm.py:631-633.  The reference pins transformers 4.40.2 (tuple KV cache);
this image has 5.x, so the cache object is HF's DynamicCache — same arithmetic, same O(KV) growth per step.

Used by bench.py for `cpu_baseline` and for the `--impl reference` arm ("port": the reference's Python cannot be imported
unmodified here).  Bounded sample: a few layers of the named architecture at full width, scaled to
the full depth by a two-point fit (1 layer vs 1 + sample_layers).
"""
from __future__ import annotations

import os
import time

import torch

ARCHS = {
    # name: (hidden, intermediate, layers, heads, kv_heads, vocab) — must match utils/synthetic.py of the product
    "llama-2-7b": (4096, 11008, 32, 32, 32, 32000),
    "tinyllama-1.1b": (2048, 5632, 22, 32, 4, 32000),
    "llama-3-8b": (4096, 14336, 32, 32, 8, 128256),
    "llama-3-70b": (8192, 28672, 80, 64, 8, 128256),
    "tiny-test": (256, 512, 2, 4, 2, 512),
}


def _build(arch: str, n_layers: int):
    from transformers import LlamaConfig, LlamaForCausalLM
    H, I, L, h, kv, V = ARCHS[arch]
    cfg = LlamaConfig(hidden_size=H, intermediate_size=I, num_hidden_layers=n_layers, num_attention_heads=h,
                      num_key_value_heads=kv, vocab_size=V, rms_norm_eps=1e-5, max_position_embeddings=8192,
                      tie_word_embeddings=False)
    torch.manual_seed(1234)
    with torch.device("cpu"):
        model = LlamaForCausalLM(cfg).to(torch.float32).eval()
    return model, cfg


def _time_steps(model, cfg, B: int, ctx: int, steps: int, warmup: int) -> float:
    """seconds per decode step (median) with a KV cache holding ctx-1 tokens per sequence (left padding = none: the
    synthetic batch is rectangular, every prompt the same length, like the GPU arm's)."""
    from transformers import DynamicCache
    d = cfg.hidden_size // cfg.num_attention_heads
    cache = DynamicCache()
    g = torch.Generator().manual_seed(0)
    for li in range(cfg.num_hidden_layers):
        k = torch.randn(B, cfg.num_key_value_heads, ctx - 1, d, generator=g) * 0.5
        v = torch.randn(B, cfg.num_key_value_heads, ctx - 1, d, generator=g) * 0.5
        cache.update(k, v, li)
    ids = torch.randint(3, cfg.vocab_size, (B, 1), generator=g)
    times = []
    with torch.inference_mode():
        for i in range(warmup + steps):
            L = ctx - 1 + i
            attn = torch.ones(B, L + 1, dtype=torch.long)
            pos = torch.full((B, 1), L, dtype=torch.long)
            t0 = time.perf_counter()
            out = model(input_ids=ids, attention_mask=attn, position_ids=pos, past_key_values=cache, use_cache=True)
            dt = time.perf_counter() - t0
            ids = out.logits[:, -1, :].argmax(-1, keepdim=True)  # Greedy
            cache = out.past_key_values
            if i >= warmup:
                times.append(dt)
    times.sort()
    return times[len(times) // 2]


def time_decode(arch: str, B: int, ctx: int, steps: int, warmup: int, sample_layers: int = 2, max_cpu_steps: int = 6) -> dict:
    """-> {"value": tokens/s extrapolated to the full model, "cores": threads used, "sample": description,
           "ms_per_step": extrapolated full-depth milliseconds per step (value == B / that), "steps_timed": steps really timed per
           sample, "extrapolated": True}.  A 32-layer fp32 step at B = 64 takes ~6 s on 16 cores, so the sample is bounded."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_full = ARCHS[arch][2]
    steps = max(1, min(steps, max_cpu_steps))
    warmup = max(1, min(warmup, 2))
    sample_layers = max(1, min(sample_layers, n_full - 1))
    m1, c1 = _build(arch, 1)
    t1 = _time_steps(m1, c1, B, ctx, steps, warmup)
    del m1
    m2, c2 = _build(arch, 1 + sample_layers)
    t2 = _time_steps(m2, c2, B, ctx, steps, warmup)
    del m2
    t_layer = max((t2 - t1) / sample_layers, 1e-9)
    t_full = t1 + (n_full - 1) * t_layer
    sample = (f"HF LlamaForCausalLM fp32 eager CPU decode step, B={B}, context={ctx}, full width; timed 1 and "
              f"{1 + sample_layers} of {n_full} layers ({steps} steps each, median) and extrapolated linearly in depth "
              f"(t1={t1 * 1e3:.1f} ms, per-layer={t_layer * 1e3:.1f} ms, full depth={t_full * 1e3:.1f} ms per step)")
    return {"value": B / t_full, "cores": cores, "sample": sample, "ms_per_step": t_full * 1e3, "steps_timed": steps,
            "extrapolated": True}


def time_decode_gpt2(B: int = 4, ctx: int = 192, steps: int = 8, warmup: int = 2) -> dict:
    """BASELINE config[0]: the reference's CPU CausalLM path on gpt2 (124M: 12 layers, 768 wide, 12 heads, vocab 50257), fp32,
    bs = 4, sequence 128 -> 256 (mid context 192), greedy - integration_tests/test_cases_gpt2.yaml runs this model through the
    same path.  The whole model is run (no extrapolation); weights are random-init (no checkpoint in the image), which does not
    change the arithmetic per step."""
    from transformers import DynamicCache, GPT2Config, GPT2LMHeadModel
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    cfg = GPT2Config()
    model = GPT2LMHeadModel(cfg).to(torch.float32).eval()
    g = torch.Generator().manual_seed(0)
    prompt = torch.randint(0, cfg.vocab_size, (B, ctx - 1), generator=g)
    times = []
    with torch.inference_mode():
        out = model(input_ids=prompt, use_cache=True)  # prefill builds the KV cache the decode steps grow
        cache = out.past_key_values
        ids = out.logits[:, -1, :].argmax(-1, keepdim=True)
        for i in range(warmup + steps):
            L = ctx - 1 + i
            attn = torch.ones(B, L + 1, dtype=torch.long)
            t0 = time.perf_counter()
            out = model(input_ids=ids, attention_mask=attn, past_key_values=cache, use_cache=True)
            dt = time.perf_counter() - t0
            ids = out.logits[:, -1, :].argmax(-1, keepdim=True)
            cache = out.past_key_values
            if i >= warmup:
                times.append(dt)
    times.sort()
    t = times[len(times) // 2]
    return {"workload": "gpt2 CausalLM fp32 CPU, bs=4, seq 128->256 (BASELINE config[0])", "value": B / t, "unit": "tokens/s",
            "ms_per_step": t * 1e3, "steps_timed": steps, "cores": cores, "extrapolated": False, "weights": "random-init"}
