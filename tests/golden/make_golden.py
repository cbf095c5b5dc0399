"""Generates the golden fixtures in tests/golden/ by EXECUTING THE REFERENCE'S OWN PYTHON in this container
(/root/reference is read-only and absent on the GPU box, so the outputs are committed as small .npz/.json files).

    python tests/golden/make_golden.py          # needs /root/reference; CPU only

What runs is reference code, imported from /root/reference/server/text_generation_server by file path:
  utils/gptq/quant_linear.py   QuantLinear.pack                       -> gptq_pack.npz
  utils/layers.py              PositionRotaryEmbedding / LinearScaling -> rope_tables.npz
  utils/weights.py             Weights.get_multi_weights_col/row, get_partial_sharded (tp = 2) -> weights_shards.npz
  utils/tokens.py + logits_process.py  HeterogeneousNextTokenChooser   -> chooser.npz
  models/custom_modeling/flash_llama_modeling.py  FlashLlamaForCausalLM.forward (prefill + decode, CPU fp16)
                                                                      -> flash_llama_ref.npz
  models/custom_modeling/flash_neox_modeling.py   FlashGPTNeoXForCausalLM.forward (both residual forms; `--neox-only` regenerates
                                                  just this one)      -> flash_neox_ref.npz
  models/flash_causal_lm.py    FlashCausalLMBatch bookkeeping         -> batch_bookkeeping.npz
  proto/generate.proto (parsed, not executed)                         -> generate_proto_fields.json
Stubs, and only these: modules that are absent from the image (`accelerate`, `loguru`-free paths, `rotary_emb`,
`dropout_layer_norm`, `flash_attn_2_cuda`, the protoc-generated `pb.generate_pb2`) are replaced by thin shims; the
three un-vendored CUDA kernels are shimmed with the oracle's restatement of their arithmetic (oracle/llama.py, oracle/neox.py),
so `flash_llama_ref.npz` / `flash_neox_ref.npz` pin the reference's *graph wiring, weight loading and KV handling* — its own Python — while the
kernel arithmetic stays "parity unpinned" (SURVEY.md §8c).
"""
from __future__ import annotations

import importlib.util
import json
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/server/text_generation_server"
sys.path.insert(0, ROOT)

from oracle import llama as oll  # noqa: E402


def _pkg(name, path=None):
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    sys.modules[name] = m
    return m


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def install_stubs():
    import transformers  # noqa: F401  (before the `accelerate` shim below: transformers probes for it at import)
    import transformers.activations  # noqa: F401
    import transformers.generation.logits_process  # noqa: F401
    import tgis_b200  # noqa: F401
    from tgis_b200 import pb as my_pb

    tgs = _pkg("text_generation_server", REF)
    utils = _pkg("text_generation_server.utils", os.path.join(REF, "utils"))
    utils.print_rank_n = lambda *a, **k: None
    _pkg("text_generation_server.utils.gptq", os.path.join(REF, "utils", "gptq"))
    pbpkg = _pkg("text_generation_server.pb")
    pbpkg.generate_pb2 = my_pb
    sys.modules["text_generation_server.pb.generate_pb2"] = my_pb
    # absent third-party modules
    acc = types.ModuleType("accelerate")
    import contextlib
    acc.init_empty_weights = contextlib.nullcontext
    sys.modules["accelerate"] = acc
    if "loguru" not in sys.modules:
        try:
            import loguru  # noqa: F401
        except ImportError:
            lg = types.ModuleType("loguru")
            lg.logger = types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None)
            sys.modules["loguru"] = lg
    # rotary_emb.apply_rotary(x1, x2, cos, sin, out1, out2, conj): oracle restatement, in place
    rot = types.ModuleType("rotary_emb")

    def apply_rotary(x1, x2, cos, sin, o1, o2, conj):
        a, b = x1.float(), x2.float()
        c, s = cos.float(), sin.float()
        r1 = (a * c - b * s).to(x1.dtype)
        r2 = (a * s + b * c).to(x1.dtype)
        o1.copy_(r1)
        o2.copy_(r2)
    rot.apply_rotary = apply_rotary
    sys.modules["rotary_emb"] = rot
    # dropout_layer_norm.dropout_add_ln_fwd(...) in RMSNorm mode: oracle restatement
    dln = types.ModuleType("dropout_layer_norm")

    def dropout_add_ln_fwd(x, residual, gamma, beta, rowscale, colscale, x0_subset, z_subset, p, eps, rs, zn, gen, res_fp32, is_rms):
        if is_rms:
            assert beta is None
            normed, res = oll.rmsnorm_residual(x, residual, gamma, eps)
        else:  # LayerNorm mode (FastLayerNorm, utils/layers.py:376-392): the GPT-NeoX family
            from oracle import neox as onx
            normed, res = onx.layernorm_residual(x, residual, gamma, beta, eps)
        return normed, (res if residual is not None else None), None, None, None
    dln.dropout_add_ln_fwd = dropout_add_ln_fwd
    sys.modules["dropout_layer_norm"] = dln
    # utils.flash_attn.attention: the module refuses to import without CUDA; shim with the oracle's attention
    fa = types.ModuleType("text_generation_server.utils.flash_attn")

    def attention(q, k, v, cu_seqlens, max_s, softmax_scale, cu_seqlens_q=None, max_s_q=None, causal=True):
        cu = [int(c) for c in cu_seqlens]
        if cu_seqlens_q is None:
            return oll.attention_prefill(q, k, v, cu, softmax_scale)
        ks = [k[cu[b]:cu[b + 1]] for b in range(len(cu) - 1)]
        vs = [v[cu[b]:cu[b + 1]] for b in range(len(cu) - 1)]
        return oll.attention_decode(q, ks, vs, softmax_scale)
    fa.attention = attention
    sys.modules["text_generation_server.utils.flash_attn"] = fa
    return my_pb


# ----------------------------------------------------------------------------------------------------------
def gold_gptq_pack():
    ql = _load("text_generation_server.utils.gptq.quant_linear", "utils/gptq/quant_linear.py")
    g = torch.Generator().manual_seed(11)
    N, K, gs = 64, 256, 128
    lin = torch.nn.Linear(K, N, bias=False)
    lin.weight.data = torch.randn(N, K, generator=g) * 0.05
    G = K // gs
    # per-(out-feature, group) scales/zeros as GPTQ produces them: [N, G]
    w = lin.weight.data.reshape(N, G, gs)
    scale = (w.amax(-1) - w.amin(-1)).clamp(min=1e-5) / 15
    zero = torch.round(-w.amin(-1) / scale).clamp(1, 16)
    q = ql.QuantLinear.new(4, gs, K, N, False)
    q.pack(lin, scale.clone(), zero.clone(), None)
    np.savez(os.path.join(HERE, "gptq_pack.npz"), weight=lin.weight.data.numpy(), scales_in=scale.numpy(), zeros_in=zero.numpy(),
             qweight=q.qweight.numpy(), qzeros=q.qzeros.numpy(), scales=q.scales.numpy(), g_idx=q.g_idx.numpy(), groupsize=gs)
    print("gptq_pack.npz", q.qweight.shape, q.qzeros.shape)


def gold_rope(layers):
    out = {}
    for name, (d, base, factor) in {"d64": (64, 10000.0, 1.0), "d128": (128, 10000.0, 1.0), "d128_theta5e5": (128, 500000.0, 1.0),
                                    "d64_linear2": (64, 10000.0, 2.0)}.items():
        if factor == 1.0:
            rot = layers.PositionRotaryEmbedding.static(dim=d, base=base, device="cpu")
        else:
            rot = layers.LinearScalingPositionRotaryEmbedding.static(dim=d, base=base, scaling_factor=factor, device="cpu")
        pos = torch.tensor([0, 1, 5, 63, 200, 1023])
        cos, sin = rot.get_cos_sin(pos, 1024, torch.float16)
        out[f"{name}_cos"] = cos.squeeze(1).numpy()
        out[f"{name}_sin"] = sin.squeeze(1).numpy()
    out["positions"] = np.array([0, 1, 5, 63, 200, 1023])
    np.savez(os.path.join(HERE, "rope_tables.npz"), **out)
    print("rope_tables.npz")


class _Group:
    def __init__(self, rank, size):
        self._r, self._s = rank, size

    def rank(self):
        return self._r

    def size(self):
        return self._s


def gold_weights(weights_mod, tmpdir):
    from safetensors.torch import save_file
    # the exllama (GPTQ CUDA) branch of get_multi_weights_row is the hot path; without the un-vendored kernels the
    # module-level flag is False and the loader would take the Triton branch (full scales/zeros + sharded g_idx)
    sys.modules["text_generation_server.utils.layers"].HAS_GPTQ_CUDA = True
    cfg = oll.LlamaConfig(256, 512, 1, 4, 2, 512)
    out = {}
    for quant in (None, "gptq"):
        sd = oll.make_state_dict(cfg, seed=5, quantize=quant)
        path = os.path.join(tmpdir, f"w_{quant}.safetensors")
        save_file({k: v.contiguous() for k, v in sd.items()}, path)
        for rank in (0, 1):
            w = weights_mod.Weights([path], device="cpu", dtype=torch.float16, process_group=_Group(rank, 2))
            if quant == "gptq":
                w.gptq_bits, w.gptq_groupsize = 4, 128
            p = "model.layers.0"
            col = w.get_multi_weights_col([f"{p}.self_attn.q_proj", f"{p}.self_attn.k_proj", f"{p}.self_attn.v_proj"], quant, 0)
            row = w.get_multi_weights_row(f"{p}.mlp.down_proj", quant)
            emb = w.get_partial_sharded("model.embed_tokens.weight", dim=0)
            tag = f"{quant or 'fp16'}_r{rank}"
            if quant == "gptq":
                for nm, t in zip(("qweight", "qzeros", "scales"), col[:3]):
                    out[f"{tag}_col_{nm}"] = t.numpy()
                for nm, t in zip(("qweight", "qzeros", "scales"), row[:3]):
                    out[f"{tag}_row_{nm}"] = t.numpy()
                out[f"{tag}_row_gidx_is_none"] = np.array(row[3] is None)
            else:
                out[f"{tag}_col"] = col.numpy()
                out[f"{tag}_row"] = row.numpy()
            out[f"{tag}_emb"] = emb.numpy()
    np.savez(os.path.join(HERE, "weights_shards.npz"), **out)
    print("weights_shards.npz", len(out))


def gold_chooser(my_pb):
    _load("text_generation_server.utils.dist", "utils/dist.py")
    _load("text_generation_server.utils.token_types", "utils/token_types.py")
    # transformers 5.x dropped LogitsWarper (the reference pins 4.40.2): alias it for the import only
    import transformers
    if not hasattr(transformers, "LogitsWarper"):
        transformers.LogitsWarper = transformers.LogitsProcessor
    import transformers.generation.logits_process as tlp
    if not hasattr(tlp, "LogitsWarper"):
        tlp.LogitsWarper = tlp.LogitsProcessor
    _load("text_generation_server.utils.logits_process", "utils/logits_process.py")
    tokens = _load("text_generation_server.utils.tokens", "utils/tokens.py")
    P = my_pb.NextTokenChooserParameters
    params = [
        P(temperature=0.0, top_p=1.0),                                                 # greedy
        P(temperature=0.7, top_k=5, top_p=1.0, seed=3),                                # top-k sampling
        P(temperature=1.3, top_p=0.8, seed=4, repetition_penalty=1.3),                 # top-p + repetition penalty
        P(temperature=1.0, typical_p=0.6, top_p=1.0, seed=5, min_new_tokens=2),        # typical + min_new_tokens
        P(temperature=0.0, top_p=1.0, length_penalty=P.LengthPenalty(start_index=1, decay_factor=1.5)),
    ]
    B, V, steps = len(params), 97, 4
    g = torch.Generator().manual_seed(21)
    all_scores = torch.randn(steps, B, V, generator=g) * 3
    input_ids = torch.randint(0, V, (B, 6), generator=g)
    ch = tokens.HeterogeneousNextTokenChooser.from_pb(pb=params, model_eos_token_id=2, model_pad_token_id=0,
                                                      return_logprobs=[False, True, False, True, False], dtype=torch.float32,
                                                      device=torch.device("cpu"))
    ids_out, scores_out, lp_out = [], [], []
    for s in range(steps):
        ids, sc, lp = ch(input_ids, all_scores[s].clone())
        ids_out.append(ids.numpy().copy())
        scores_out.append(sc.numpy().copy())
        lp_out.append(lp.numpy().copy())
    # after filtering to requests [1, 3, 4] (prune path)
    ch2 = ch.filter([1, 3, 4])
    ids_f, _, _ = ch2(input_ids[[1, 3, 4]], all_scores[0][[1, 3, 4]].clone())
    np.savez(os.path.join(HERE, "chooser.npz"), all_scores=all_scores.numpy(), input_ids=input_ids.numpy(),
             ids=np.stack(ids_out), scores=np.stack(scores_out), logprobs=np.stack(lp_out), ids_filtered=ids_f.numpy(),
             params=np.frombuffer(b"\n".join(p.SerializeToString().hex().encode() for p in params), dtype=np.uint8))
    print("chooser.npz", np.stack(ids_out).tolist())


def gold_flash_llama(layers, weights_mod, tmpdir):
    from safetensors.torch import save_file
    fl = _load("text_generation_server.models.custom_modeling.flash_llama_modeling", "models/custom_modeling/flash_llama_modeling.py")
    dist_mod = sys.modules.get("text_generation_server.utils.dist") or _load("text_generation_server.utils.dist", "utils/dist.py")
    cases = {"mha": oll.LlamaConfig(128, 256, 2, 2, 2, 160), "gqa": oll.LlamaConfig(256, 512, 2, 8, 2, 192)}
    out = {}
    for name, cfg in cases.items():
        sd = oll.make_state_dict(cfg, seed=31, std=0.08)
        path = os.path.join(tmpdir, f"fl_{name}.safetensors")
        save_file({k: v.contiguous() for k, v in sd.items()}, path)
        w = weights_mod.Weights([path], device="cpu", dtype=torch.float16, process_group=dist_mod.FakeGroup(0, 1))
        hf = fl.LlamaConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                            num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                            num_key_value_heads=cfg.num_key_value_heads, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta)
        hf.quantize = None
        model = fl.FlashLlamaForCausalLM(hf, w)
        g = torch.Generator().manual_seed(9)
        lens = [5, 12, 1]
        prompts = [torch.randint(0, cfg.vocab_size, (L,), generator=g) for L in lens]
        input_ids = torch.cat(prompts)
        position_ids = torch.cat([torch.arange(L) for L in lens])
        cu = torch.tensor([0, 5, 17, 18], dtype=torch.int32)
        with torch.no_grad():
            logits, present = model.forward(input_ids, position_ids, cu, None, max(lens), None, None, None)
            out[f"{name}_prefill_logits"] = logits.numpy()
            # decode exactly as FlashCausalLM.generate_token does for batch > 1 (flash_causal_lm.py:436-447, 457-458)
            B = len(lens)
            pad = present.new_zeros(present.shape[0], 1, *present.shape[2:])
            pieces, start = [], 0
            for i in range(1, B + 1):
                pieces += [present[:, start:int(cu[i])], pad]
                start = int(cu[i])
            past = torch.cat(pieces, dim=1)
            cu_q = torch.arange(B + 1, dtype=torch.int32)
            nxt = logits[(cu[1:] - 1).long()].float().argmax(-1)   # last prompt token rows (flash_causal_lm.py:519)
            cu = cu + cu_q
            out[f"{name}_first_tokens"] = nxt.numpy()
            pos = torch.tensor(lens)
            for step in range(2):
                logits, present = model.forward(nxt, pos, cu, cu_q, max(lens) + 1 + step, None, past, None)
                out[f"{name}_decode{step}_logits"] = logits.numpy()
                pieces, start = [], 0
                for i in range(1, B + 1):
                    pieces += [present[:, start:int(cu[i])], pad]
                    start = int(cu[i])
                past = torch.cat(pieces, dim=1)
                cu = cu + cu_q
                pos = pos + 1
                nxt = logits.float().argmax(-1)
                out[f"{name}_decode{step}_tokens"] = nxt.numpy()
        out[f"{name}_input_ids"] = input_ids.numpy()
        out[f"{name}_lens"] = np.array(lens)
        out[f"{name}_cfg"] = np.array([cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.num_attention_heads,
                                       cfg.num_key_value_heads, cfg.vocab_size])
    np.savez(os.path.join(HERE, "flash_llama_ref.npz"), **out)
    print("flash_llama_ref.npz", sorted(out)[:4], "...")


def gold_flash_neox(layers, weights_mod, tmpdir):
    """The reference's own FlashGPTNeoXForCausalLM (flash_neox_modeling.py) on CPU: prefill + 2 decode steps, both residual
    forms -> flash_neox_ref.npz (graph wiring, QKV re-layout, partial rotary and KV handling pinned; the three CUDA
    extensions are the oracle shims above)."""
    from safetensors.torch import save_file
    from oracle import neox as onx
    fn = _load("text_generation_server.models.custom_modeling.flash_neox_modeling", "models/custom_modeling/flash_neox_modeling.py")
    dist_mod = sys.modules.get("text_generation_server.utils.dist") or _load("text_generation_server.utils.dist", "utils/dist.py")
    from transformers import GPTNeoXConfig
    out = {}
    for name, parallel in (("parallel", True), ("sequential", False)):
        cfg = onx.NeoXConfig(128, 512, 2, 2, 160, rotary_pct=0.25, use_parallel_residual=parallel)
        sd = onx.make_state_dict(cfg, seed=17, std=0.06)
        path = os.path.join(tmpdir, f"neox_{name}.safetensors")
        save_file({k: v.contiguous() for k, v in sd.items()}, path)
        w = weights_mod.Weights([path], device="cpu", dtype=torch.float16, process_group=dist_mod.FakeGroup(0, 1))
        hf = GPTNeoXConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                           num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads, rotary_pct=cfg.rotary_pct,
                           rotary_emb_base=10000, layer_norm_eps=cfg.layer_norm_eps, use_parallel_residual=parallel, hidden_act="gelu")
        hf.quantize = None
        model = fn.FlashGPTNeoXForCausalLM(hf, w)
        g = torch.Generator().manual_seed(9)
        lens = [5, 12, 1]
        prompts = [torch.randint(0, cfg.vocab_size, (L,), generator=g) for L in lens]
        input_ids = torch.cat(prompts)
        position_ids = torch.cat([torch.arange(L) for L in lens])
        cu = torch.tensor([0, 5, 17, 18], dtype=torch.int32)
        with torch.no_grad():
            logits, present = model.forward(input_ids, position_ids, cu, None, max(lens), None, None, None)
            out[f"{name}_prefill_logits"] = logits.numpy()
            B = len(lens)
            pad = present.new_zeros(present.shape[0], 1, *present.shape[2:])

            def repad(present, cu):
                pieces, start = [], 0
                for i in range(1, B + 1):
                    pieces += [present[:, start:int(cu[i])], pad]
                    start = int(cu[i])
                return torch.cat(pieces, dim=1)
            past = repad(present, cu)
            cu_q = torch.arange(B + 1, dtype=torch.int32)
            nxt = logits[(cu[1:] - 1).long()].float().argmax(-1)
            cu = cu + cu_q
            pos = torch.tensor(lens)
            for step in range(2):
                out[f"{name}_decode{step}_input"] = nxt.numpy()
                logits, present = model.forward(nxt, pos, cu, cu_q, max(lens) + 1 + step, None, past, None)
                out[f"{name}_decode{step}_logits"] = logits.numpy()
                past = repad(present, cu)
                cu = cu + cu_q
                pos = pos + 1
                nxt = logits.float().argmax(-1)
        out[f"{name}_input_ids"] = input_ids.numpy()
        out[f"{name}_lens"] = np.array(lens)
    np.savez(os.path.join(HERE, "flash_neox_ref.npz"), **out)
    print("flash_neox_ref.npz", sorted(out)[:4], "...")


def gold_flash_santacoder(layers, weights_mod, tmpdir):
    """The reference's own FlashSantacoderForCausalLM (flash_santacoder_modeling.py, multi-query attention) on CPU: prefill + 2
    decode steps -> flash_santacoder_ref.npz (c_attn q | kv split, learned positions, KV placement, sequential residual,
    tied head pinned; LayerNorm and attention are the oracle shims above)."""
    from safetensors.torch import save_file
    from oracle import santacoder as osc
    fs = _load("text_generation_server.models.custom_modeling.flash_santacoder_modeling",
               "models/custom_modeling/flash_santacoder_modeling.py")
    dist_mod = sys.modules.get("text_generation_server.utils.dist") or _load("text_generation_server.utils.dist", "utils/dist.py")
    from transformers import GPTBigCodeConfig
    cfg = osc.SantacoderConfig(128, 512, 2, 4, 160, n_positions=64)
    sd = osc.make_state_dict(cfg, seed=23, std=0.06)
    path = os.path.join(tmpdir, "santacoder.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    w = weights_mod.Weights([path], device="cpu", dtype=torch.float16, process_group=dist_mod.FakeGroup(0, 1))
    hf = GPTBigCodeConfig(vocab_size=cfg.vocab_size, n_positions=cfg.n_positions, n_embd=cfg.hidden_size, n_layer=cfg.num_hidden_layers,
                          n_head=cfg.num_attention_heads, n_inner=cfg.n_inner, activation_function=cfg.activation_function,
                          layer_norm_epsilon=cfg.layer_norm_epsilon, multi_query=True)
    hf.quantize, hf.transpose = None, False  # tgis_native.py:84: transpose only for GPT2-architecture checkpoints
    model = fs.FlashSantacoderForCausalLM(hf, w)
    g = torch.Generator().manual_seed(13)
    lens = [6, 11, 2]
    prompts = [torch.randint(0, cfg.vocab_size, (L,), generator=g) for L in lens]
    input_ids = torch.cat(prompts)
    position_ids = torch.cat([torch.arange(L) for L in lens])
    cu = torch.tensor([0, 6, 17, 19], dtype=torch.int32)
    out = {}
    with torch.no_grad():
        logits, present = model.forward(input_ids, position_ids, cu, None, max(lens), None, None, None)
        out["prefill_logits"] = logits.numpy()
        B = len(lens)
        pad = present.new_zeros(present.shape[0], 1, *present.shape[2:])

        def repad(present, cu):
            pieces, start = [], 0
            for i in range(1, B + 1):
                pieces += [present[:, start:int(cu[i])], pad]
                start = int(cu[i])
            return torch.cat(pieces, dim=1)
        past = repad(present, cu)
        cu_q = torch.arange(B + 1, dtype=torch.int32)
        nxt = logits[(cu[1:] - 1).long()].float().argmax(-1)
        cu = cu + cu_q
        pos = torch.tensor(lens)
        for step in range(2):
            out[f"decode{step}_input"] = nxt.numpy()
            logits, present = model.forward(nxt, pos, cu, cu_q, max(lens) + 1 + step, None, past, None)
            out[f"decode{step}_logits"] = logits.numpy()
            past = repad(present, cu)
            cu = cu + cu_q
            pos = pos + 1
            nxt = logits.float().argmax(-1)
    out["input_ids"] = input_ids.numpy()
    out["lens"] = np.array(lens)
    np.savez(os.path.join(HERE, "flash_santacoder_ref.npz"), **out)
    print("flash_santacoder_ref.npz", sorted(out)[:4], "...")


FALCON_CASES = {  # name -> (hidden, layers, heads, kv heads, vocab, new_decoder_architecture, parallel_attn, bias)
    "mqa_parallel": (128, 2, 4, 1, 160, False, True, False),
    "gqa_large": (256, 2, 8, 2, 160, True, True, False),
    "mqa_sequential_bias": (128, 2, 2, 1, 160, False, False, True),
}


def gold_flash_rw(layers, weights_mod, tmpdir):
    """The reference's own FlashRWForCausalLM (flash_rw_modeling.py: Falcon / RefinedWeb) on CPU in its three layer forms:
    prefill + 2 decode steps -> flash_rw_ref.npz (fused-QKV layouts, rotary on q and k, KV placement, parallel / sequential /
    two-LayerNorm residual wiring pinned; LayerNorm, rotary and attention are the oracle shims above)."""
    from safetensors.torch import save_file
    from oracle import falcon as ofa
    fr = _load("text_generation_server.models.custom_modeling.flash_rw_modeling", "models/custom_modeling/flash_rw_modeling.py")
    dist_mod = sys.modules.get("text_generation_server.utils.dist") or _load("text_generation_server.utils.dist", "utils/dist.py")
    out = {}
    for name, (H, L, h, kvh, V, large, parallel, bias) in FALCON_CASES.items():
        cfg = ofa.FalconConfig(H, L, h, kvh, V, new_decoder_architecture=large, parallel_attn=parallel, bias=bias)
        sd = ofa.make_state_dict(cfg, seed=29, std=0.06)
        path = os.path.join(tmpdir, f"falcon_{name}.safetensors")
        save_file({k: v.contiguous() for k, v in sd.items()}, path)
        w = weights_mod.Weights([path], device="cpu", dtype=torch.float16, process_group=dist_mod.FakeGroup(0, 1))
        rw = fr.RWConfig(model_type="RefinedWeb" if large else "RefinedWebModel", vocab_size=V, hidden_size=H, num_hidden_layers=L,
                         num_attention_heads=h, num_kv_heads=kvh, new_decoder_architecture=large, bias=bias, parallel_attn=parallel)
        rw.quantize = None
        model = fr.FlashRWForCausalLM(rw, w)
        g = torch.Generator().manual_seed(15)
        lens = [4, 13, 2]
        prompts = [torch.randint(0, V, (n,), generator=g) for n in lens]
        input_ids = torch.cat(prompts)
        position_ids = torch.cat([torch.arange(n) for n in lens])
        cu = torch.tensor([0, 4, 17, 19], dtype=torch.int32)
        with torch.no_grad():
            logits, present = model.forward(input_ids, position_ids, cu, None, max(lens), None, None, None)
            out[f"{name}_prefill_logits"] = logits.numpy()
            B = len(lens)
            pad = present.new_zeros(present.shape[0], 1, *present.shape[2:])

            def repad(present, cu):
                pieces, start = [], 0
                for i in range(1, B + 1):
                    pieces += [present[:, start:int(cu[i])], pad]
                    start = int(cu[i])
                return torch.cat(pieces, dim=1)
            past = repad(present, cu)
            cu_q = torch.arange(B + 1, dtype=torch.int32)
            nxt = logits[(cu[1:] - 1).long()].float().argmax(-1)
            cu = cu + cu_q
            pos = torch.tensor(lens)
            for step in range(2):
                out[f"{name}_decode{step}_input"] = nxt.numpy()
                logits, present = model.forward(nxt, pos, cu, cu_q, max(lens) + 1 + step, None, past, None)
                out[f"{name}_decode{step}_logits"] = logits.numpy()
                past = repad(present, cu)
                cu = cu + cu_q
                pos = pos + 1
                nxt = logits.float().argmax(-1)
        out[f"{name}_input_ids"] = input_ids.numpy()
        out[f"{name}_lens"] = np.array(lens)
    np.savez(os.path.join(HERE, "flash_rw_ref.npz"), **out)
    print("flash_rw_ref.npz", sorted(out)[:4], "...")


def gold_proto():
    """field table of proto/generate.proto (message -> [name, number, type, label]) for tests/test_pb.py"""
    text = open("/root/reference/proto/generate.proto").read()
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"\{\s*\}", ";", text)  # empty messages / rpc bodies
    text = re.sub(r"message\s+(\w+)\s*;", r"message \1 { }", text)
    msgs = {}
    stack = []
    for m in re.finditer(r"(message|enum)\s+(\w+)\s*\{|\}|(?:(repeated|optional)\s+)?([\w.]+)\s+(\w+)\s*=\s*(\d+)\s*;|(\w+)\s*=\s*(\d+)\s*;|service\s+\w+\s*\{|rpc\s+(\w+)\s*\((\w+)\)\s*returns\s*\((\w+)\)", text):
        tok = m.group(0)
        if tok.startswith("message") or tok.startswith("enum"):
            stack.append((m.group(1), ".".join([s[1] for s in stack if s[0] == "message"] + [m.group(2)])))
            if m.group(1) == "message":
                msgs.setdefault(stack[-1][1], [])
        elif tok.startswith("service"):
            stack.append(("service", "service"))
        elif tok == "}":
            stack.pop()
        elif tok.startswith("rpc"):
            msgs.setdefault("__rpc__", []).append([m.group(9), m.group(10), m.group(11)])
        elif m.group(5):
            if stack and stack[-1][0] == "message":
                msgs[stack[-1][1]].append([m.group(5), int(m.group(6)), m.group(4), m.group(3) or ""])
        elif m.group(7) and stack and stack[-1][0] == "enum":
            msgs.setdefault("__enum__." + stack[-1][1], []).append([m.group(7), int(m.group(8))])
    with open(os.path.join(HERE, "generate_proto_fields.json"), "w") as f:
        json.dump(msgs, f, indent=1, sort_keys=True)
    print("generate_proto_fields.json", len(msgs))


def main():
    import tempfile
    my_pb = install_stubs()
    gold_proto()
    gold_gptq_pack()
    layers = _load("text_generation_server.utils.layers", "utils/layers.py")
    weights_mod = _load("text_generation_server.utils.weights", "utils/weights.py")
    gold_rope(layers)
    with tempfile.TemporaryDirectory() as tmp:
        gold_weights(weights_mod, tmp)
        gold_flash_llama(layers, weights_mod, tmp)
        gold_flash_neox(layers, weights_mod, tmp)
        gold_flash_santacoder(layers, weights_mod, tmp)
        gold_flash_rw(layers, weights_mod, tmp)
    gold_chooser(my_pb)
    gold_batch(my_pb)


if __name__ == "__main__" and "--neox-only" in sys.argv:
    import tempfile
    install_stubs()
    _layers = _load("text_generation_server.utils.layers", "utils/layers.py")
    _weights = _load("text_generation_server.utils.weights", "utils/weights.py")
    with tempfile.TemporaryDirectory() as _tmp:
        gold_flash_neox(_layers, _weights, _tmp)
elif __name__ == "__main__" and "--falcon-only" in sys.argv:
    import tempfile
    install_stubs()
    _layers = _load("text_generation_server.utils.layers", "utils/layers.py")
    _weights = _load("text_generation_server.utils.weights", "utils/weights.py")
    with tempfile.TemporaryDirectory() as _tmp:
        gold_flash_rw(_layers, _weights, _tmp)
elif __name__ == "__main__" and "--santacoder-only" in sys.argv:
    import tempfile
    install_stubs()
    _layers = _load("text_generation_server.utils.layers", "utils/layers.py")
    _weights = _load("text_generation_server.utils.weights", "utils/weights.py")
    with tempfile.TemporaryDirectory() as _tmp:
        gold_flash_santacoder(_layers, _weights, _tmp)
elif __name__ == "__main__" and "--batch-only" not in sys.argv:
    main()


# ----------------------------------------------------------------------------------------------------------
def gold_batch(my_pb):
    """FlashCausalLMBatch.from_pb / FlashCausalLM.generate_token / prune / concatenate of the REFERENCE, on CPU, with a
    stand-in `model.forward` (deterministic logits from the token ids; KV tensors of the reference's contiguous layout):
    pins the host-side batch bookkeeping (ids, positions, cu_seqlens, all_input_ids_tensor, lengths, chooser counters)
    through prefill, decode steps, an add-on batch + concatenate, and a prune.  -> batch_bookkeeping.npz"""
    import tgis_b200  # noqa: F401
    from tgis_b200.utils.synthetic import make_tokenizer
    for name, rel in (("dist", "utils/dist.py"), ("token_types", "utils/token_types.py"), ("layers", "utils/layers.py")):
        if f"text_generation_server.utils.{name}" not in sys.modules:
            _load(f"text_generation_server.utils.{name}", rel)
    _pkg("text_generation_server.inference_engine", os.path.join(REF, "inference_engine"))
    sys.modules["text_generation_server.inference_engine"].get_inference_engine_class = lambda name: None
    _load("text_generation_server.inference_engine.engine", "inference_engine/engine.py")
    _load("text_generation_server.prompt_cache", "prompt_cache.py")
    models = _pkg("text_generation_server.models", os.path.join(REF, "models"))
    _load("text_generation_server.models.types", "models/types.py")
    model_mod = _load("text_generation_server.models.model", "models/model.py")
    models.Model = model_mod.Model
    hub = types.ModuleType("text_generation_server.utils.hub")
    hub.get_model_path = lambda *a, **k: None
    sys.modules["text_generation_server.utils.hub"] = hub
    import transformers
    import transformers.generation.logits_process as tlp
    if not hasattr(transformers, "LogitsWarper"):
        transformers.LogitsWarper = transformers.LogitsProcessor
    if not hasattr(tlp, "LogitsWarper"):
        tlp.LogitsWarper = tlp.LogitsProcessor
    for name, rel in (("dist", "utils/dist.py"), ("token_types", "utils/token_types.py"), ("logits_process", "utils/logits_process.py"),
                      ("tokens", "utils/tokens.py"), ("layers", "utils/layers.py")):
        if f"text_generation_server.utils.{name}" not in sys.modules:
            _load(f"text_generation_server.utils.{name}", rel)
    fcl = _load("text_generation_server.models.flash_causal_lm", "models/flash_causal_lm.py")

    V = 64
    tok = make_tokenizer(V)

    def fake_logits(input_ids, position_ids):
        g = (input_ids.to(torch.int64) * 7919 + position_ids.to(torch.int64) * 104729) % 1000003
        base = torch.arange(V, dtype=torch.int64)[None, :]
        return (((g[:, None] * (base + 3)) % 97).float() / 9.7 - 5.0).to(torch.float16)

    class FakeModel:
        def forward(self, input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds, past_key_values, prealloc):
            T = input_ids.shape[0]
            if past_key_values is None:
                n = T if prealloc is None else prealloc
                present = torch.zeros(1, n, 2, 1, 1, dtype=torch.float16)
            else:
                present = past_key_values
            return fake_logits(input_ids, position_ids), present

    lm = object.__new__(fcl.FlashCausalLM)
    lm.model = FakeModel()
    lm.present_pad = None
    lm.device = torch.device("cpu")
    lm.tokenizer = tok

    def req(i, text, n_in, n_out, truncate=False):
        return my_pb.Request(id=i, inputs=text, input_length=n_in, truncate=truncate, max_output_length=n_out,
                             parameters=my_pb.NextTokenChooserParameters(temperature=0.0, top_p=1.0, min_new_tokens=2))

    out = {}

    def snap(tag, b, toks=None):
        out[f"{tag}_input_ids"] = b.input_ids.numpy().copy()
        out[f"{tag}_position_ids"] = b.position_ids.numpy().copy()
        out[f"{tag}_cu_seqlens"] = b.cu_seqlens.numpy().copy()
        out[f"{tag}_all_input_ids"] = b.all_input_ids_tensor.numpy().copy()
        out[f"{tag}_input_lengths"] = np.array(b.input_lengths)
        out[f"{tag}_total_lengths"] = np.array(list(b.total_lengths))
        out[f"{tag}_max_seqlen"] = np.array(b.max_seqlen)
        out[f"{tag}_request_ids"] = np.array([r.id for r in b.requests])
        out[f"{tag}_current_tokens"] = np.array(b.next_token_chooser.current_tokens)
        if toks is not None:
            out[f"{tag}_tokens"] = np.array([[t.request_id, t.token_id] for t in toks])

    msg_a = my_pb.Batch(id=0, requests=[req(0, "test <tok5> <tok9>", 3, 6), req(1, "test " * 50, 5, 6, truncate=True), req(2, "<tok7>", 1, 6)])
    msg_b = my_pb.Batch(id=1, requests=[req(3, "<tok11> test", 2, 4)])
    A, errs = fcl.FlashCausalLMBatch.from_pb(msg_a, tok, torch.float16, torch.device("cpu"), None, None, True)
    snap("a0", A)
    toks = lm.generate_token(A, first=True)[0]
    snap("a1", A, toks)
    for s in range(2):
        toks = lm.generate_token(A)[0]
        snap(f"a{2 + s}", A, toks)
    Bb, _ = fcl.FlashCausalLMBatch.from_pb(msg_b, tok, torch.float16, torch.device("cpu"), None, None, True)
    toks = lm.generate_token(Bb, first=True, for_concat=True)[0]
    snap("b1", Bb, toks)
    C = fcl.FlashCausalLMBatch.concatenate([A, Bb])
    snap("c0", C)
    toks = lm.generate_token(C)[0]
    snap("c1", C, toks)
    C = fcl.FlashCausalLMBatch.prune(C, [1])
    snap("p0", C)
    toks = lm.generate_token(C)[0]
    snap("p1", C, toks)
    out["msg_a"] = np.frombuffer(msg_a.SerializeToString(), dtype=np.uint8)
    out["msg_b"] = np.frombuffer(msg_b.SerializeToString(), dtype=np.uint8)
    np.savez(os.path.join(HERE, "batch_bookkeeping.npz"), **out)
    print("batch_bookkeeping.npz", len(out))


if __name__ == "__main__" and "--batch-only" in sys.argv:
    gold_batch(install_stubs() if "text_generation_server" not in sys.modules else sys.modules["text_generation_server.pb"].generate_pb2)
