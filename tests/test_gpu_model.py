"""GPU parity: FlashLlamaForCausalLM (C++ step runtime over the CUDA kernels) against the CPU oracle's restatement of
the reference graph, on synthetic checkpoints — prefill logits, then greedy decode steps with the paged KV cache.

Tolerance: logits within 1e-3 * max|logit| + 1 fp16 ulp per element per step (both sides round to fp16 at the same
points; only accumulation order differs) — loosened by the number of layers through which rounding differences
propagate: rel 4e-3.  Greedy ids must match wherever the oracle's top-2 gap exceeds 2 fp16 ulp.
"""
import os
import types

import pytest
import torch

from oracle import llama as oll

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _config(cfg: oll.LlamaConfig, quantize):
    return types.SimpleNamespace(
        hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_hidden_layers,
        num_attention_heads=cfg.num_attention_heads, num_key_value_heads=cfg.num_key_value_heads, vocab_size=cfg.vocab_size,
        rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta, rope_scaling=None, quantize=quantize, hidden_act="silu",
        attention_bias=False, mlp_bias=False, max_position_embeddings=512, model_type="llama")


def build_model(tmp_path, cfg, quantize, seed=1234, act_order=False):
    from safetensors.torch import save_file
    import tgis_b200  # noqa: F401
    from tgis_b200.models.custom_modeling.flash_llama_modeling import FlashLlamaForCausalLM
    from tgis_b200.utils.dist import FakeGroup
    from tgis_b200.utils.paged import PagedKVCacheManager
    from tgis_b200.utils.weights import Weights

    sd = oll.make_state_dict(cfg, seed=seed, quantize=quantize, act_order=act_order)
    path = os.path.join(tmp_path, "model.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    weights = Weights([path], device=DEV, dtype=torch.float16, process_group=FakeGroup(0, 1))
    model = FlashLlamaForCausalLM(_config(cfg, quantize), weights)
    model.kv_cache_manager = PagedKVCacheManager(cfg.num_hidden_layers, cfg.num_attention_heads, cfg.hidden_size,
                                                 kv_heads=cfg.num_key_value_heads, device=DEV, total_num_gpu_blocks=256)
    oracle = oll.LlamaOracle(oll.build_shards(cfg, sd, 1))
    return model, oracle


def _check_logits(got, ref, what, rel=4e-3):
    got, ref = got.float().cpu(), ref.float()
    assert torch.isfinite(got).all(), f"{what}: non-finite logits"
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= rel * scale + 2e-3, f"{what}: max logit err {err:.4e} vs scale {scale:.3e}"


def _check_greedy(got_ids, ref_logits, what):
    ref = ref_logits.float()
    top2 = ref.topk(2, dim=-1)
    gap = top2.values[:, 0] - top2.values[:, 1]
    ulp = torch.maximum(top2.values[:, 0].abs(), torch.tensor(1.0)) * 2.0 ** -10
    decisive = gap > 2 * ulp
    ref_ids = top2.indices[:, 0]
    bad = decisive & (got_ids.cpu() != ref_ids)
    assert not bad.any(), f"{what}: greedy ids differ outside the tie band at rows {torch.nonzero(bad).flatten().tolist()}"
    return decisive


CASES = [
    ("mha_d128_fp16", oll.LlamaConfig(256, 512, 2, 2, 2, 512), None),
    ("gqa_d64_fp16", oll.LlamaConfig(512, 1024, 3, 8, 2, 1000), None),
    ("mha_d128_gptq", oll.LlamaConfig(256, 512, 2, 2, 2, 512), "gptq"),
    ("gqa_d128_gptq", oll.LlamaConfig(1024, 2048, 2, 8, 2, 768), "gptq"),
    ("gqa_d128_gptq_actorder", oll.LlamaConfig(1024, 2048, 2, 8, 2, 768), "gptq"),
]


@pytest.mark.parametrize("name,cfg,quantize", CASES, ids=[c[0] for c in CASES])
def test_prefill_then_decode_matches_oracle(tmp_path, name, cfg, quantize):
    model, oracle = build_model(str(tmp_path), cfg, quantize, act_order=name.endswith("actorder"))
    mgr = model.kv_cache_manager
    g = torch.Generator().manual_seed(7)
    lens = [5, 17, 1, 40, 16]
    prompts = [torch.randint(0, cfg.vocab_size, (L,), generator=g).tolist() for L in lens]
    n_new = 6
    ref_tokens, ref_logits = oracle.generate_greedy(prompts, n_new)

    from tgis_b200.utils.paged import PagedKVState
    B = len(prompts)
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    sids = mgr.allocate_tokens(lens, reserve_tokens=[n_new] * B)
    kv = PagedKVState(sequence_ids=sids, block_table=mgr.block_table_tensor(sids),
                      context_lens=torch.tensor(lens, dtype=torch.int32, device=DEV),
                      slot_mapping=mgr.slot_mapping_for(sids, [0] * B, lens), max_blocks=0)
    input_ids = torch.tensor([t for p in prompts for t in p], dtype=torch.int64, device=DEV)
    position_ids = torch.cat([torch.arange(L) for L in lens]).to(DEV)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    last = (cu_t[1:] - 1).to(torch.int64)
    logits, _ = model.forward(input_ids, position_ids, cu_t, None, max(lens), None, kv, None, last)
    torch.cuda.synchronize()
    _check_logits(logits, ref_logits[0], f"{name} prefill")
    # feed the ORACLE's tokens so every step is compared on identical inputs
    cur = list(lens)
    from tgis_b200 import ops
    for step in range(1, n_new):
        nxt = ref_tokens[:, step - 1].to(DEV)
        pos = torch.tensor(cur, dtype=torch.int64, device=DEV)
        kv.slot_mapping = mgr.slot_mapping_for(sids, cur, [1] * B)
        cur = [c + 1 for c in cur]
        kv.context_lens = torch.tensor(cur, dtype=torch.int32, device=DEV)
        logits, _ = model.forward(nxt, pos, torch.arange(B + 1, dtype=torch.int32, device=DEV),
                                  torch.arange(B + 1, dtype=torch.int32, device=DEV), max(cur), None, kv, None, None)
        torch.cuda.synchronize()
        _check_logits(logits, ref_logits[step], f"{name} decode step {step}")
        _check_greedy(ops.argmax(logits), ref_logits[step], f"{name} decode step {step}")
    mgr.free_sequences(sids)
    assert mgr.free_blocks == mgr.total_num_gpu_blocks
