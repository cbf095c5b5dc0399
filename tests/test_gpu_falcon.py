"""Flash Falcon / RefinedWeb on the B200 kernels vs the CPU oracle (oracle/falcon.py, itself pinned against the reference's
own module graph and transformers): prefill + decode logits per step and greedy ids outside the fp16 tie band, for the
multi-query parallel form (incl. 20 query heads on one KV head: more than one decode launch shares), the sequential form
with biases, and the grouped large form (two LayerNorms, per-group fused projection re-laid-out at load)."""
import os
import types

import pytest
import torch

from oracle import falcon as ofa

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _config(cfg: ofa.FalconConfig):
    return types.SimpleNamespace(
        model_type="RefinedWeb" if cfg.new_decoder_architecture else "RefinedWebModel", hidden_size=cfg.hidden_size,
        n_layer=cfg.num_hidden_layers, n_head=cfg.n_head, n_head_kv=cfg.n_head_kv, vocab_size=cfg.vocab_size,
        new_decoder_architecture=cfg.new_decoder_architecture, parallel_attn=cfg.parallel_attn, bias=cfg.bias,
        layer_norm_epsilon=cfg.layer_norm_epsilon, multi_query=True, alibi=False, quantize=None, max_position_embeddings=512)


def build(tmp_path, cfg, seed=11):
    from safetensors.torch import save_file
    import tgis_b200  # noqa: F401
    from tgis_b200.models.custom_modeling.flash_rw_modeling import FlashRWForCausalLM
    from tgis_b200.utils.dist import FakeGroup
    from tgis_b200.utils.paged import PagedKVCacheManager
    from tgis_b200.utils.weights import Weights

    sd = ofa.make_state_dict(cfg, seed=seed, std=0.04)
    path = os.path.join(tmp_path, "model.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    weights = Weights([path], device=DEV, dtype=torch.float16, process_group=FakeGroup(0, 1))
    model = FlashRWForCausalLM(_config(cfg), weights)
    model.kv_cache_manager = PagedKVCacheManager(cfg.num_hidden_layers, cfg.n_head, cfg.hidden_size, kv_heads=cfg.n_head_kv,
                                                 device=DEV, total_num_gpu_blocks=128)
    return model, ofa.FalconOracle(cfg, sd)


def _check_logits(got, ref, what, rel=4e-3):
    got, ref = got.float().cpu(), ref.float()
    assert torch.isfinite(got).all(), f"{what}: non-finite logits"
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= rel * scale + 2e-3, f"{what}: max logit err {err:.4e} vs scale {scale:.3e}"


CASES = [
    ("mqa_parallel_h4_d64", ofa.FalconConfig(256, 2, 4, 1, 512, new_decoder_architecture=False, parallel_attn=True)),
    ("mqa_parallel_h20_d64", ofa.FalconConfig(1280, 2, 20, 1, 384, new_decoder_architecture=False, parallel_attn=True)),
    ("mqa_sequential_bias_d128", ofa.FalconConfig(256, 2, 2, 1, 512, new_decoder_architecture=False, parallel_attn=False, bias=True)),
    ("gqa_large_h8_kv2_d64", ofa.FalconConfig(512, 2, 8, 2, 384, new_decoder_architecture=True, parallel_attn=True)),
]


@pytest.mark.parametrize("name,cfg", CASES, ids=[c[0] for c in CASES])
def test_falcon_prefill_then_decode_matches_oracle(tmp_path, name, cfg):
    from tgis_b200 import ops
    from tgis_b200.utils.paged import PagedKVState
    model, oracle = build(str(tmp_path), cfg)
    mgr = model.kv_cache_manager
    g = torch.Generator().manual_seed(7)
    lens = [5, 17, 1, 33, 16]
    prompts = [torch.randint(0, cfg.vocab_size, (L,), generator=g).tolist() for L in lens]
    n_new = 5
    ref_tokens, ref_logits = oracle.generate_greedy(prompts, n_new)
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
    with torch.inference_mode():
        logits, _ = model.forward(input_ids, position_ids, cu_t, None, max(lens), None, kv, None, last)
        torch.cuda.synchronize()
        _check_logits(logits, ref_logits[0], f"{name} prefill")
        cur = list(lens)
        for step in range(1, n_new):
            nxt = ref_tokens[:, step - 1].to(DEV)  # the ORACLE's tokens: every step is compared on identical inputs
            pos = torch.tensor(cur, dtype=torch.int64, device=DEV)
            kv.slot_mapping = mgr.slot_mapping_for(sids, cur, [1] * B)
            cur = [c + 1 for c in cur]
            kv.context_lens = torch.tensor(cur, dtype=torch.int32, device=DEV)
            ar = torch.arange(B + 1, dtype=torch.int32, device=DEV)
            logits, _ = model.forward(nxt, pos, ar, ar, max(cur), None, kv, None, None)
            torch.cuda.synchronize()
            _check_logits(logits, ref_logits[step], f"{name} decode step {step}")
            ref = ref_logits[step].float()
            top2 = ref.topk(2, dim=-1)
            decisive = (top2.values[:, 0] - top2.values[:, 1]) > 2 * torch.maximum(top2.values[:, 0].abs(), torch.tensor(1.0)) * 2.0 ** -10
            bad = decisive & (ops.argmax(logits).cpu() != top2.indices[:, 0])
            assert not bad.any(), f"{name} step {step}: greedy ids differ outside the tie band"
    mgr.free_sequences(sids)
    assert mgr.free_blocks == mgr.total_num_gpu_blocks
