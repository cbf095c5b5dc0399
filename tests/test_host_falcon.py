"""CPU tests of the Falcon / RefinedWeb host logic: the per-KV-group fused projection of the large form re-laid-out to
[q heads | k heads | v heads] (and cut by groups for tensor parallelism) must select the same rows as the reference's
`view(groups, heads_per_group + 2, d)` split (flash_rw_modeling.py:259-266, restated in oracle/falcon.py: split_qkv), and
`RWConfig.of` must read both config spellings."""
import os
import types

import pytest
import torch

from oracle import falcon as ofa


def test_grouped_qkv_relayout_and_group_sharding(tmp_path):
    from safetensors.torch import save_file
    import tgis_b200  # noqa: F401
    from tgis_b200.models.custom_modeling.flash_rw_modeling import RWConfig, load_grouped_qkv, load_row
    from tgis_b200.utils.dist import FakeGroup
    from tgis_b200.utils.layers import FastLinear
    from tgis_b200.utils.weights import Weights

    cfg = ofa.FalconConfig(128, 1, 8, 4, 96, new_decoder_architecture=True, parallel_attn=True, bias=True)
    sd = ofa.make_state_dict(cfg, seed=5)
    path = os.path.join(str(tmp_path), "model.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    H, d, hpg = 128, 16, 2
    w, b = sd["transformer.h.0.self_attention.query_key_value.weight"], sd["transformer.h.0.self_attention.query_key_value.bias"]
    # the reference's split applied to the ROWS of the weight (each output feature is one row): which rows are q / k / v
    q_rows, k_rows, v_rows = ofa.split_qkv(cfg, torch.arange(w.shape[0])[None].float())
    q_rows, k_rows, v_rows = q_rows[0].long(), k_rows[0].long(), v_rows[0].long()  # [h, d], [kv, d], [kv, d]
    rw = RWConfig(model_type="RefinedWeb", hidden_size=H, num_attention_heads=8, num_kv_heads=4, bias=True, parallel_attn=True)
    dense_w = sd["transformer.h.0.self_attention.dense.weight"]
    for world in (1, 2, 4):
        for rank in range(world):
            weights = Weights([path], device="cpu", dtype=torch.float16, process_group=FakeGroup(rank, world))
            gl = 4 // world
            lin = load_grouped_qkv(rw, "transformer.h.0.self_attention.query_key_value", weights, gl, hpg, d, H).linear
            groups = slice(rank * gl, (rank + 1) * gl)
            want = torch.cat([q_rows.view(4, hpg, d)[groups].reshape(-1), k_rows[groups].reshape(-1), v_rows[groups].reshape(-1)])
            assert torch.equal(lin.weight, w[want]) and torch.equal(lin.bias, b[want])
            row = load_row(rw, "transformer.h.0.self_attention.dense", weights, bias=True)
            assert isinstance(row, FastLinear)  # parallel_attn: bare linear, the layer all-reduces once (:29-32)
            assert torch.equal(row.weight, dense_w[:, rank * gl * hpg * d:(rank + 1) * gl * hpg * d])
            assert (row.bias is not None) == (rank == 0)


def test_rwconfig_reads_both_spellings():
    import tgis_b200  # noqa: F401
    from tgis_b200.models.custom_modeling.flash_rw_modeling import FlashRWForCausalLM, RWConfig
    from transformers import FalconConfig
    # Falcon-7B as transformers spells it: multi_query, num_kv_heads left at its default (= heads)
    seven = RWConfig.of(FalconConfig(hidden_size=4544, num_attention_heads=71, num_hidden_layers=32, multi_query=True,
                                     new_decoder_architecture=False, parallel_attn=True, bias=False))
    assert (seven.n_head, seven.n_head_kv, seven.new_decoder_architecture, seven.parallel_attn) == (71, 1, False, True)
    assert FlashRWForCausalLM.kv_cache_layout(seven, 1) == (1, 1)
    # Falcon-40B
    forty = RWConfig.of(FalconConfig(hidden_size=8192, num_attention_heads=128, num_kv_heads=8, num_hidden_layers=60,
                                     new_decoder_architecture=True, parallel_attn=True))
    assert (forty.n_head, forty.n_head_kv, forty.new_decoder_architecture) == (128, 8, True)
    assert FlashRWForCausalLM.kv_cache_layout(forty, 4) == (8, 4)
    assert forty.num_hidden_layers == 60 and forty.num_attention_heads == 128
    # the original RefinedWeb checkpoints' spelling (flash_rw_modeling.py:36-118)
    old = RWConfig.of(types.SimpleNamespace(model_type="RefinedWebModel", n_head=71, n_head_kv=1, n_layer=32, hidden_size=4544,
                                            vocab_size=65024, parallel_attn=True, bias=False, multi_query=True))
    assert (old.n_head, old.n_head_kv, old.n_layer, old.new_decoder_architecture) == (71, 1, 32, False)
    assert RWConfig(model_type="RefinedWeb", n_head=128, n_head_kv=8).new_decoder_architecture
    with pytest.raises(NotImplementedError):
        RWConfig(alibi=True)


def test_engine_registers_the_family():
    import tgis_b200  # noqa: F401
    from tgis_b200 import inference_engine
    assert {"falcon", "RefinedWeb", "RefinedWebModel"} <= set(inference_engine.FLASH_TYPES)
