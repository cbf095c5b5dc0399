"""CPU tests of the GPT-NeoX host logic: tensor-parallel slicing and the [h, 3, d] -> [3, h, d] QKV re-layout of
flash_neox_modeling.py:40-80 through this repo's `Weights` / `load_qkv` / `load_row`, checked against plain indexing."""
import os
import types

import torch

from oracle import neox as onx


def _cfg(parallel=True):
    return types.SimpleNamespace(quantize=None, use_parallel_residual=parallel)


def test_neox_qkv_relayout_and_tp_slicing(tmp_path):
    from safetensors.torch import save_file
    import tgis_b200  # noqa: F401
    from tgis_b200.models.custom_modeling.flash_neox_modeling import load_qkv, load_row
    from tgis_b200.utils.dist import FakeGroup
    from tgis_b200.utils.layers import FastLinear, TensorParallelColumnLinear, TensorParallelRowLinear
    from tgis_b200.utils.weights import Weights

    cfg = onx.NeoXConfig(64, 256, 1, 4, 96)
    sd = onx.make_state_dict(cfg, seed=5)
    path = os.path.join(str(tmp_path), "model.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    h, d, H = 4, 16, 64
    full_w = sd["gpt_neox.layers.0.attention.query_key_value.weight"].view(h, 3, d, H)
    full_b = sd["gpt_neox.layers.0.attention.query_key_value.bias"].view(h, 3, d)
    dense_w = sd["gpt_neox.layers.0.attention.dense.weight"]
    for world in (1, 2):
        outs = []
        for rank in range(world):
            weights = Weights([path], device="cpu", dtype=torch.float16, process_group=FakeGroup(rank, world))
            hl = h // world
            lin = load_qkv(_cfg(True), "gpt_neox.layers.0.attention.query_key_value", weights, hl, d, H)
            assert isinstance(lin, FastLinear)  # parallel residual: bare linear, the layer all-reduces once (:255-260)
            heads = full_w[rank * hl:(rank + 1) * hl]  # this rank's heads, still interleaved [hl, 3, d, H]
            assert torch.equal(lin.weight, heads.permute(1, 0, 2, 3).reshape(3 * hl * d, H))
            assert torch.equal(lin.bias, full_b[rank * hl:(rank + 1) * hl].permute(1, 0, 2).reshape(-1))
            row = load_row(_cfg(True), "gpt_neox.layers.0.attention.dense", weights, bias=True)
            assert torch.equal(row.weight, dense_w[:, rank * hl * d:(rank + 1) * hl * d])
            assert (row.bias is not None) == (rank == 0)  # bias on rank 0 only (:43-47)
            outs.append(row)
            # sequential residual: the TP wrappers that all-reduce themselves
            assert isinstance(load_qkv(_cfg(False), "gpt_neox.layers.0.attention.query_key_value", weights, hl, d, H),
                              TensorParallelColumnLinear)
            assert isinstance(load_row(_cfg(False), "gpt_neox.layers.0.attention.dense", weights, bias=True), TensorParallelRowLinear)
        # the row shards tile the full matrix
        assert torch.equal(torch.cat([o.weight for o in outs], dim=1), dense_w)


def test_neox_engine_registers_the_family():
    import tgis_b200  # noqa: F401
    from tgis_b200 import inference_engine
    assert "gpt_neox" in inference_engine.FLASH_TYPES and "llama" in inference_engine.FLASH_TYPES


def test_synthetic_neox_checkpoint_serves_every_tensor_the_model_loads():
    """bench.py's GPT-NeoX workload builds the model from utils/synthetic.py: every name the loader asks for exists, at tp 1 and 2"""
    import torch
    import tgis_b200  # noqa: F401
    from tgis_b200.models.custom_modeling.flash_neox_modeling import FlashGPTNeoXForCausalLM
    from tgis_b200.utils.dist import FakeGroup
    from tgis_b200.utils.synthetic import SyntheticWeights, model_config
    cfg = model_config("tiny-neox", quantize=None, max_position_embeddings=512)
    sizes = []
    for world in (1, 2):
        model = FlashGPTNeoXForCausalLM(cfg, SyntheticWeights(cfg, "cpu", torch.float16, FakeGroup(0, world)))
        qkv = model.model.layers[0].attention.query_key_value
        sizes.append(tuple(qkv.weight.shape))
        inv = model.model.layers[0].attention.rotary_emb.inv_freq
        assert inv.dtype == torch.float32 and inv.shape[0] == int(64 * 0.25) // 2
    assert sizes == [(768, 256), (384, 256)]
