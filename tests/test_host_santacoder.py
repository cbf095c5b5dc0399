"""CPU tests of the Santacoder (gpt_bigcode, multi-query) host logic: how the fused `c_attn` projection is cut for tensor
parallelism — this rank's block of query rows, then the shared key / value rows every rank keeps
(flash_santacoder_modeling.py:19-159) — through this repo's `Weights` / `load_multi_mqa` / `load_col` / `load_row`, checked
against plain indexing, for plain, transposed (GPT2-style Conv1D) and GPTQ tensors."""
import os
import types

import pytest
import torch

from oracle import gptq as ogptq
from oracle import santacoder as osc


def _cfg(transpose=False, quantize=None):
    return types.SimpleNamespace(quantize=quantize, transpose=transpose)


@pytest.fixture()
def checkpoint(tmp_path):
    from safetensors.torch import save_file
    cfg = osc.SantacoderConfig(64, 256, 1, 4, 96, n_positions=32)
    sd = osc.make_state_dict(cfg, seed=5)
    path = os.path.join(str(tmp_path), "model.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    return cfg, sd, path


def test_c_attn_is_cut_into_query_block_plus_shared_kv(checkpoint):
    import tgis_b200  # noqa: F401
    from tgis_b200.models.custom_modeling.flash_santacoder_modeling import load_col, load_multi_mqa, load_row
    from tgis_b200.utils.dist import FakeGroup
    from tgis_b200.utils.layers import TensorParallelColumnLinear, TensorParallelRowLinear
    from tgis_b200.utils.weights import Weights

    cfg, sd, path = checkpoint
    h, d, H = 4, 16, 64
    w, b = sd["transformer.h.0.attn.c_attn.weight"], sd["transformer.h.0.attn.c_attn.bias"]
    proj_w, fc_w = sd["transformer.h.0.attn.c_proj.weight"], sd["transformer.h.0.mlp.c_fc.weight"]
    for world in (1, 2, 4):
        rows, cols = [], []
        for rank in range(world):
            weights = Weights([path], device="cpu", dtype=torch.float16, process_group=FakeGroup(rank, world))
            hl = h // world
            lin = load_multi_mqa(_cfg(), "transformer.h.0.attn", weights, True, d, hl, H)
            assert isinstance(lin, TensorParallelColumnLinear)
            q_rows = slice(rank * hl * d, (rank + 1) * hl * d)
            assert torch.equal(lin.linear.weight, torch.cat([w[q_rows], w[h * d:]]))     # [this rank's heads | k | v]
            assert torch.equal(lin.linear.bias, torch.cat([b[q_rows], b[h * d:]]))
            assert lin.linear.weight.shape == ((hl + 2) * d, H)
            row = load_row(_cfg(), "transformer.h.0.attn.c_proj", weights, bias=True)
            assert isinstance(row, TensorParallelRowLinear)
            assert torch.equal(row.linear.weight, proj_w[:, q_rows])
            assert (row.linear.bias is not None) == (rank == 0)                           # added once, on the first rank (:183-187)
            col = load_col(_cfg(), "transformer.h.0.mlp.c_fc", weights, bias=True)
            rows.append(row.linear.weight)
            cols.append(col.linear.weight)
        assert torch.equal(torch.cat(rows, dim=1), proj_w) and torch.equal(torch.cat(cols, dim=0), fc_w)


def test_transposed_gpt2_style_checkpoint(tmp_path, checkpoint):
    """GPT2-architecture checkpoints keep Conv1D weights [in, out] (`config.transpose`, tgis_native.py:84)"""
    from safetensors.torch import save_file
    import tgis_b200  # noqa: F401
    from tgis_b200.models.custom_modeling.flash_santacoder_modeling import load_col, load_multi_mqa, load_row
    from tgis_b200.utils.dist import FakeGroup
    from tgis_b200.utils.weights import Weights

    cfg, sd, _ = checkpoint
    h, d, H = 4, 16, 64
    flipped = {k: (v.t().contiguous() if k.endswith(("c_attn.weight", "c_proj.weight", "c_fc.weight")) else v) for k, v in sd.items()}
    path = os.path.join(str(tmp_path), "gpt2_style.safetensors")
    save_file(flipped, path)
    w = sd["transformer.h.0.attn.c_attn.weight"]
    for rank in range(2):
        weights = Weights([path], device="cpu", dtype=torch.float16, process_group=FakeGroup(rank, 2))
        lin = load_multi_mqa(_cfg(transpose=True), "transformer.h.0.attn", weights, True, d, 2, H)
        q_rows = slice(rank * 2 * d, (rank + 1) * 2 * d)
        assert torch.equal(lin.linear.weight, torch.cat([w[q_rows], w[h * d:]]))
        assert torch.equal(load_row(_cfg(transpose=True), "transformer.h.0.attn.c_proj", weights, True).linear.weight,
                           sd["transformer.h.0.attn.c_proj.weight"][:, q_rows])
        assert torch.equal(load_col(_cfg(transpose=True), "transformer.h.0.mlp.c_fc", weights, True).linear.weight,
                           sd["transformer.h.0.mlp.c_fc.weight"][rank * 128:(rank + 1) * 128])


def test_gptq_c_attn_columns(tmp_path):
    """GPTQ tensors are [in, out]: the query block and the shared kv tail are column ranges, 8 to an int32 in qzeros"""
    from safetensors.torch import save_file
    import tgis_b200  # noqa: F401
    from tgis_b200.models.custom_modeling.flash_santacoder_modeling import _q_block_then_kv
    from tgis_b200.utils.dist import FakeGroup
    from tgis_b200.utils.weights import Weights

    h, d, H = 4, 32, 128
    g = torch.Generator().manual_seed(1)
    dense = (torch.randn((h + 2) * d, H, generator=g) * 0.05).half()
    qweight, qzeros, scales, _ = ogptq.quantize_rtn(dense, groupsize=64)
    q = {"qweight": qweight, "qzeros": qzeros, "scales": scales}
    path = os.path.join(str(tmp_path), "q.safetensors")
    save_file({"a.c_attn.qweight": qweight, "a.c_attn.qzeros": qzeros, "a.c_attn.scales": scales}, path)
    kv = 2 * d
    for world in (1, 2):
        for rank in range(world):
            weights = Weights([path], device="cpu", dtype=torch.float16, process_group=FakeGroup(rank, world))
            block = h * d // world
            cols = list(range(rank * block, (rank + 1) * block)) + list(range(h * d, (h + 2) * d))
            qw = _q_block_then_kv(weights._get_slice("a.c_attn.qweight"), 1, kv, world, rank)
            sc = _q_block_then_kv(weights._get_slice("a.c_attn.scales"), 1, kv, world, rank)
            qz = _q_block_then_kv(weights._get_slice("a.c_attn.qzeros"), 1, kv // 8, world, rank)
            assert torch.equal(qw, q["qweight"][:, cols]) and torch.equal(sc, q["scales"][:, cols])
            assert torch.equal(qz, q["qzeros"][:, [c // 8 for c in cols[::8]]])
            # the cut tensors dequantize to the same rows of the dense matrix the full tensors dequantize to
            full = ogptq.dequantize(q["qweight"], q["qzeros"], q["scales"], None, 64)
            part = ogptq.dequantize(qw, qz, sc, None, 64)
            assert torch.equal(part, full[:, cols])


def test_engine_registers_the_family():
    import tgis_b200  # noqa: F401
    from tgis_b200 import inference_engine
    assert "gpt_bigcode" in inference_engine.FLASH_TYPES
