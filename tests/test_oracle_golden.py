"""CPU tests: the oracle (and the host-side ports) against the golden fixtures produced by running the reference's own
Python (tests/golden/make_golden.py).  These pin the oracle before it is trusted as the checker of the CUDA path."""
import os

import numpy as np
import pytest
import torch

from oracle import gptq as ogptq
from oracle import llama as oll

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _npz(name):
    return np.load(os.path.join(G, name))


# ------------------------------------------------------------------------------------------------ GPTQ
def test_gptq_pack_layout_matches_reference_pack():
    """reference QuantLinear.pack (quant_linear.py:290-345) -> our unpack recovers exactly what pack rounded."""
    z = _npz("gptq_pack.npz")
    gs = int(z["groupsize"])
    w = torch.from_numpy(z["weight"])            # [N, K]
    scales_in = torch.from_numpy(z["scales_in"])  # [N, G]
    zeros_in = torch.from_numpy(z["zeros_in"])
    N, K = w.shape
    # what pack computes (quant_linear.py:300-311): round((w + zero*scale) / scale) with the fp16-rounded scale
    s16 = scales_in.t().contiguous().half()       # [G, N]
    sz = (zeros_in.t() * scales_in.t())
    g_idx = torch.arange(K) // gs
    expect_q = torch.round((w.t() + sz[g_idx]) / s16[g_idx]).to(torch.int64).numpy().astype(np.uint32) & 0xF
    q = ogptq.unpack_rows_int4(z["qweight"])
    assert q.shape == (K, N)
    assert np.array_equal(q.astype(np.uint32), expect_q)
    zs = ogptq.unpack_cols_int4(z["qzeros"])
    assert np.array_equal(zs.astype(np.int64) + 1, zeros_in.t().numpy().astype(np.int64))  # stored minus one (:329)
    # and our packers are the exact inverse
    assert np.array_equal(ogptq.pack_rows_int4(q), z["qweight"])
    assert np.array_equal(ogptq.pack_cols_int4(zs), z["qzeros"])
    assert np.array_equal(z["g_idx"], g_idx.numpy())


def test_gptq_dequant_formula_of_record():
    """W = fp16(scale * (q - (z + 1))) (quant_linear.py:184-192) reproduces the packed weights to quantisation error."""
    z = _npz("gptq_pack.npz")
    gs = int(z["groupsize"])
    wd = ogptq.dequantize(z["qweight"], z["qzeros"], torch.from_numpy(z["scales"]), None, gs)  # [K, N]
    w = torch.from_numpy(z["weight"]).t()
    scale = torch.from_numpy(z["scales"]).float()[torch.arange(w.shape[0]) // gs]
    assert ((wd.float() - w).abs() <= 0.5 * scale + 1e-3).all()
    # explicit g_idx == implicit k // groupsize
    wd2 = ogptq.dequantize(z["qweight"], z["qzeros"], torch.from_numpy(z["scales"]), torch.from_numpy(z["g_idx"]), gs)
    assert torch.equal(wd, wd2)


def test_quantize_rtn_roundtrip():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(64, 256, generator=g) * 0.05
    qw, qz, sc, gi = ogptq.quantize_rtn(w, 128)
    assert qw.shape == (32, 64) and qz.shape == (2, 8) and sc.shape == (2, 64) and gi.shape == (256,)
    wd = ogptq.dequantize(qw, qz, sc, gi, 128)
    assert ((wd.float() - w.t()).abs() <= 0.51 * sc.float()[gi.long()] + 1e-3).all()


# ------------------------------------------------------------------------------------------------ RoPE
@pytest.mark.parametrize("name,d,theta,factor", [("d64", 64, 10000.0, 1.0), ("d128", 128, 10000.0, 1.0),
                                                   ("d128_theta5e5", 128, 500000.0, 1.0), ("d64_linear2", 64, 10000.0, 2.0)])
def test_rope_tables_match_reference(name, d, theta, factor):
    z = _npz("rope_tables.npz")
    pos = torch.from_numpy(z["positions"])
    cos, sin = oll.rope_tables(d, theta, 1024, factor)
    assert np.array_equal(cos[pos].numpy(), z[f"{name}_cos"])
    assert np.array_equal(sin[pos].numpy(), z[f"{name}_sin"])


def test_product_rope_tables_match_reference():
    import tgis_b200  # noqa: F401
    from tgis_b200.utils.layers import LinearScalingPositionRotaryEmbedding, PositionRotaryEmbedding
    z = _npz("rope_tables.npz")
    pos = torch.from_numpy(z["positions"])
    rot = PositionRotaryEmbedding.static(128, 10000.0, "cpu")
    cos, sin = rot.get_cos_sin(pos, 1024, torch.float16)
    assert np.array_equal(cos.squeeze(1).numpy(), z["d128_cos"]) and np.array_equal(sin.squeeze(1).numpy(), z["d128_sin"])
    rot = LinearScalingPositionRotaryEmbedding.static(64, 10000.0, 2.0, "cpu")
    cos, sin = rot.get_cos_sin(pos, 1024, torch.float16)
    assert np.array_equal(cos.squeeze(1).numpy(), z["d64_linear2_cos"])


# ------------------------------------------------------------------------------------------------ TP slicing
@pytest.mark.parametrize("quant", [None, "gptq"])
def test_tp_shards_match_reference_weights_loader(quant, tmp_path):
    """oracle.build_shards and the product's Weights port both reproduce the reference Weights slicing at tp = 2."""
    from safetensors.torch import save_file
    import tgis_b200  # noqa: F401
    from tgis_b200.utils.weights import Weights

    z = _npz("weights_shards.npz")
    cfg = oll.LlamaConfig(256, 512, 1, 4, 2, 512)
    sd = oll.make_state_dict(cfg, seed=5, quantize=quant)
    shards = oll.build_shards(cfg, sd, 2)
    path = os.path.join(str(tmp_path), "w.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)

    class Group:
        def __init__(self, r):
            self.r = r

        def rank(self):
            return self.r

        def size(self):
            return 2

    for rank in (0, 1):
        tag = f"{quant or 'fp16'}_r{rank}"
        L = shards[rank].layers[0]
        w = Weights([path], device="cpu", dtype=torch.float16, process_group=Group(rank))
        p = "model.layers.0"
        col = w.get_multi_weights_col([f"{p}.self_attn.q_proj", f"{p}.self_attn.k_proj", f"{p}.self_attn.v_proj"], quant, 0)
        row = w.get_multi_weights_row(f"{p}.mlp.down_proj", quant)
        emb = w.get_partial_sharded("model.embed_tokens.weight", dim=0)
        assert np.array_equal(emb.numpy(), z[f"{tag}_emb"]) and np.array_equal(shards[rank].embed.numpy(), z[f"{tag}_emb"])
        if quant == "gptq":
            for i, nm in enumerate(("qweight", "qzeros", "scales")):
                assert np.array_equal(col[i].numpy(), z[f"{tag}_col_{nm}"])
                assert np.array_equal(getattr(L.qkv, nm).numpy(), z[f"{tag}_col_{nm}"])
                assert np.array_equal(row[i].numpy(), z[f"{tag}_row_{nm}"])
                assert np.array_equal(getattr(L.down, nm).numpy(), z[f"{tag}_row_{nm}"])
            assert bool(z[f"{tag}_row_gidx_is_none"]) == (row[3] is None) == (L.down.g_idx is None)
        else:
            assert np.array_equal(col.numpy(), z[f"{tag}_col"]) and np.array_equal(L.qkv.weight.numpy(), z[f"{tag}_col"])
            assert np.array_equal(row.numpy(), z[f"{tag}_row"]) and np.array_equal(L.down.weight.numpy(), z[f"{tag}_row"])


# ------------------------------------------------------------------------------------------------ model graph
@pytest.mark.parametrize("name", ["mha", "gqa"])
def test_oracle_graph_matches_reference_flash_llama(name):
    """The reference's FlashLlamaForCausalLM.forward (its own Python, third-party kernels shimmed) vs LlamaOracle:
    prefill logits for every prompt token and two decode steps, tensor by tensor.  Both sides share the kernel
    restatements, so the wiring (qkv split, RoPE order, KV placement, gate/up order, residual flow, head) must agree
    exactly up to CPU fp16 GEMM rounding: <= 2 fp16 ulp of the logit scale."""
    z = _npz("flash_llama_ref.npz")
    H, I, nl, h, kv, V = [int(x) for x in z[f"{name}_cfg"]]
    cfg = oll.LlamaConfig(H, I, nl, h, kv, V)
    sd = oll.make_state_dict(cfg, seed=31, std=0.08)
    orc = oll.LlamaOracle(oll.build_shards(cfg, sd, 1))
    lens = [int(x) for x in z[f"{name}_lens"]]
    ids = torch.from_numpy(z[f"{name}_input_ids"])
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    pos = torch.cat([torch.arange(L) for L in lens])
    logits, kvs = orc.forward(ids, pos, cu, None, prefill=True, all_logits=True)
    ref = torch.from_numpy(z[f"{name}_prefill_logits"])
    tol = 2 * 2.0 ** -10 * max(1.0, ref.float().abs().max().item())
    assert (logits.float() - ref.float()).abs().max().item() <= tol
    nxt = torch.from_numpy(z[f"{name}_first_tokens"])
    last = torch.tensor([c - 1 for c in cu[1:]])
    assert torch.equal(logits[last].float().argmax(-1), nxt)
    cur = list(lens)
    for step in range(2):
        logits, kvs = orc.forward(nxt, torch.tensor(cur), list(range(len(lens) + 1)), kvs, prefill=False)
        ref = torch.from_numpy(z[f"{name}_decode{step}_logits"])
        assert (logits.float() - ref.float()).abs().max().item() <= tol, f"decode step {step}"
        nxt = torch.from_numpy(z[f"{name}_decode{step}_tokens"])
        cur = [c + 1 for c in cur]


# ------------------------------------------------------------------------------------------------ chooser
def test_chooser_port_matches_reference_chooser():
    """tgis_b200.utils.tokens.HeterogeneousNextTokenChooser vs the reference class run on the same scores with the same
    seeds (CPU RNG): chosen ids, warped scores and logprobs bit-identical, before and after filter()."""
    import tgis_b200  # noqa: F401
    from tgis_b200 import pb
    from tgis_b200.utils.tokens import HeterogeneousNextTokenChooser

    z = _npz("chooser.npz")
    params = [pb.NextTokenChooserParameters.FromString(bytes.fromhex(h)) for h in bytes(z["params"]).decode().split("\n")]
    ch = HeterogeneousNextTokenChooser.from_pb(pb=params, model_eos_token_id=2, model_pad_token_id=0,
                                               return_logprobs=[False, True, False, True, False], dtype=torch.float32,
                                               device=torch.device("cpu"))
    all_scores = torch.from_numpy(z["all_scores"])
    input_ids = torch.from_numpy(z["input_ids"])
    for s in range(all_scores.shape[0]):
        ids, sc, lp = ch(input_ids, all_scores[s].clone())
        assert np.array_equal(ids.numpy(), z["ids"][s]), f"step {s}"
        assert np.array_equal(sc.numpy(), z["scores"][s], equal_nan=True)
        assert np.allclose(lp.numpy(), z["logprobs"][s], rtol=0, atol=0, equal_nan=True)
    ch2 = ch.filter([1, 3, 4])
    ids, _, _ = ch2(input_ids[[1, 3, 4]], all_scores[0][[1, 3, 4]].clone())
    assert np.array_equal(ids.numpy(), z["ids_filtered"])
