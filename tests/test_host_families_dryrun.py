"""The host wiring of the flash families that run op by op from Python (GPT-NeoX, Santacoder, Falcon), executed END TO END
on CPU: the product's modeling code (weight loading, fused-projection layouts and re-layouts, the strided views it hands to
the attention ops, residual forms, KV placement through the paged block table) with the kernels replaced by
tests/dryrun_ops.py (the oracle's arithmetic), compared with the family's oracle over a prefill and four decode steps.

Both sides then do the same arithmetic, so the logits must agree to fp16 rounding (<= 2 ulp of the logit scale): any
difference is a wiring bug.  The kernels themselves are covered by the `-m gpu` tests."""
import os
import types

import pytest
import torch

from oracle import falcon as ofa
from oracle import neox as onx
from oracle import santacoder as osc
from tests.dryrun_ops import patch_ops


def _run(model, oracle, vocab, n_heads, hidden, kv_heads, n_layers):
    from tgis_b200.utils.paged import PagedKVCacheManager, PagedKVState
    model.kv_cache_manager = PagedKVCacheManager(n_layers, n_heads, hidden, kv_heads=kv_heads, device="cpu", total_num_gpu_blocks=48)
    mgr = model.kv_cache_manager
    g = torch.Generator().manual_seed(7)
    lens = [5, 17, 1, 33, 16]
    prompts = [torch.randint(0, vocab, (L,), generator=g).tolist() for L in lens]
    n_new = 5
    ref_tokens, ref_logits = oracle.generate_greedy(prompts, n_new)
    B = len(prompts)
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    sids = mgr.allocate_tokens(lens, reserve_tokens=[n_new] * B)
    kv = PagedKVState(sequence_ids=sids, block_table=mgr.block_table_tensor(sids), context_lens=torch.tensor(lens, dtype=torch.int32),
                      slot_mapping=mgr.slot_mapping_for(sids, [0] * B, lens), max_blocks=0)
    input_ids = torch.tensor([t for p in prompts for t in p], dtype=torch.int64)
    position_ids = torch.cat([torch.arange(L) for L in lens])
    cu_t = torch.tensor(cu, dtype=torch.int32)
    last = (cu_t[1:] - 1).to(torch.int64)

    def check(got, ref, what):
        tol = 2 * 2.0 ** -10 * max(1.0, ref.float().abs().max().item())
        err = (got.float() - ref.float()).abs().max().item()
        assert err <= tol, f"{what}: max logit difference {err:.3e} > {tol:.3e}"

    with torch.inference_mode():
        logits, _ = model.forward(input_ids, position_ids, cu_t, None, max(lens), None, kv, None, last)
        check(logits, ref_logits[0], "prefill")
        cur = list(lens)
        for step in range(1, n_new):
            nxt = ref_tokens[:, step - 1]
            pos = torch.tensor(cur, dtype=torch.int64)
            kv.slot_mapping = mgr.slot_mapping_for(sids, cur, [1] * B)
            cur = [c + 1 for c in cur]
            kv.context_lens = torch.tensor(cur, dtype=torch.int32)
            ar = torch.arange(B + 1, dtype=torch.int32)
            logits, _ = model.forward(nxt, pos, ar, ar, max(cur), None, kv, None, None)
            check(logits, ref_logits[step], f"decode step {step}")
    mgr.free_sequences(sids)
    assert mgr.free_blocks == mgr.total_num_gpu_blocks


def _weights(tmp_path, sd):
    from safetensors.torch import save_file
    from tgis_b200.utils.dist import FakeGroup
    from tgis_b200.utils.weights import Weights
    path = os.path.join(str(tmp_path), "model.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    return Weights([path], device="cpu", dtype=torch.float16, process_group=FakeGroup(0, 1))


NEOX = [("parallel_d64_rot25", onx.NeoXConfig(256, 1024, 2, 4, 512, rotary_pct=0.25, use_parallel_residual=True)),
        ("sequential_d128_rot100", onx.NeoXConfig(256, 1024, 2, 2, 512, rotary_pct=1.0, use_parallel_residual=False)),
        ("parallel_d128_rot50_tanh", onx.NeoXConfig(512, 2048, 3, 4, 384, rotary_pct=0.5, use_parallel_residual=True, hidden_act="gelu_fast"))]


@pytest.mark.parametrize("name,cfg", NEOX, ids=[c[0] for c in NEOX])
def test_neox_host_wiring(tmp_path, monkeypatch, name, cfg):
    from tgis_b200.models.custom_modeling import flash_neox_modeling as m
    patch_ops(monkeypatch, m)
    sd = onx.make_state_dict(cfg, seed=11, std=0.04)
    ns = types.SimpleNamespace(model_type="gpt_neox", hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                               num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                               vocab_size=cfg.vocab_size, rotary_pct=cfg.rotary_pct, rotary_emb_base=cfg.rotary_emb_base,
                               layer_norm_eps=cfg.layer_norm_eps, use_parallel_residual=cfg.use_parallel_residual,
                               hidden_act=cfg.hidden_act, quantize=None, max_position_embeddings=512)
    model = m.FlashGPTNeoXForCausalLM(ns, _weights(tmp_path, sd))
    _run(model, onx.NeoXOracle(cfg, sd), cfg.vocab_size, cfg.num_attention_heads, cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers)


SANTACODER = [("h4_d64", osc.SantacoderConfig(256, 1024, 2, 4, 512, n_positions=256)),
              ("h2_d128_gelu", osc.SantacoderConfig(256, 1024, 2, 2, 512, n_positions=256, activation_function="gelu")),
              ("h24_d64_two_launches", osc.SantacoderConfig(1536, 2048, 1, 24, 384, n_positions=256))]


@pytest.mark.parametrize("name,cfg", SANTACODER, ids=[c[0] for c in SANTACODER])
def test_santacoder_host_wiring(tmp_path, monkeypatch, name, cfg):
    from tgis_b200.models.custom_modeling import flash_santacoder_modeling as m
    fake = patch_ops(monkeypatch, m)
    sd = osc.make_state_dict(cfg, seed=11, std=0.04)
    ns = types.SimpleNamespace(model_type="gpt_bigcode", hidden_size=cfg.hidden_size, n_inner=cfg.n_inner,
                               num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                               vocab_size=cfg.vocab_size, n_positions=cfg.n_positions, layer_norm_epsilon=cfg.layer_norm_epsilon,
                               activation_function=cfg.activation_function, multi_query=True, transpose=False, quantize=None)
    model = m.FlashSantacoderForCausalLM(ns, _weights(tmp_path, sd))
    assert m.FlashSantacoderForCausalLM.kv_cache_layout(ns, 1) == (1, 1)
    _run(model, osc.SantacoderOracle(cfg, sd), cfg.vocab_size, cfg.num_attention_heads, cfg.hidden_size, 1, cfg.num_hidden_layers)
    launches = [c for c in fake.calls if c[0] == "attn_decode_paged"]
    per_layer_step = len(launches) // (cfg.num_hidden_layers * 4)
    assert per_layer_step == -(-cfg.num_attention_heads // 16)  # 24 heads: 16 + 8


FALCON = [("mqa_parallel_h4_d64", ofa.FalconConfig(256, 2, 4, 1, 512, new_decoder_architecture=False, parallel_attn=True)),
          ("mqa_parallel_h20_d64", ofa.FalconConfig(1280, 1, 20, 1, 384, new_decoder_architecture=False, parallel_attn=True)),
          ("mqa_sequential_bias_d128", ofa.FalconConfig(256, 2, 2, 1, 512, new_decoder_architecture=False, parallel_attn=False, bias=True)),
          ("gqa_large_h8_kv2_d64", ofa.FalconConfig(512, 2, 8, 2, 384, new_decoder_architecture=True, parallel_attn=True))]


@pytest.mark.parametrize("name,cfg", FALCON, ids=[c[0] for c in FALCON])
def test_falcon_host_wiring(tmp_path, monkeypatch, name, cfg):
    from tgis_b200.models.custom_modeling import flash_rw_modeling as m
    patch_ops(monkeypatch, m)
    sd = ofa.make_state_dict(cfg, seed=11, std=0.04)
    ns = types.SimpleNamespace(model_type="RefinedWeb" if cfg.new_decoder_architecture else "RefinedWebModel", hidden_size=cfg.hidden_size,
                               n_layer=cfg.num_hidden_layers, n_head=cfg.n_head, n_head_kv=cfg.n_head_kv, vocab_size=cfg.vocab_size,
                               new_decoder_architecture=cfg.new_decoder_architecture, parallel_attn=cfg.parallel_attn, bias=cfg.bias,
                               layer_norm_epsilon=cfg.layer_norm_epsilon, multi_query=True, alibi=False, quantize=None,
                               max_position_embeddings=512)
    model = m.FlashRWForCausalLM(ns, _weights(tmp_path, sd))
    _run(model, ofa.FalconOracle(cfg, sd), cfg.vocab_size, cfg.n_head, cfg.hidden_size, cfg.n_head_kv, cfg.num_hidden_layers)


def test_python_fused_step_protocol(tmp_path, monkeypatch):
    """make_step / run_step (python_step.PythonFusedGreedy): the decode forward written into the caller's logits buffer plus
    the in-step arg-max with a banned id, against the oracle's greedy continuation."""
    from tgis_b200.models.custom_modeling import flash_santacoder_modeling as m
    from tgis_b200.models.custom_modeling import python_step
    from tgis_b200.utils.paged import PagedKVCacheManager, PagedKVState
    patch_ops(monkeypatch, m, python_step)
    cfg = osc.SantacoderConfig(256, 1024, 2, 4, 512, n_positions=256)
    sd = osc.make_state_dict(cfg, seed=3, std=0.04)
    ns = types.SimpleNamespace(model_type="gpt_bigcode", hidden_size=256, n_inner=1024, num_hidden_layers=2, num_attention_heads=4,
                               vocab_size=512, n_positions=256, layer_norm_epsilon=1e-5, activation_function=cfg.activation_function,
                               multi_query=True, transpose=False, quantize=None)
    model = m.FlashSantacoderForCausalLM(ns, _weights(tmp_path, sd))
    mgr = model.kv_cache_manager = PagedKVCacheManager(2, 4, 256, kv_heads=1, device="cpu", total_num_gpu_blocks=16)
    lens = [6, 3]
    g = torch.Generator().manual_seed(1)
    prompts = [torch.randint(0, 512, (L,), generator=g).tolist() for L in lens]
    ref_tokens, ref_logits = osc.SantacoderOracle(cfg, sd).generate_greedy(prompts, 3)
    sids = mgr.allocate_tokens(lens, reserve_tokens=[4, 4])
    kv = PagedKVState(sequence_ids=sids, block_table=mgr.block_table_tensor(sids), context_lens=torch.tensor(lens, dtype=torch.int32),
                      slot_mapping=mgr.slot_mapping_for(sids, [0, 0], lens), max_blocks=0)
    with torch.inference_mode():
        model.forward(torch.tensor(prompts[0] + prompts[1]), torch.cat([torch.arange(L) for L in lens]),
                      torch.tensor([0, 6, 9], dtype=torch.int32), None, 6, None, kv, None, None)
        # what FlashCausalLM._decode_fused_greedy does around the step
        input_ids, position_ids = ref_tokens[:, 0].clone(), torch.tensor(lens)
        kv.slot_mapping = mgr.slot_mapping_for(sids, lens, [1, 1])
        kv.context_lens = torch.tensor([L + 1 for L in lens], dtype=torch.int32)
        logits, next_ids = torch.empty(2, 512, dtype=torch.float16), torch.empty(2, dtype=torch.int64)
        step = model.make_step(T=2, B=2, is_prefill=False, max_s=12, input_ids=input_ids, position_ids=position_ids, kv=kv,
                               logits=logits, next_ids=next_ids)
        step.banned = torch.tensor([-1, int(ref_tokens[1, 1])])  # row 1 may not emit the oracle's choice
        model.run_step(step)
    assert (logits.float() - ref_logits[1].float()).abs().max().item() <= 2 * 2.0 ** -10 * max(1.0, ref_logits[1].float().abs().max().item())
    assert int(next_ids[0]) == int(ref_tokens[0, 1])
    second = ref_logits[1][1].float().clone()
    second[int(ref_tokens[1, 1])] = float("-inf")
    assert int(next_ids[1]) == int(second.argmax())
    with pytest.raises(NotImplementedError):
        model.make_step(T=2, B=2, is_prefill=True, max_s=12, input_ids=input_ids, position_ids=position_ids, kv=kv, logits=logits)


# ------------------------------------------------------------------------------------------ tensor parallel, world size 2 (gloo)
def _tp_worker(rank, world, port, family, path, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import tgis_b200  # noqa: F401
    from tests.dryrun_ops import DryRunOps
    from tgis_b200.utils import flash_attn, layers
    from tgis_b200.utils.dist import initialize_torch_distributed
    from tgis_b200.utils.paged import PagedKVCacheManager, PagedKVState
    from tgis_b200.utils.weights import Weights
    fake = DryRunOps()
    if family == "santacoder":
        from tgis_b200.models.custom_modeling import flash_santacoder_modeling as m
        ns = types.SimpleNamespace(model_type="gpt_bigcode", hidden_size=256, n_inner=1024, num_hidden_layers=2, num_attention_heads=4,
                                   vocab_size=512, n_positions=256, layer_norm_epsilon=1e-5, activation_function="gelu_pytorch_tanh",
                                   multi_query=True, transpose=False, quantize=None)
        build, heads, kv_total = m.FlashSantacoderForCausalLM, 4, 1
    elif family == "falcon_large":
        from tgis_b200.models.custom_modeling import flash_rw_modeling as m
        ns = types.SimpleNamespace(model_type="RefinedWeb", hidden_size=512, n_layer=2, n_head=8, n_head_kv=2, vocab_size=384,
                                   new_decoder_architecture=True, parallel_attn=True, bias=False, layer_norm_epsilon=1e-5,
                                   multi_query=True, alibi=False, quantize=None, max_position_embeddings=512)
        build, heads, kv_total = m.FlashRWForCausalLM, 8, 2
    else:
        from tgis_b200.models.custom_modeling import flash_neox_modeling as m
        ns = types.SimpleNamespace(model_type="gpt_neox", hidden_size=256, intermediate_size=1024, num_hidden_layers=2,
                                   num_attention_heads=4, vocab_size=512, rotary_pct=0.25, rotary_emb_base=10000.0, layer_norm_eps=1e-5,
                                   use_parallel_residual=(family == "neox_parallel"), hidden_act="gelu", quantize=None,
                                   max_position_embeddings=512)
        build, heads, kv_total = m.FlashGPTNeoXForCausalLM, 4, 4
    from tgis_b200.models.custom_modeling import python_step
    for mod in (layers, flash_attn, python_step, m):
        if hasattr(mod, "_ops"):
            mod._ops = lambda: fake
    pg = initialize_torch_distributed(world, rank)  # gloo on CPU
    model = build(ns, Weights([path], device="cpu", dtype=torch.float16, process_group=pg))
    kv_heads, kv_world = getattr(build, "kv_cache_layout", lambda c, w: (kv_total, w))(ns, world)
    mgr = model.kv_cache_manager = PagedKVCacheManager(2, heads, ns.hidden_size, kv_heads=kv_heads, tensor_parallel_size=kv_world,
                                                       device="cpu", total_num_gpu_blocks=16)
    lens = [7, 18, 2]
    g = torch.Generator().manual_seed(4)
    prompts = [torch.randint(0, ns.vocab_size, (L,), generator=g).tolist() for L in lens]
    sids = mgr.allocate_tokens(lens, reserve_tokens=[3] * 3)
    kv = PagedKVState(sequence_ids=sids, block_table=mgr.block_table_tensor(sids), context_lens=torch.tensor(lens, dtype=torch.int32),
                      slot_mapping=mgr.slot_mapping_for(sids, [0] * 3, lens), max_blocks=0)
    outs = []
    with torch.inference_mode():
        cu = torch.tensor([0, 7, 25, 27], dtype=torch.int32)
        logits, _ = model.forward(torch.tensor([t for p in prompts for t in p]), torch.cat([torch.arange(L) for L in lens]), cu, None,
                                  max(lens), None, kv, None, (cu[1:] - 1).long())
        outs.append(logits.clone())
        nxt = logits.float().argmax(-1)
        kv.slot_mapping = mgr.slot_mapping_for(sids, lens, [1] * 3)
        kv.context_lens = torch.tensor([L + 1 for L in lens], dtype=torch.int32)
        ar = torch.arange(4, dtype=torch.int32)
        logits, _ = model.forward(nxt, torch.tensor(lens), ar, ar, max(lens) + 1, None, kv, None, None)
        outs.append(logits.clone())
    out_q.put((rank, [o.numpy() for o in outs], prompts))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("family", ["santacoder", "falcon_large", "neox_parallel", "neox_sequential"])
def test_world_size_2_family_wiring_over_gloo(tmp_path, family):
    """Two CPU ranks over gloo: per-rank head / group / column shards, the replicated multi-query KV head, the all-reduces and
    the vocab-sharded head gather of the product's host code reproduce the single-rank oracle."""
    import socket
    import torch.multiprocessing as mp
    from safetensors.torch import save_file
    if family == "santacoder":
        cfg = osc.SantacoderConfig(256, 1024, 2, 4, 512, n_positions=256)
        sd, oracle = osc.make_state_dict(cfg, seed=9, std=0.04), None
        oracle = osc.SantacoderOracle(cfg, sd)
    elif family == "falcon_large":
        cfg = ofa.FalconConfig(512, 2, 8, 2, 384, new_decoder_architecture=True, parallel_attn=True)
        sd = ofa.make_state_dict(cfg, seed=9, std=0.04)
        oracle = ofa.FalconOracle(cfg, sd)
    else:
        cfg = onx.NeoXConfig(256, 1024, 2, 4, 512, rotary_pct=0.25, use_parallel_residual=(family == "neox_parallel"))
        sd = onx.make_state_dict(cfg, seed=9, std=0.04)
        oracle = onx.NeoXOracle(cfg, sd)
    path = os.path.join(str(tmp_path), "model.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tp_worker, args=(r, 2, port, family, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, outs, prompts = q.get(timeout=90)
        res[rank] = [torch.from_numpy(o) for o in outs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_tokens, ref_logits = oracle.generate_greedy(prompts, 2)
    for step in range(2):
        ref = ref_logits[step].float()
        # fp16 partial sums of two ranks vs one fp32 accumulation: a few ulp of the activation scale, not 2 ulp of the logits
        tol = 4e-3 * ref.abs().max().item() + 2e-3
        for rank in range(2):
            assert res[rank][step].shape == ref.shape, f"rank {rank}: the vocab shards were not gathered"
            assert (res[rank][step].float() - ref).abs().max().item() <= tol, f"{family} rank {rank} step {step}"
        assert torch.equal(res[0][step], res[1][step]), "ranks disagree"
