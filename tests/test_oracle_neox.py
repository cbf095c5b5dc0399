"""The GPT-NeoX oracle against an independent implementation: transformers' GPTNeoXForCausalLM (eager, fp32, CPU)."""
import pytest
import torch

from oracle import neox as onx


@pytest.mark.parametrize("parallel", [True, False])
def test_neox_oracle_matches_transformers(parallel):
    from transformers import GPTNeoXConfig, GPTNeoXForCausalLM
    cfg = onx.NeoXConfig(hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, vocab_size=160,
                         rotary_pct=0.25, use_parallel_residual=parallel)
    sd = onx.make_state_dict(cfg, seed=3, std=0.05)
    hf_cfg = GPTNeoXConfig(hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, vocab_size=160,
                           rotary_pct=0.25, rotary_emb_base=10000, use_parallel_residual=parallel, layer_norm_eps=1e-5,
                           hidden_act="gelu", max_position_embeddings=64, attention_bias=True, tie_word_embeddings=False)
    hf_cfg._attn_implementation = "eager"
    hf = GPTNeoXForCausalLM(hf_cfg).float().eval()
    own = {k: v.float() for k, v in sd.items() if not k.endswith("inv_freq")}
    missing, unexpected = hf.load_state_dict(own, strict=False)
    assert not unexpected, unexpected
    assert all("rotary_emb" in m or "masked_bias" in m or m.endswith(".bias") and "attention.bias" in m for m in missing), missing
    g = torch.Generator().manual_seed(0)
    prompt = torch.randint(0, cfg.vocab_size, (11,), generator=g)
    oracle = onx.NeoXOracle(cfg, sd)
    toks, logits = oracle.generate_greedy([prompt.tolist()], 3)
    with torch.no_grad():
        ref = hf(prompt[None]).logits[0, -1]
    scale = ref.abs().max().item()
    assert (logits[0][0].float() - ref).abs().max().item() <= 4e-3 * scale + 2e-3
    # decode steps against HF re-run on the extended sequence
    seq = prompt.tolist()
    for s in range(1, 3):
        seq.append(int(toks[0, s - 1]))
        with torch.no_grad():
            ref = hf(torch.tensor(seq)[None]).logits[0, -1]
        assert (logits[s][0].float() - ref).abs().max().item() <= 4e-3 * ref.abs().max().item() + 2e-3


@pytest.mark.parametrize("name,parallel", [("parallel", True), ("sequential", False)])
def test_neox_oracle_matches_reference_flash_neox_graph(name, parallel):
    """tests/golden/flash_neox_ref.npz is the reference's OWN FlashGPTNeoXForCausalLM (flash_neox_modeling.py) executed on CPU by
    tests/golden/make_golden.py (its three CUDA extensions shimmed with the oracle's restatements): the graph wiring - QKV
    [h,3,d] -> [3,h,d] re-layout, partial rotary, KV placement, parallel / sequential residual, final norm, head - must agree
    with oracle/neox.py up to CPU fp16 GEMM rounding: <= 2 fp16 ulp of the logit scale, prefill and two decode steps."""
    import os
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "flash_neox_ref.npz"))
    cfg = onx.NeoXConfig(128, 512, 2, 2, 160, rotary_pct=0.25, use_parallel_residual=parallel)
    sd = onx.make_state_dict(cfg, seed=17, std=0.06)
    oracle = onx.NeoXOracle(cfg, sd)
    lens = [int(x) for x in z[f"{name}_lens"]]
    ids = torch.from_numpy(z[f"{name}_input_ids"])
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    pos = torch.cat([torch.arange(L) for L in lens])
    logits = oracle.forward(ids, pos, cu, decode=False)
    ref = torch.from_numpy(z[f"{name}_prefill_logits"])
    tol = 2 * 2.0 ** -10 * max(1.0, ref.float().abs().max().item())
    assert (logits.float() - ref.float()).abs().max().item() <= tol
    cur = list(lens)
    for step in range(2):
        nxt = torch.from_numpy(z[f"{name}_decode{step}_input"])
        logits = oracle.forward(nxt, torch.tensor(cur), list(range(len(lens) + 1)), decode=True)
        cur = [c + 1 for c in cur]
        ref = torch.from_numpy(z[f"{name}_decode{step}_logits"])
        assert (logits.float() - ref.float()).abs().max().item() <= tol, f"decode step {step}"
