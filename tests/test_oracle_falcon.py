"""The Falcon / RefinedWeb oracle (multi-query and grouped-query attention, three layer forms), pinned twice: against the
reference's OWN FlashRWForCausalLM executed on CPU (tests/golden/flash_rw_ref.npz, written by tests/golden/make_golden.py with
the CUDA extensions shimmed by the oracle's restatements) and against an independent implementation, transformers'
FalconForCausalLM (eager, fp32, CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import falcon as ofa

CASES = {  # the cases of tests/golden/make_golden.py: FALCON_CASES
    "mqa_parallel": ofa.FalconConfig(128, 2, 4, 1, 160, new_decoder_architecture=False, parallel_attn=True),
    "gqa_large": ofa.FalconConfig(256, 2, 8, 2, 160, new_decoder_architecture=True, parallel_attn=True),
    "mqa_sequential_bias": ofa.FalconConfig(128, 2, 2, 1, 160, new_decoder_architecture=False, parallel_attn=False, bias=True),
}


@pytest.mark.parametrize("name", list(CASES))
def test_falcon_oracle_matches_transformers(name):
    from transformers import FalconConfig, FalconForCausalLM
    cfg = CASES[name]
    sd = ofa.make_state_dict(cfg, seed=3, std=0.05)
    hf_cfg = FalconConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_hidden_layers,
                          num_attention_heads=cfg.n_head, num_kv_heads=cfg.n_head_kv, new_decoder_architecture=cfg.new_decoder_architecture,
                          multi_query=True, parallel_attn=cfg.parallel_attn, bias=cfg.bias, alibi=False,
                          layer_norm_epsilon=cfg.layer_norm_epsilon, hidden_dropout=0.0, attention_dropout=0.0,
                          max_position_embeddings=64, tie_word_embeddings=False)
    hf_cfg._attn_implementation = "eager"
    hf = FalconForCausalLM(hf_cfg).float().eval()
    missing, unexpected = hf.load_state_dict({k: v.float() for k, v in sd.items()}, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    prompt = torch.randint(0, cfg.vocab_size, (11,), generator=torch.Generator().manual_seed(0))
    toks, logits = ofa.FalconOracle(cfg, sd).generate_greedy([prompt.tolist()], 3)
    seq = prompt.tolist()
    for s in range(3):
        with torch.no_grad():
            ref = hf(torch.tensor(seq)[None]).logits[0, -1]
        assert (logits[s][0].float() - ref).abs().max().item() <= 4e-3 * ref.abs().max().item() + 2e-3, f"step {s}"
        seq.append(int(toks[0, s]))


@pytest.mark.parametrize("name", list(CASES))
def test_falcon_oracle_matches_reference_flash_rw_graph(name):
    """graph wiring - the two fused-QKV layouts, rotary on q and k, KV placement, one / two / sequential LayerNorm forms, the
    fp16 add of the parallel branches, final norm, head - must agree with oracle/falcon.py up to CPU fp16 GEMM rounding:
    <= 2 fp16 ulp of the logit scale, prefill and two decode steps."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "flash_rw_ref.npz"))
    cfg = CASES[name]
    oracle = ofa.FalconOracle(cfg, ofa.make_state_dict(cfg, seed=29, std=0.06))
    lens = [int(x) for x in z[f"{name}_lens"]]
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    pos = torch.cat([torch.arange(L) for L in lens])
    logits = oracle.forward(torch.from_numpy(z[f"{name}_input_ids"]), pos, cu, decode=False)
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
