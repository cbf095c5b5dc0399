"""The Santacoder (gpt_bigcode, multi-query attention) oracle, pinned twice: against the reference's OWN
FlashSantacoderForCausalLM executed on CPU (tests/golden/flash_santacoder_ref.npz, written by tests/golden/make_golden.py
with the CUDA extensions shimmed by the oracle's restatements) and against an independent implementation, transformers'
GPTBigCodeForCausalLM (eager, fp32, CPU)."""
import os

import numpy as np
import torch

from oracle import santacoder as osc


def test_santacoder_oracle_matches_transformers():
    from transformers import GPTBigCodeConfig, GPTBigCodeForCausalLM
    cfg = osc.SantacoderConfig(128, 512, 2, 4, 160, n_positions=64)
    sd = osc.make_state_dict(cfg, seed=5, std=0.05)
    hf_cfg = GPTBigCodeConfig(vocab_size=cfg.vocab_size, n_positions=cfg.n_positions, n_embd=cfg.hidden_size, n_layer=cfg.num_hidden_layers,
                              n_head=cfg.num_attention_heads, n_inner=cfg.n_inner, activation_function=cfg.activation_function,
                              layer_norm_epsilon=cfg.layer_norm_epsilon, multi_query=True, attn_pdrop=0.0, resid_pdrop=0.0,
                              embd_pdrop=0.0)
    hf_cfg._attn_implementation = "eager"
    hf = GPTBigCodeForCausalLM(hf_cfg).float().eval()
    own = {k: v.float() for k, v in sd.items()}
    own["lm_head.weight"] = own["transformer.wte.weight"]  # tied head (flash_santacoder_modeling.py:447-449)
    missing, unexpected = hf.load_state_dict(own, strict=False)
    assert not unexpected, unexpected
    assert all("bias" in m and "attn" in m for m in missing), missing  # causal-mask buffers only
    prompt = torch.randint(0, cfg.vocab_size, (11,), generator=torch.Generator().manual_seed(0))
    oracle = osc.SantacoderOracle(cfg, sd)
    toks, logits = oracle.generate_greedy([prompt.tolist()], 3)
    seq = prompt.tolist()
    for s in range(3):
        with torch.no_grad():
            ref = hf(torch.tensor(seq)[None]).logits[0, -1]
        assert (logits[s][0].float() - ref).abs().max().item() <= 4e-3 * ref.abs().max().item() + 2e-3, f"step {s}"
        seq.append(int(toks[0, s]))


def test_santacoder_oracle_matches_reference_flash_santacoder_graph():
    """graph wiring - q | kv split of c_attn, one shared KV head, learned positions, KV placement, sequential residual, final
    norm, tied head - must agree with oracle/santacoder.py up to CPU fp16 GEMM rounding: <= 2 fp16 ulp of the logit scale."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "flash_santacoder_ref.npz"))
    cfg = osc.SantacoderConfig(128, 512, 2, 4, 160, n_positions=64)
    sd = osc.make_state_dict(cfg, seed=23, std=0.06)
    oracle = osc.SantacoderOracle(cfg, sd)
    lens = [int(x) for x in z["lens"]]
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    pos = torch.cat([torch.arange(L) for L in lens])
    logits = oracle.forward(torch.from_numpy(z["input_ids"]), pos, cu, decode=False)
    ref = torch.from_numpy(z["prefill_logits"])
    tol = 2 * 2.0 ** -10 * max(1.0, ref.float().abs().max().item())
    assert (logits.float() - ref.float()).abs().max().item() <= tol
    cur = list(lens)
    for step in range(2):
        nxt = torch.from_numpy(z[f"decode{step}_input"])
        logits = oracle.forward(nxt, torch.tensor(cur), list(range(len(lens) + 1)), decode=True)
        cur = [c + 1 for c in cur]
        ref = torch.from_numpy(z[f"decode{step}_logits"])
        assert (logits.float() - ref.float()).abs().max().item() <= tol, f"decode step {step}"
