"""Long-trajectory greedy parity report: B sequences generated through FlashCausalLM.generate_token for n_new
tokens each (prefill, then the CUDA-graph-replayed fused step at contexts L0 .. L0 + n_new), checked token by token against the
CPU oracle run teacher-forced on the generated ids.  A token counts when the oracle's top-1 / top-2 logit gap exceeds 2 fp16 ulp
at that magnitude (the tie band); it must then equal the oracle's arg-max.  Writes a JSON report.

  python tools/parity_report.py [--arch llama-2-7b] [--quantize gptq] [--batch 64] [--prompt 1024] [--new 1024] [--layers 2]
                                [--out profiles/r2_parity_report.json]
Default = 64 x 1024 generated tokens on a 2-layer model of Llama-2-7B widths (vocab 4096: the head's cost on the CPU side).
"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="llama-2-7b")
    ap.add_argument("--quantize", default="gptq")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--prompt", type=int, default=1024)
    ap.add_argument("--new", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--vocab", type=int, default=4096)
    ap.add_argument("--out", default="profiles/r2_parity_report.json")
    a = ap.parse_args()
    quantize = None if a.quantize in ("none", "fp16") else a.quantize
    from safetensors.torch import save_file

    from oracle import llama as oll  # checker
    import tgis_b200  # noqa: F401
    from tgis_b200 import pb
    from tgis_b200.inference_engine import InferenceEngine
    from tgis_b200.models.flash_causal_lm import FlashCausalLM
    from tgis_b200.utils.dist import FakeGroup
    from tgis_b200.utils.synthetic import ARCHS, llama_config, make_tokenizer
    from tgis_b200.utils.weights import Weights

    H, I, _, h, kv, _ = ARCHS[a.arch]
    max_len = a.prompt + a.new
    cfg = llama_config(a.arch, quantize=quantize, max_position_embeddings=max_len + 64, num_layers=a.layers)
    cfg.vocab_size = a.vocab
    ocfg = oll.LlamaConfig(H, I, a.layers, h, kv, a.vocab, cfg.rms_norm_eps, cfg.rope_theta)
    t0 = time.time()
    sd = oll.make_state_dict(ocfg, seed=21, quantize=quantize, std=0.02)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "model.safetensors")
        save_file({k: v.contiguous() for k, v in sd.items()}, path)
        weights = Weights([path], device="cuda:0", dtype=torch.float16, process_group=FakeGroup(0, 1))
        tok = make_tokenizer(a.vocab)
        engine = InferenceEngine(tmp, None, torch.float16, quantize, cfg, max_len, weights=weights, tokenizer=tok)
        blocks = a.batch * ((max_len + 16) // 16 + 1) + 8
        model = FlashCausalLM(tmp, None, "tgis_native", torch.float16, quantize, cfg, engine=engine, num_kv_blocks=blocks)
    g = torch.Generator().manual_seed(5)
    prompts = [torch.randint(4, a.vocab, (a.prompt,), generator=g).tolist() for _ in range(a.batch)]
    reqs = [pb.Request(id=i, inputs=" ".join(f"<tok{t}>" for t in p), input_length=len(p), max_output_length=a.new,
                       parameters=pb.NextTokenChooserParameters(temperature=0.0, top_p=1.0, min_new_tokens=a.new))
            for i, p in enumerate(prompts)]
    got = torch.zeros(a.batch, a.new, dtype=torch.long)
    with torch.inference_mode():
        batch, errs = model.batch_type.from_pb(pb.Batch(id=0, requests=reqs), tok, torch.float16, model.device, None, None, True)
        assert not errs
        out = model.generate_token(batch, first=True)
        for step in range(a.new):
            for t in out[0]:
                got[t.request_id, step] = t.token_id
            if step + 1 < a.new:
                out = model.generate_token(batch)
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    graph_steps = max(0, a.new - 1 - 2)
    del model, batch
    torch.cuda.empty_cache()
    # oracle, teacher-forced on the generated ids
    oracle = oll.LlamaOracle(oll.build_shards(ocfg, sd, 1))
    eos = cfg.eos_token_id
    stats = dict(checked=0, ties=0, mismatches=0, first_mismatches=[])

    def on_step(step, logits):
        lg = logits.float().clone()
        lg[:, eos] = float("-inf")
        top2 = lg.topk(2, -1)
        gap = top2.values[:, 0] - top2.values[:, 1]
        decisive = gap > 2 * top2.values[:, 0].abs().clamp(min=1.0) * 2.0 ** -10
        same = top2.indices[:, 0] == got[:, step]
        stats["checked"] += int(decisive.sum())
        stats["ties"] += int((~decisive).sum())
        bad = decisive & ~same
        stats["mismatches"] += int(bad.sum())
        for b in torch.nonzero(bad).flatten().tolist()[:4]:
            if len(stats["first_mismatches"]) < 16:
                stats["first_mismatches"].append({"sequence": b, "step": step, "got": int(got[b, step]), "oracle": int(top2.indices[b, 0]),
                                                  "gap": float(gap[b])})
        if step % 64 == 0:
            print(f"oracle step {step}: checked {stats['checked']} ties {stats['ties']} mismatches {stats['mismatches']}", flush=True)

    t1 = time.time()
    oracle.generate_greedy(prompts, a.new, banned_token=eos, forced=got, keep_logits=False, on_step=on_step)
    report = {
        "what": "greedy ids of FlashCausalLM.generate_token vs the CPU oracle run teacher-forced on them (SURVEY.md §8c exactness rule)",
        "model": f"{a.arch} widths, {a.layers} layers, vocab {a.vocab}, {quantize or 'fp16'}", "batch": a.batch, "prompt_len": a.prompt,
        "new_tokens_per_sequence": a.new, "generated_tokens": a.batch * a.new, "contexts": [a.prompt, a.prompt + a.new - 1],
        "decode_steps_through_cuda_graph_replay": graph_steps,
        "tokens_checked_outside_tie_band": stats["checked"], "tokens_inside_tie_band": stats["ties"],
        "mismatches_outside_tie_band": stats["mismatches"], "first_mismatches": stats["first_mismatches"],
        "tie_band": "oracle top-1 minus top-2 logit <= 2 * max(|top-1|, 1) * 2^-10",
        "seconds": {"gpu_generation_incl_model_build": round(t_gpu, 1), "cpu_oracle": round(time.time() - t1, 1)},
    }
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report))


if __name__ == "__main__":
    main()
