"""GPU parity through the public API: FlashCausalLMBatch.from_pb -> FlashCausalLM.generate_token (prefill, decode,
concatenate, prune) against the CPU oracle's greedy generation, plus the reference's batching-integrity properties
(scripts/batch_integrity_checks/batching_integrity_checks.py:96-153): batched == single, pruned == un-pruned,
concatenated == separate.

Exactness: token ids must equal the oracle's wherever the oracle's top-2 logit gap exceeds 2 fp16 ulp;
the synthetic model is seeded so that no step of these cases falls inside the tie band (asserted).
"""
import os
import types

import pytest
import torch

from oracle import llama as oll

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
EOS = 2


def _setup(tmp_path, quantize):
    from safetensors.torch import save_file
    import tgis_b200  # noqa: F401
    from tgis_b200.inference_engine import InferenceEngine
    from tgis_b200.models.flash_causal_lm import FlashCausalLM
    from tgis_b200.utils.dist import FakeGroup
    from tgis_b200.utils.synthetic import llama_config, make_tokenizer
    from tgis_b200.utils.weights import Weights

    cfg = llama_config("tiny-test", quantize=quantize, max_position_embeddings=256)
    ocfg = oll.LlamaConfig(cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.num_attention_heads,
                           cfg.num_key_value_heads, cfg.vocab_size, cfg.rms_norm_eps, cfg.rope_theta)
    sd = oll.make_state_dict(ocfg, seed=99, quantize=quantize, std=0.08)
    path = os.path.join(str(tmp_path), "model.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    weights = Weights([path], device=DEV, dtype=torch.float16, process_group=FakeGroup(0, 1))
    tok = make_tokenizer(cfg.vocab_size)
    engine = InferenceEngine(str(tmp_path), None, torch.float16, quantize, cfg, 256, weights=weights, tokenizer=tok)
    model = FlashCausalLM(str(tmp_path), None, "tgis_native", torch.float16, quantize, cfg, engine=engine, num_kv_blocks=128)
    oracle = oll.LlamaOracle(oll.build_shards(ocfg, sd, 1))
    return model, oracle, tok


def _text(ids):
    return " ".join("test" if i == 3 else f"<tok{i}>" for i in ids)


def _pb_batch(batch_id, prompts, n_new, first_id=0, logprobs=False):
    from tgis_b200 import pb
    reqs = []
    for i, p in enumerate(prompts):
        reqs.append(pb.Request(id=first_id + i, inputs=_text(p), input_length=len(p), truncate=False, max_output_length=n_new,
                               parameters=pb.NextTokenChooserParameters(temperature=0.0, top_p=1.0, min_new_tokens=n_new),
                               details=pb.RequestedDetails(logprobs=logprobs)))
    return pb.Batch(id=batch_id, requests=reqs)


def _prompts(seed, lens, vocab):
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(4, vocab, (L,), generator=g).tolist() for L in lens]


def _oracle_tokens(oracle, prompts, n_new):
    """-> (tokens [B, n_new], n_exact [B]): per sequence, the number of leading steps whose oracle top-2 gap exceeds
    2 fp16 ulp — up to there the ids must match exactly; past a near-tie the trajectories may legitimately fork."""
    toks, logits = oracle.generate_greedy(prompts, n_new, banned_token=EOS)
    n_exact = [n_new] * len(prompts)
    for step, lg in enumerate(logits):
        l = lg.float().clone()
        l[:, EOS] = float("-inf")
        top2 = l.topk(2, -1).values
        decisive = (top2[:, 0] - top2[:, 1]) > 2 * top2[:, 0].abs().clamp(min=1.0) * 2.0 ** -10
        for b in range(len(prompts)):
            if not decisive[b]:
                n_exact[b] = min(n_exact[b], step)
    return toks, n_exact


def _assert_prefix_equal(got, ref, n_exact, min_total):
    total = 0
    for b, n in enumerate(n_exact):
        assert list(got[b][:n]) == list(ref[b][:n]), f"sequence {b}: {got[b]} vs oracle {list(ref[b])} (exact for {n} steps)"
        total += n
    assert total >= min_total, f"test case too weak: only {total} decisive tokens"


@pytest.mark.parametrize("quantize", [None, "gptq"])
def test_generate_token_matches_oracle(tmp_path, quantize):
    model, oracle, tok = _setup(tmp_path, quantize)
    n_new = 12
    prompts = _prompts(1, [7, 33, 16, 1], 512)
    ref, n_exact = _oracle_tokens(oracle, prompts, n_new)
    with torch.inference_mode():
        batch, errs = model.batch_type.from_pb(_pb_batch(0, prompts, n_new), tok, torch.float16, model.device, None, None, True)
        assert not errs and len(batch) == 4
        got = [[] for _ in prompts]
        out = model.generate_token(batch, first=True)
        for t in out[0]:
            got[t.request_id].append(t.token_id)
        for step in range(n_new - 1):  # steps 0,1 eager fused; from step 2 on the CUDA graph replays
            out = model.generate_token(batch)
            assert out[1] is None and not out[2]
            for t in out[0]:
                got[t.request_id].append(t.token_id)
    _assert_prefix_equal(got, ref.tolist(), n_exact, min_total=24)
    # all_input_ids_tensor holds prompt + generated tokens (flash_causal_lm.py:533-535)
    for i, p in enumerate(prompts):
        assert batch.all_input_ids_tensor[i, :len(p) + n_new].tolist() == p + got[i]
    model.kv_cache_manager.free_sequences(batch.sequence_ids)
    assert model.kv_cache_manager.free_blocks == model.kv_cache_manager.total_num_gpu_blocks


def test_generate_token_general_chooser_path_matches_fused(tmp_path):
    """requests that ask for logprobs take the un-fused path (chooser + get_token_info): same tokens, finite logprobs."""
    model, oracle, tok = _setup(tmp_path, None)
    n_new = 6
    prompts = _prompts(2, [9, 20], 512)
    ref, n_exact = _oracle_tokens(oracle, prompts, n_new)
    with torch.inference_mode():
        batch, _ = model.batch_type.from_pb(_pb_batch(0, prompts, n_new, logprobs=True), tok, torch.float16, model.device, None, None, True)
        got = [[] for _ in prompts]
        toks = model.generate_token(batch, first=True)[0]
        for _ in range(n_new - 1):
            for t in toks:
                got[t.request_id].append(t.token_id)
                assert t.logprob <= 0.0 and t.logprob > -50.0
            toks = model.generate_token(batch)[0]
        for t in toks:
            got[t.request_id].append(t.token_id)
    _assert_prefix_equal(got, ref.tolist(), n_exact, min_total=6)


def test_concatenate_and_prune_are_kv_free_and_exact(tmp_path):
    """continuous batching: A (2 requests) runs 3 steps, B (2 requests) is prefilled and concatenated, request 1 is pruned
    two steps later; every request's tokens equal its single-request oracle generation."""
    model, oracle, tok = _setup(tmp_path, None)
    mgr = model.kv_cache_manager
    n_new = 10
    pa, pbs = _prompts(3, [12, 5], 512), _prompts(4, [20, 3], 512)
    ref, n_ex = {}, {}
    for i, p in enumerate(pa + pbs):
        t, ne = _oracle_tokens(oracle, [p], n_new)
        ref[i], n_ex[i] = t[0].tolist(), ne[0]
    got = {i: [] for i in range(4)}

    def take(out):
        for t in out[0]:
            got[t.request_id].append(t.token_id)

    with torch.inference_mode():
        A, _ = model.batch_type.from_pb(_pb_batch(0, pa, n_new, first_id=0), tok, torch.float16, model.device, None, None, True)
        take(model.generate_token(A, first=True))
        for _ in range(3):
            take(model.generate_token(A))
        Bb, _ = model.batch_type.from_pb(_pb_batch(1, pbs, n_new, first_id=2), tok, torch.float16, model.device, None, None, True)
        take(model.generate_token(Bb, first=True, for_concat=True))
        pool_before = mgr.pool.clone()
        C = model.batch_type.concatenate([A, Bb])
        assert torch.equal(mgr.pool, pool_before), "concatenate must not move KV"
        assert len(C) == 4 and C.batch_id == 0
        take(model.generate_token(C))
        take(model.generate_token(C))
        free_before = mgr.free_blocks
        C = model.batch_type.prune(C, [1])  # request 1 completed (router sends ascending completed ids)
        assert len(C) == 3 and mgr.free_blocks > free_before
        assert model.batch_type.prune(C, []) is C  # flash_causal_lm.py:296-298
        while min(len(got[i]) for i in (0, 2, 3)) < n_new:
            done = [i for i in (0, 2, 3) if len(got[i]) >= n_new and any(r.id == i for r in C.requests)]
            if done:
                C = model.batch_type.prune(C, done)
            take(model.generate_token(C))
        assert model.batch_type.prune(C, [r.id for r in C.requests]) is None
    for i in (0, 2, 3):
        assert got[i][:n_ex[i]] == ref[i][:n_ex[i]], f"request {i}: {got[i]} vs {ref[i]}"
    n1 = min(len(got[1]), n_ex[1])
    assert got[1][:n1] == ref[1][:n1]
    assert sum(n_ex.values()) >= 20, "test case too weak"
    assert mgr.free_blocks == mgr.total_num_gpu_blocks, "every block returned after all requests finished"


def test_prompt_prefix_equals_the_same_tokens_typed_in(tmp_path):
    """flash_causal_lm.py:97-107, :157-168: a request whose prefix embeddings are the embedding rows of some tokens must
    generate exactly what the request with those tokens prepended to its input generates (same KV, same positions), and the
    plain request batched with it must be unaffected by going in as embeddings."""
    from tgis_b200 import pb
    model, oracle, tok = _setup(tmp_path, None)
    n_new = 8
    prefix_ids, tail, other = _prompts(5, [6], 512)[0], _prompts(6, [9], 512)[0], _prompts(7, [14], 512)[0]
    ref, n_exact = _oracle_tokens(oracle, [prefix_ids + tail, other], n_new)
    table = model.model.get_input_embeddings().weight

    class Store:
        def get(self, prefix_id):
            assert prefix_id == "tuned"
            return table[torch.tensor(prefix_ids, device=table.device)].clone()

    reqs = _pb_batch(0, [tail, other], n_new).requests
    reqs[0].prefix_id = "tuned"
    with torch.inference_mode():
        batch, errs = model.batch_type.from_pb(pb.Batch(id=0, requests=list(reqs)), tok, torch.float16, model.device,
                                               model.word_embeddings, Store(), True)
        assert not errs and batch.input_ids is None and batch.input_lengths == [len(prefix_ids) + len(tail), len(other)]
        got = [[] for _ in range(2)]
        out = model.generate_token(batch, first=True)
        for _ in range(n_new):
            for t in out[0]:
                got[t.request_id].append(t.token_id)
            if len(got[0]) == n_new:
                break
            out = model.generate_token(batch)
    _assert_prefix_equal(got, ref.tolist(), n_exact, min_total=8)
    model.kv_cache_manager.free_sequences(batch.sequence_ids)


def test_paged_convention_and_speculative_verification(tmp_path):
    """The paged calling convention (paged_llama_modeling.py:443-462: forward(input_ids, position_ids, cache_data, ...)) on the same
    kernels, and the verification forward of speculative decoding (models/paged_causal_lm.py:481-562): k candidate sequences per
    parent, n tokens each, every token a row of the generation-form cache_data.  A candidate fed the true continuation must give,
    token by token, the logits that plain one-token-at-a-time decoding gives; the acceptance rule keeps it and frees the rest."""
    from tgis_b200.utils import paged

    model, oracle, tok = _setup(tmp_path, None)
    net, mgr = model.model, model.kv_cache_manager
    prompts = _prompts(31, [7, 19, 33], 512)
    n_spec = 4  # tokens verified per candidate: the last accepted one + 3 speculated
    with torch.inference_mode():
        # prefill through the paged convention
        pos, cd = paged.prepare_inputs_for_prefill([len(p) for p in prompts], mgr)
        ids = torch.tensor([t for p in prompts for t in p], dtype=torch.int64, device=DEV)
        logits, embeds = net(ids, pos, cd, None, True)
        last = (cd.context_lengths[1:] - 1).long()
        assert embeds.shape == (ids.shape[0], net.config.hidden_size)
        ref_tokens, ref_logits = oracle.generate_greedy(prompts, n_spec + 1, banned_token=None)
        _logits_close(logits[last], ref_logits[0], "paged-convention prefill")
        parents = cd.sequence_ids
        first = logits[last].float().argmax(-1)
        # plain decoding, one token per step, on CHILD copies so that the parents stay where they are
        plain_ids = [mgr.add_child_sequences(p, 1)[0] for p in parents]
        fed, plain_logits, chain = first, [], [first]
        for _ in range(n_spec):
            pos1, cd1 = paged.prepare_inputs_without_speculation(plain_ids, mgr)
            lg = net(fed, pos1, cd1)
            plain_logits.append(lg)
            fed = lg.float().argmax(-1)
            chain.append(fed)
        mgr.free_sequences(plain_ids)
        # speculative verification: candidate 0 = garbage, candidate 1 = the true continuation, in ONE forward
        pos_s, cd_s, children = paged.prepare_candidates(parents, n_candidates=2, n_tokens=n_spec, kv_cache_manager=mgr)
        truth = torch.stack(chain[:n_spec], 1)                                   # [B, n]: token fed at each of the n positions
        garbage = truth.clone()
        garbage[:, 1:] = (garbage[:, 1:] + 7) % 500 + 4
        cand = torch.stack([garbage, truth], 1)                                  # [B, k, n]
        lg_s = net(cand.reshape(-1), pos_s, cd_s).view(len(prompts), 2, n_spec, -1)
        for j in range(n_spec):
            _logits_close(lg_s[:, 1, j], plain_logits[j].cpu(), f"speculative row {j} vs plain decode step {j}")
        survivors, accepted = paged.accept_candidates(cand.cpu(), lg_s.float().argmax(-1).cpu(), children, mgr)
        for b, kids in enumerate(children):
            assert survivors[b] == kids[1] and accepted[b] == [int(c[b]) for c in chain[1:n_spec + 1]]
            assert mgr.sequence_length(kids[1]) == len(prompts[b]) + n_spec
        mgr.free_sequences(survivors, recursive=True)
    assert mgr.free_blocks == mgr.total_num_gpu_blocks


def _logits_close(got, ref, what):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    tol = 4e-3 * ref.abs().max().item() + 2e-3
    assert err.max().item() <= tol, f"{what}: max err {err.max().item():.4e} > {tol:.4e}"
