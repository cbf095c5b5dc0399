"""GPU tests of the fused next-token chooser (csrc/chooser.cu) against the torch chooser of this package, which is pinned bit for
bit against the reference's classes (tests/golden/chooser.npz): HeterogeneousNextTokenChooser.__call__ (utils/tokens.py:238-271 of
the reference) with the Heterogeneous* warpers (utils/logits_process.py:93-317).

What must agree exactly: greedy ids, the set of tokens that survive repetition penalty + temperature + top-k, the warped values
themselves (bit for bit: same fp16 arithmetic).  Top-p: the kernel drops tied values as a group and sums probabilities in fixed
point where torch cuts inside a tie group and accumulates an fp16 cumsum, so the surviving sets may differ by tokens within
2^-10 of cumulative mass of the cut; log-probabilities within 2e-3.  Sampling: reproducible, lock-step across "ranks", and
distributed as softmax of the warped scores (chi-square over many draws)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _params(pb, **kw):
    d = dict(temperature=0.0, top_k=0, top_p=1.0, typical_p=0.0, min_new_tokens=0)
    d.update(kw)
    return pb.NextTokenChooserParameters(**d)


def _chooser(pbs, return_logprobs=None, eos=2, pad=0, current=None):
    from tgis_b200.utils.tokens import HeterogeneousNextTokenChooser
    return HeterogeneousNextTokenChooser.from_pb(pb=pbs, model_eos_token_id=eos, model_pad_token_id=pad,
                                                 return_logprobs=return_logprobs or [False] * len(pbs), dtype=torch.float16,
                                                 device=torch.device(DEV), current_tokens=current)


def _case(B, V, S, seed):
    g = torch.Generator().manual_seed(seed)
    scores = (torch.randn(B, V, generator=g) * 3).half().to(DEV)
    history = torch.randint(0, V, (B, S + 4), generator=g).to(DEV)
    pos = torch.randint(S // 2, S + 1, (B,), generator=g).to(DEV)
    pos[0] = S  # the longest row defines the history window every row sees
    return scores, history, pos


@pytest.mark.parametrize("V", [512, 32000, 128256])
def test_warped_scores_and_greedy_choice_match_the_torch_chooser(V):
    import tgis_b200  # noqa: F401
    from tgis_b200 import pb
    B, S = 6, 40
    scores, history, pos = _case(B, V, S, V)
    pbs = [_params(pb), _params(pb, repetition_penalty=1.2), _params(pb, temperature=0.7, top_k=5, seed=1),
           _params(pb, temperature=1.3, top_k=50, repetition_penalty=1.1, seed=2), _params(pb, min_new_tokens=3),
           _params(pb, top_k=1, temperature=2.0, seed=3, length_penalty=pb.NextTokenChooserParameters.LengthPenalty(start_index=1, decay_factor=1.5))]
    cur = [0, 0, 0, 0, 1, 4]
    ref_ch = _chooser(pbs, [True] * B, current=list(cur))
    dev_ch = _chooser(pbs, [True] * B, current=list(cur))
    ref_ids, ref_scores, ref_lp = ref_ch(input_ids=history[:, :S], scores=scores.clone())
    ids, lps, rks = dev_ch.choose_on_device(history, pos, scores, True, True)
    torch.cuda.synchronize()
    warped = dev_ch.device_chooser().scratch
    keep_ref = torch.isfinite(ref_scores)
    # no top-p here: the kernel's warped values are the torch chain's, bit for bit, wherever torch keeps a token; the kernel leaves
    # dropped tokens in the scratch row and remembers a cut-off instead, so compare through the chosen ids / log-probs / ranks too
    assert torch.equal(warped[keep_ref], ref_scores[keep_ref])
    greedy_rows = [0, 1, 4]
    assert ids[greedy_rows].tolist() == ref_ids[greedy_rows].tolist()
    assert ids[5].item() == ref_scores[5].argmax().item()  # top_k = 1: sampling has one survivor
    assert dev_ch.current_tokens == ref_ch.current_tokens
    for b in range(B):
        t = ids[b].item()
        assert keep_ref[b, t], f"row {b}: chose a token the torch chain dropped"
        # torch's log_softmax of fp16 scores is itself rounded to fp16 (ulp 2^-7 at |logprob| ~ 10)
        assert abs(lps[b].item() - ref_lp[b, t].item()) <= 1.2e-2, (b, lps[b].item(), ref_lp[b, t].item())
        assert rks[b].item() == int((ref_scores[b] > ref_scores[b, t]).sum()) + 1


def test_top_p_survivors_agree_up_to_ties_at_the_cut():
    import tgis_b200  # noqa: F401
    from tgis_b200 import pb
    B, V, S = 5, 32000, 8
    scores, history, pos = _case(B, V, S, 11)
    ps = [0.9, 0.5, 0.25, 0.99, 0.7]
    pbs = [_params(pb, temperature=1.0, top_p=p, top_k=(200 if i == 4 else 0), seed=i) for i, p in enumerate(ps)]
    ref_ch, dev_ch = _chooser(pbs), _chooser(pbs)
    _, ref_scores, _ = ref_ch(input_ids=history[:, :S], scores=scores.clone())
    ids, lps, _ = dev_ch.choose_on_device(history, pos, scores, True, False)
    torch.cuda.synchronize()
    probs = torch.softmax(scores.float(), -1)
    for b in range(B):
        keep_ref = torch.isfinite(ref_scores[b])
        # the kernel's surviving set = {x >= cut}: recover the cut from the reported log-probability of the chosen token
        t = ids[b].item()
        x = scores[b].float()
        logz = x[t] - lps[b]
        # survivors' mass implied by the kernel vs the torch chain's
        mass_dev = torch.exp(logz - torch.logsumexp(x if b != 4 else torch.where(x >= x.topk(200).values[-1], x, x.new_tensor(-math.inf)), 0)).item()
        ref_set = keep_ref
        base = probs[b] if b != 4 else torch.softmax(torch.where(x >= x.topk(200).values[-1], x, x.new_tensor(-math.inf)), -1)
        mass_ref = base[ref_set].sum().item()
        assert mass_dev >= ps[b] - 2e-3, (b, mass_dev)                    # at least the requested mass survives
        assert abs(mass_dev - mass_ref) <= 4e-3, (b, mass_dev, mass_ref)    # and the same mass as the torch chain, up to the tie group / fp16 cumsum


def test_sampling_is_reproducible_lockstep_and_follows_the_warped_distribution():
    import tgis_b200  # noqa: F401
    from tgis_b200 import pb
    V, S, N = 64, 4, 4000
    g = torch.Generator().manual_seed(5)
    row = (torch.randn(V, generator=g) * 1.5).half()
    scores = row.repeat(N, 1).to(DEV)
    history = torch.zeros(N, S + 2, dtype=torch.int64, device=DEV)
    pos = torch.full((N,), S, dtype=torch.int64, device=DEV)
    pbs = [_params(pb, temperature=0.8, top_k=12, seed=1000 + i) for i in range(N)]
    a, b = _chooser(pbs), _chooser(pbs)
    ids_a, _, _ = a.choose_on_device(history, pos, scores, False, False)
    ids_b, _, _ = b.choose_on_device(history, pos, scores, False, False)          # another "rank": same seeds, same draws
    first = ids_a.clone()
    assert torch.equal(first, ids_b)
    second, _, _ = a.choose_on_device(history, pos, scores, False, False)           # next step: the draw counter advanced
    torch.cuda.synchronize()
    assert not torch.equal(first, second)
    # distribution of the N independently seeded draws vs softmax of the warped row
    w = (row.float() / torch.tensor(0.8, dtype=torch.float16).float()).half().float()
    kth = w.topk(12).values[-1]
    p = torch.softmax(torch.where(w >= kth, w, w.new_tensor(-math.inf)), -1)
    counts = torch.bincount(torch.cat([first, second]).cpu(), minlength=V).float()
    assert counts[p == 0].sum() == 0
    exp = p * 2 * N
    chi2 = (((counts - exp) ** 2)[p > 0] / exp[p > 0]).sum().item()
    dof = int((p > 0).sum()) - 1
    assert chi2 < dof + 6 * math.sqrt(2 * dof), (chi2, dof)                          # ~6 sigma


def test_generate_token_samples_inside_the_graph_and_matches_the_op_by_op_path(tmp_path, monkeypatch):
    """A batch mixing greedy + logprobs, sampling with top-k / top-p and repetition penalty: the CUDA-graph replayed fused step
    (chooser kernel inside the graph) produces exactly the tokens, log-probabilities and ranks of the op-by-op path that calls the
    same kernel after an eager forward (B200_CUDA_GRAPHS=false)."""
    from tests.test_gpu_generate import _prompts, _setup, _text
    import tgis_b200  # noqa: F401
    from tgis_b200 import pb

    model, oracle, tok = _setup(tmp_path, None)
    prompts = _prompts(41, [5, 17, 9, 12], 512)
    n_new = 9

    def batch_pb():
        params = [_params(pb), _params(pb, temperature=0.9, top_k=20, seed=7), _params(pb, temperature=1.1, top_p=0.8, seed=8),
                  _params(pb, repetition_penalty=1.3, min_new_tokens=4)]
        reqs = [pb.Request(id=i, inputs=_text(p), input_length=len(p), max_output_length=n_new, parameters=params[i],
                           details=pb.RequestedDetails(logprobs=(i != 2), ranks=(i == 0))) for i, p in enumerate(prompts)]
        return pb.Batch(id=0, requests=reqs)

    def run():
        out = {i: [] for i in range(len(prompts))}
        with torch.inference_mode():
            batch, errs = model.batch_type.from_pb(batch_pb(), tok, torch.float16, model.device, None, None, True)
            assert not errs
            res = model.generate_token(batch, first=True)
            for step in range(n_new):
                for t in res[0]:
                    out[t.request_id].append((t.token_id, round(t.logprob, 3) if t.logprob else 0.0, t.rank))
                if step + 1 < n_new:  # the requests asked for n_new tokens: their rows of the history tensor end there
                    res = model.generate_token(batch)
        model.kv_cache_manager.free_sequences(batch.sequence_ids)
        return out

    fused = run()
    monkeypatch.setattr(model, "_can_fuse_greedy", lambda batch: False)  # op by op: eager forward, then the same chooser kernel
    eager = run()
    assert fused == eager
    assert all(r for _, _, r in fused[0]) and all(lp <= 0 for _, lp, _ in fused[1])
