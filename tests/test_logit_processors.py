"""The vectorised (heterogeneous) logit processors against one-request-at-a-time restatements, the way the reference tests
them (/root/reference/server/tests/test_logit_processors.py:47-161: batch of 2, vocabulary of 25, one no-op setting and one
active setting per processor).  The sequential side here is transformers' own processor where the reference uses it
(repetition penalty, temperature, top-k) and a plain per-row loop for the two warpers TGIS defines itself (top-p keeps the
smallest set whose mass reaches p, counted from the top; typical-p keeps the tokens whose surprise is closest to the
entropy until their mass reaches `mass`).  CPU only; `tests/golden/chooser.npz` pins the same classes bit for bit against
the reference implementation itself.
"""
import math

import pytest
import torch
from transformers.generation.logits_process import (RepetitionPenaltyLogitsProcessor, TemperatureLogitsWarper,
                                                    TopKLogitsWarper)

import tgis_b200  # noqa: F401
from tgis_b200.utils.logits_process import (HeterogeneousRepetitionPenaltyLogitsProcessor, HeterogeneousTemperatureLogitsWarper,
                                            HeterogeneousTopKLogitsWarper, HeterogeneousTopPLogitsWarper,
                                            HeterogeneousTypicalLogitsWarper)

B, V = 2, 25
INPUT_IDS = torch.tensor([[1, 2, 1, 3, 4, 6, 7, 1, 1, 1], [1, 7, 0, 3, 4, 6, 7, 1, 1, 1]], dtype=torch.long)
SCORES = torch.softmax(torch.rand((B, V), generator=torch.Generator().manual_seed(11), dtype=torch.float32), dim=-1)


def _same(rows, batched):
    assert len(rows) == batched.shape[0]
    for row, got in zip(rows, batched):
        assert torch.allclose(row.reshape(-1), got), (row, got)


def test_repetition_penalty_rows_agree():
    penalties = [1.0, 2.5]  # 1.0: no penalty
    batched = HeterogeneousRepetitionPenaltyLogitsProcessor(penalties, torch.float32, None)(INPUT_IDS, SCORES.clone())
    rows = [RepetitionPenaltyLogitsProcessor(penalty=p)(ids[None], s[None].clone()) for p, s, ids in zip(penalties, SCORES, INPUT_IDS)]
    _same(rows, batched)


def test_repetition_penalty_spares_the_excluded_id():
    """id_to_exclude (the pad id of a left-padded batch) keeps its score whenever the batch has more than one row"""
    proc = HeterogeneousRepetitionPenaltyLogitsProcessor([2.0, 2.0], torch.float32, None, id_to_exclude=1)
    out = proc(INPUT_IDS, SCORES.clone())
    assert torch.equal(out[:, 1], SCORES[:, 1]) and not torch.equal(out[:, 3], SCORES[:, 3])


def test_temperature_rows_agree():
    temperatures = [0.25, 1.0]
    batched = HeterogeneousTemperatureLogitsWarper(temperatures, torch.float32, None)(INPUT_IDS, SCORES.clone())
    rows = [TemperatureLogitsWarper(t)(None, s[None].clone()) if t != 1.0 else s for t, s in zip(temperatures, SCORES)]
    _same(rows, batched)


@pytest.mark.parametrize("top_k", [[0, 3], [1, 3], [V + 5, 2]])
def test_top_k_rows_agree(top_k):
    batched = HeterogeneousTopKLogitsWarper(top_k, None)(INPUT_IDS, SCORES.clone())
    rows = [TopKLogitsWarper(top_k=k)(None, s[None].clone()) if k != 0 else s for k, s in zip(top_k, SCORES)]  # 0 = off
    _same(rows, batched)


def _top_p_row(s, p):
    order = torch.argsort(s, descending=True)
    probs = torch.softmax(s, dim=-1)[order]
    out = torch.full_like(s, -math.inf)
    mass = 0.0
    for rank, idx in enumerate(order.tolist()):
        # a token goes when the mass of everything below it, itself included, is at most 1 - p; the best one always stays
        tail = 1.0 - mass
        if rank == 0 or tail > 1.0 - p + 1e-7:
            out[idx] = s[idx]
        mass += probs[rank].item()
    return out


def test_top_p_rows_agree():
    top_p = [0.9, 0.0]
    batched = HeterogeneousTopPLogitsWarper(top_p, torch.float32, None)(INPUT_IDS, SCORES.clone())
    _same([_top_p_row(s, p) for p, s in zip(top_p, SCORES)], batched)
    assert int(torch.isfinite(batched[1]).sum()) == 1  # p = 0 keeps min_tokens_to_keep = 1 token: the arg-max
    assert int(torch.argmax(batched[1])) == int(torch.argmax(SCORES[1]))


def _typical_row(s, mass):
    logp = torch.log_softmax(s, dim=-1)
    p = logp.exp()
    entropy = -(p * logp).sum()
    order = torch.argsort((-logp - entropy).abs(), descending=False, stable=True)
    out = torch.full_like(s, -math.inf)
    acc = 0.0
    for idx in order.tolist():  # closest surprise first, until the kept mass reaches `mass` (the crossing token is kept)
        out[idx] = s[idx]
        acc += p[idx].item()
        if acc >= mass:
            break
    return out


def test_typical_p_rows_agree():
    masses = [0.7, 0.9]
    batched = HeterogeneousTypicalLogitsWarper(masses, torch.float32, None)(INPUT_IDS, SCORES.clone())
    _same([_typical_row(s, m) for m, s in zip(masses, SCORES)], batched)


def test_filter_drops_rows_and_reports_no_ops():
    """`filter(indices)` keeps the surviving requests' settings and returns None once nothing is left to do
    (utils/logits_process.py:132-143 and the same contract in the four warpers)"""
    rep = HeterogeneousRepetitionPenaltyLogitsProcessor([1.0, 2.5], torch.float32, None)
    assert rep.filter([0]) is None
    temp = HeterogeneousTemperatureLogitsWarper([0.5, 1.0], torch.float32, None)
    kept = temp.filter([0])
    assert kept is temp and torch.allclose(kept(None, SCORES[:1].clone()), SCORES[:1] / 0.5)
    assert HeterogeneousTemperatureLogitsWarper([0.5, 1.0], torch.float32, None).filter([1]) is None
    topk = HeterogeneousTopKLogitsWarper([0, 3], None)
    assert topk.filter([0]) is None
    topk = HeterogeneousTopKLogitsWarper([0, 3], None).filter([1])
    assert int(torch.isfinite(topk(None, SCORES[1:].clone())).sum()) == 3
    assert HeterogeneousTopPLogitsWarper([1.0, 0.3], torch.float32, None).filter([0]) is None
    assert HeterogeneousTypicalLogitsWarper([1.0, 0.3], torch.float32, None).filter([0]) is None
