"""Next-token choosing and token-info extraction (caller side of the logits).

Behavioural mirror of /root/reference/server/text_generation_server/utils/tokens.py: `Sampling` (:32-41), `Greedy`
(:44-46), `HeterogeneousNextTokenChooser` (:161-333), `HeterogeneousSampling` (:336-384), `get_token_info` (:388-425),
`get_input_tokens_info` (:429-506).  Greedy batches build no warpers and arg-max the fp16 logits (:197-219) — here with
the library's arg-max kernel when the scores live on the GPU.
"""
from __future__ import annotations

import os
from itertools import chain, repeat
from typing import List, Optional, Tuple, Union

import torch

from .logits_process import (
    HeterogeneousRepetitionPenaltyLogitsProcessor,
    HeterogeneousTemperatureLogitsWarper,
    HeterogeneousTopKLogitsWarper,
    HeterogeneousTopPLogitsWarper,
    HeterogeneousTypicalLogitsWarper,
)
from .token_types import InputTokens, TokenInfo, TopToken

FP32_LOGITS = os.getenv("FP32_LOGITS_PROCESS") == "true"
INT_ZEROS, FLOAT_ZEROS, NONES = repeat(0), repeat(0.0), repeat(None)
SINGLE_ZERO, SINGLE_NONE, SINGLE_NAN = [0], [None], [float("nan")]


class Sampling:
    def __init__(self, seed: Optional[int] = None, device: str = "cpu"):
        self.generator = None if seed is None else torch.Generator(device).manual_seed(seed)
        self.seed = seed   # the fused device chooser keys its Philox stream with it ...
        self.draws = 0     # ... and counts the draws made so far (host mirror of the device counter)

    def __call__(self, logits):
        probs = torch.nn.functional.softmax(logits, -1)
        q = torch.empty_like(probs).exponential_(1, generator=self.generator)  # Gumbel-style draw without a host sync
        return probs.div_(q).argmax()


class Greedy:
    def __call__(self, logits):
        if logits.is_cuda and logits.dtype == torch.float16 and logits.dim() == 2 and logits.stride(1) == 1:
            from .. import ops
            return ops.argmax(logits)
        return logits.argmax(dim=-1)


class HeterogeneousSampling:
    def __init__(self, do_sample: List[bool], seeds: List[Optional[Union[int, Sampling]]], device):
        self.greedy_indices, self.sampling_mapping, self.samplings = [], {}, []
        for i, (sample, seed) in enumerate(zip(do_sample, seeds)):
            if sample:
                s = seed if isinstance(seed, Sampling) else Sampling(seed, device)
                self.sampling_mapping[i] = s
                self.samplings.append(s)
            else:
                self.greedy_indices.append(i)
                self.samplings.append(None)
        self.greedy = Greedy()

    def __call__(self, logits):
        out = torch.empty(logits.shape[0], dtype=torch.int64, device=logits.device)
        if self.greedy_indices:
            torch.argmax(logits, -1, out=out)
        for i, s in self.sampling_mapping.items():
            out[i] = s(logits[i])
        return out

    def filter(self, indices):
        greedy, mapping = [], {}
        for i, idx in enumerate(indices):
            if idx in self.sampling_mapping:
                mapping[i] = self.sampling_mapping[idx]
            else:
                greedy.append(i)
        self.greedy_indices, self.sampling_mapping = greedy, mapping
        self.samplings = [self.samplings[i] for i in indices]
        return self


class HeterogeneousNextTokenChooser:
    def __init__(self, temperature: List[float], top_k: List[float], top_p: List[float], typical_p: List[float],
                 seeds: List[Optional[Union[int, Sampling]]], repetition_penalty: List[float],
                 length_penalty: List[Optional[Tuple[int, float]]], min_new_tokens: List[int], return_logprobs: List[bool],
                 eos_token_id: Optional[int] = None, pad_token_id: Optional[int] = None, device=None, dtype=None,
                 current_tokens: Optional[List[int]] = None):
        warpers = []
        self.repetition_processor = (
            HeterogeneousRepetitionPenaltyLogitsProcessor(
                repetition_penalty, dtype, device, id_to_exclude=eos_token_id if eos_token_id == pad_token_id else None)
            if any(x != 1.0 for x in repetition_penalty) else None)
        do_sample = [x != 0.0 for x in temperature]
        if any(do_sample):
            if any(x != 1.0 for x in temperature):
                warpers.append(HeterogeneousTemperatureLogitsWarper([t if t != 0 else 1 for t in temperature], dtype, device))
            if any(x != 0 for x in top_k):
                warpers.append(HeterogeneousTopKLogitsWarper(top_k, device))
            if any(x < 1.0 for x in top_p):
                warpers.append(HeterogeneousTopPLogitsWarper(top_p, dtype, device))
            if any(x < 1.0 for x in typical_p):
                warpers.append(HeterogeneousTypicalLogitsWarper(typical_p, dtype, device))
            self.choice = HeterogeneousSampling(do_sample, seeds, device)
        else:
            self.choice = Greedy()
        self.warpers = warpers
        # per-request parameters as given: the fused device chooser (csrc/chooser.cu) takes them as arrays
        self.temperature, self.top_k, self.top_p, self.typical_p = list(temperature), list(top_k), list(top_p), list(typical_p)
        self.repetition_penalty = list(repetition_penalty)
        self._device = None  # DeviceChooser, built on first use
        self.eos_token_id, self.pad_token_id = eos_token_id, pad_token_id
        self.length_penalty = length_penalty
        self.min_new_tokens = min_new_tokens
        self.current_tokens = current_tokens if current_tokens is not None else [0] * len(do_sample)
        self.do_sample = do_sample
        self.dtype, self.device = dtype, device
        self.return_logprobs = return_logprobs

    @property
    def samplings(self):
        if isinstance(self.choice, Greedy):
            return [None] * len(self.do_sample)
        return self.choice.samplings

    @property
    def is_plain_greedy(self) -> bool:
        """True when the step needs nothing but arg-max (+ the min_new_tokens EOS mask): the fused device path."""
        return (isinstance(self.choice, Greedy) and self.repetition_processor is None and not self.warpers
                and not any(self.return_logprobs) and all(lp is None for lp in self.length_penalty))

    def eos_masked_rows(self) -> List[int]:
        return [i for i, (c, m) in enumerate(zip(self.current_tokens, self.min_new_tokens)) if c < m]

    # ------------------------------------------------------------------------------------------ fused device chooser
    def device_eligible(self, scores: torch.Tensor) -> bool:
        """One b200_choose_tokens launch can stand in for __call__: fp16 scores on the GPU, vocabulary within the kernel's bitmap,
        nobody asks for typical-p (it stays on the torch path) and logits are not forced to fp32."""
        return (scores.is_cuda and scores.dtype == torch.float16 and scores.dim() == 2 and scores.stride(1) == 1
                and scores.shape[1] <= 131072 and not FP32_LOGITS and all(x >= 1.0 for x in self.typical_p))

    def step_masks(self) -> Tuple[List[int], List[float]]:
        """Host bookkeeping of one step, exactly the loop at the top of __call__ (tokens.py:240-252): -> per row the id to mask
        (EOS while fewer than min_new_tokens were produced, else -1) and the length-penalty factor pow(decay, past) - 1 (else 0)."""
        banned, factors = [], []
        for idx, (cur, mn, lp) in enumerate(zip(self.current_tokens, self.min_new_tokens, self.length_penalty)):
            ban, fac = -1, 0.0
            if cur < mn:
                ban = self.eos_token_id
                self.current_tokens[idx] += 1
            elif lp is not None:
                past = cur - lp[0]
                if past > 0:
                    fac = pow(lp[1], past) - 1
                self.current_tokens[idx] += 1
            banned.append(ban)
            factors.append(fac)
        return banned, factors

    def device_chooser(self) -> "DeviceChooser":
        if self._device is None:
            self._device = DeviceChooser(self)
        return self._device

    def choose_on_device(self, all_input_ids: torch.Tensor, position_ids: torch.Tensor, scores: torch.Tensor,
                         want_logprobs: bool, want_ranks: bool, out_ids: Optional[torch.Tensor] = None):
        """-> (next_ids [B] int64, logprobs [B] float32 or None, ranks [B] int32 or None), all on the device.
        all_input_ids / position_ids: the batch's history tensor and the position of each row's input token."""
        dc = self.device_chooser()
        banned, factors = self.step_masks()
        dc.set_step(banned, factors)
        return dc.launch(all_input_ids, position_ids, scores, want_logprobs, want_ranks, out_ids)

    def __call__(self, input_ids: torch.Tensor, scores: torch.Tensor):
        if FP32_LOGITS:
            scores = scores.to(torch.float32)
        masked = []
        for idx, (cur, mn, lp) in enumerate(zip(self.current_tokens, self.min_new_tokens, self.length_penalty)):
            if cur < mn:
                masked.append(idx)
                self.current_tokens[idx] += 1
            elif lp is not None:
                past = cur - lp[0]
                if past > 0:
                    eos = scores[idx, self.eos_token_id]
                    scores[idx, self.eos_token_id] = eos + torch.abs(eos) * (pow(lp[1], past) - 1)
                self.current_tokens[idx] += 1
        if masked:
            if len(masked) == scores.shape[0]:
                scores[:, self.eos_token_id] = -float("inf")
            else:
                scores[torch.tensor(masked, device=scores.device), self.eos_token_id] = -float("inf")
        if self.repetition_processor is not None:
            scores = self.repetition_processor(input_ids, scores)
        for warper in self.warpers:
            scores = warper(input_ids, scores)
        next_ids = self.choice(scores)
        logprobs = torch.log_softmax(scores, -1) if any(self.return_logprobs) else NONES
        return next_ids, scores, logprobs

    @classmethod
    def from_pb(cls, pb, model_eos_token_id, model_pad_token_id, return_logprobs: List[bool], dtype, device,
                samplings: Optional[List[Sampling]] = None, current_tokens: Optional[List[int]] = None):
        seeds = samplings if samplings else [p.seed if p.HasField("seed") else None for p in pb]
        return cls(
            temperature=[p.temperature for p in pb],
            repetition_penalty=[p.repetition_penalty if p.HasField("repetition_penalty") else 1.0 for p in pb],
            top_k=[p.top_k for p in pb],
            top_p=[p.top_p if p.top_p > 0 else 1.0 for p in pb],
            typical_p=[p.typical_p if p.typical_p > 0 else 1.0 for p in pb],
            length_penalty=[(p.length_penalty.start_index, p.length_penalty.decay_factor) if p.HasField("length_penalty")
                            else None for p in pb],
            seeds=seeds, min_new_tokens=[p.min_new_tokens for p in pb], eos_token_id=model_eos_token_id,
            pad_token_id=model_pad_token_id, return_logprobs=return_logprobs, device=device, dtype=dtype,
            current_tokens=current_tokens)

    def filter(self, indices):
        self.temperature, self.top_k, self.top_p = ([lst[i] for i in indices] for lst in (self.temperature, self.top_k, self.top_p))
        self.typical_p = [self.typical_p[i] for i in indices]
        self.repetition_penalty = [self.repetition_penalty[i] for i in indices]
        self._device = None  # rebuilt from the lists (and the Sampling objects' draw counts) on next use
        if self.repetition_processor is not None:
            self.repetition_processor = self.repetition_processor.filter(indices)
        self.warpers = [w for w in (warper.filter(indices) for warper in self.warpers) if w is not None]
        self.do_sample = [self.do_sample[i] for i in indices]
        self.current_tokens = [self.current_tokens[i] for i in indices]
        self.min_new_tokens = [self.min_new_tokens[i] for i in indices]
        self.length_penalty = [self.length_penalty[i] for i in indices]
        self.return_logprobs = [self.return_logprobs[i] for i in indices]
        if any(self.do_sample):
            self.choice.filter(indices)
        else:
            self.choice = Greedy()
        return self


class DeviceChooser:
    """Device-side arrays of a HeterogeneousNextTokenChooser for b200_choose_tokens (csrc/chooser.cu): built once per batch
    composition, the per-step masks are refreshed with two small host-to-device copies, the draw counters live on the device so
    that the launch can be replayed from a CUDA graph."""

    def __init__(self, chooser: HeterogeneousNextTokenChooser):
        dev = chooser.device
        B = len(chooser.do_sample)
        self.chooser, self.B = chooser, B
        f32 = dict(dtype=torch.float32, device=dev)

        def as_scores_dtype(values):
            # the torch warpers hold their parameters in the scores' dtype (fp16 1.2 is 1.2002): round the same way
            return torch.tensor(values, dtype=torch.float16).to(**f32)

        self.temperature = as_scores_dtype([t if s else 0.0 for t, s in zip(chooser.temperature, chooser.do_sample)])
        self.top_k = torch.tensor([int(k) for k in chooser.top_k], dtype=torch.int32, device=dev) if any(chooser.top_k) else None
        # logits_process.py:203 keeps `1 - top_p` in the scores' dtype; hand the kernel the p that reproduces that complement
        self.top_p = (1.0 - (1 - torch.tensor(chooser.top_p, dtype=torch.float16)).to(**f32)) if any(p < 1.0 for p in chooser.top_p) else None
        self.rep = as_scores_dtype(chooser.repetition_penalty) if any(r != 1.0 for r in chooser.repetition_penalty) else None
        self.rep_exclude = chooser.eos_token_id if (chooser.eos_token_id is not None and chooser.eos_token_id == chooser.pad_token_id) else -1
        samplings = chooser.samplings
        self.sampled_rows = [i for i, smp in enumerate(samplings) if smp is not None]
        # a request without a seed draws a stream of its own; sharded deployments always send one (router/src/validation.rs:167-177)
        seeds = [((smp.seed if smp.seed is not None else 0x9E3779B97F4A7C15 ^ id(smp)) & 0x7FFFFFFFFFFFFFFF) if smp is not None else 0
                 for smp in samplings]
        self.seeds = torch.tensor(seeds, dtype=torch.int64, device=dev)
        self.counters = torch.tensor([smp.draws if smp is not None else 0 for smp in samplings], dtype=torch.int64, device=dev)
        self.banned = torch.full((B,), -1, dtype=torch.int64, device=dev)
        self.factors = torch.zeros(B, **f32)
        self._banned_host, self._factors_host = [-1] * B, [0.0] * B
        self.uses_length_penalty = any(lp is not None for lp in chooser.length_penalty)
        self.next_ids = torch.empty(B, dtype=torch.int64, device=dev)
        self.logprobs = torch.empty(B, **f32)
        self.ranks = torch.empty(B, dtype=torch.int32, device=dev)
        self.scratch = None

    def set_step(self, banned: List[int], factors: List[float]) -> None:
        if banned != self._banned_host:
            self.banned.copy_(torch.tensor(banned, dtype=torch.int64), non_blocking=True)
            self._banned_host = banned
        if factors != self._factors_host:
            self.factors.copy_(torch.tensor(factors, dtype=torch.float32), non_blocking=True)
            self._factors_host = factors
        for i in self.sampled_rows:  # host mirror of the device draw counters
            self.chooser.samplings[i].draws += 1

    def launch(self, all_input_ids, position_ids, scores, want_logprobs: bool, want_ranks: bool, out_ids=None,
               history_len_bias: int = 0):
        """history_len_bias: every row's history is all_input_ids[:, :max(position_ids) + bias] (the reference's
        `all_input_ids_tensor[:, :max_seqlen]`): 0 when position_ids already point at the slot of the token being chosen
        (after a prefill / op-by-op decode), 1 inside the fused step, where they still index the input token."""
        import ctypes

        from .. import _lib
        B, V = scores.shape
        assert B == self.B
        if self.scratch is None or self.scratch.shape != (B, V):
            self.scratch = torch.empty(B, V, dtype=torch.float16, device=scores.device)
        ids = out_ids if out_ids is not None else self.next_ids
        p = _lib.B200ChooserParams()
        p.logits, p.warped_scratch, p.ld, p.V, p.B = scores.data_ptr(), self.scratch.data_ptr(), scores.stride(0), V, B
        p.history_len_bias = history_len_bias
        p.temperature = self.temperature.data_ptr()
        p.top_k = self.top_k.data_ptr() if self.top_k is not None else None
        p.top_p = self.top_p.data_ptr() if self.top_p is not None else None
        if self.rep is not None:
            p.rep_penalty, p.history, p.history_stride = self.rep.data_ptr(), all_input_ids.data_ptr(), all_input_ids.stride(0)
            p.position_ids = position_ids.data_ptr()
        p.rep_exclude_id = self.rep_exclude
        p.banned_ids = self.banned.data_ptr()
        p.length_penalty_factor = self.factors.data_ptr() if self.uses_length_penalty else None
        p.eos_id = self.chooser.eos_token_id if self.chooser.eos_token_id is not None else -1
        p.seeds, p.counters = self.seeds.data_ptr(), self.counters.data_ptr()
        p.next_ids = ids.data_ptr()
        p.logprobs = self.logprobs.data_ptr() if want_logprobs else None
        p.ranks = self.ranks.data_ptr() if want_ranks else None
        _lib.check(_lib.load().b200_choose_tokens(ctypes.byref(p), torch.cuda.current_stream().cuda_stream), "choose_tokens")
        return ids, (self.logprobs if want_logprobs else None), (self.ranks if want_ranks else None)


def get_token_info(request, scores: torch.Tensor, next_token: torch.Tensor, logprobs: Optional[torch.Tensor]) -> TokenInfo:
    """tokens.py:388-425; `scores` [1, V]."""
    next_token = int(next_token.item()) if isinstance(next_token, torch.Tensor) else int(next_token)
    info = TokenInfo(request_id=request.id, token_id=next_token)
    if logprobs is not None:
        info.logprob = logprobs[-1, next_token].item()
    top_n_req = request.details.top_n_toks
    if top_n_req:
        flat = scores[-1]
        top_n = min(top_n_req, flat.size(-1))
        nth = flat.topk(top_n).values[-1]
        torch.nan_to_num_(nth, neginf=torch.finfo(flat.dtype).min)
        idx = (flat >= nth).nonzero().squeeze(-1)[:(top_n * 4)]
        info.top_tokens = [TopToken(token_id=t.item()) for t in idx] if logprobs is None else _sort(
            [TopToken(token_id=t.item(), logprob=logprobs[-1, t].item()) for t in idx])
    if request.details.ranks:
        info.rank = int((scores > scores[0, next_token]).sum() + 1)
    return info


def get_input_tokens_info(request, input_token_ids, all_input_logits) -> InputTokens:
    """tokens.py:429-506."""
    return_logprobs = request.details.logprobs
    if return_logprobs:
        all_logprobs = torch.log_softmax(all_input_logits, -1)
        input_logprobs = all_logprobs.gather(1, input_token_ids[1:].unsqueeze(-1))
        logprobs_gen = chain(SINGLE_NAN, input_logprobs.squeeze(-1))
    else:
        logprobs_gen = FLOAT_ZEROS
    if request.details.ranks:
        if return_logprobs:
            ranks_gen = chain(SINGLE_ZERO, ((all_logprobs > input_logprobs).sum(dim=1) + 1))
        else:
            input_logits = all_input_logits.gather(1, input_token_ids[1:].unsqueeze(-1))
            ranks_gen = chain(SINGLE_ZERO, ((all_input_logits > input_logits).sum(dim=1) + 1))
    else:
        ranks_gen = INT_ZEROS
    top_n = request.details.top_n_toks
    if top_n:
        top_n = min(top_n, all_input_logits.size(-1))
        nth = torch.topk(all_input_logits, top_n).values[..., -1, None]
        diff = all_input_logits >= nth
        idx = [diff[i].nonzero().squeeze(-1)[:top_n * 4] for i in range(diff.shape[0])]
        if return_logprobs:
            if idx:
                combined = torch.nn.utils.rnn.pad_sequence(idx, batch_first=True)
                lp = all_logprobs.gather(1, combined)
                topn_gen = chain(SINGLE_NONE, ((idx[i], lp[i][:len(idx[i])]) for i in range(len(idx))))
            else:
                topn_gen = SINGLE_NONE
        else:
            topn_gen = chain(SINGLE_NONE, idx)
    else:
        topn_gen = NONES
    toks = []
    for tok_id, logprob, rank, top in zip(input_token_ids, logprobs_gen, ranks_gen, topn_gen):
        if top is None:
            tts = None
        elif not return_logprobs:
            tts = [TopToken(int(t)) for t in top]
        else:
            tts = _sort([TopToken(int(t), float(l)) for t, l in zip(*top)])
        toks.append(TokenInfo(token_id=int(tok_id), logprob=float(logprob), rank=int(rank), top_tokens=tts))
    return InputTokens(request_id=request.id, tokens=toks)


def _sort(tts: List[TopToken]) -> List[TopToken]:
    tts.sort(reverse=True)
    return tts
