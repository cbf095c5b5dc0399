"""Per-request (heterogeneous) logit processors applied in one vectorised pass over the [B, V] scores.

Behavioural mirror of /root/reference/server/text_generation_server/utils/logits_process.py:93-402
(repetition penalty :93-143, temperature :146-176, top-p :179-238, top-k :241-317, typical-p :320-402): same
constructor arguments, in-place semantics, `filter(indices)` contract (returns None when the processor becomes a no-op).
These are caller-side of the hot path; they run as torch ops on the GPU.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch


class HeterogeneousRepetitionPenaltyLogitsProcessor:
    def __init__(self, penalty: List[float], dtype, device, id_to_exclude: Optional[int] = None):
        self.penalty = penalty
        self.penalty_tensor = torch.tensor(penalty, dtype=dtype, device=device).unsqueeze(1)
        self.id_to_exclude = id_to_exclude

    def __call__(self, input_ids: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
        exclude = self.id_to_exclude is not None and input_ids.shape[0] != 1
        saved = scores[:, self.id_to_exclude].clone() if exclude else None
        seen = torch.gather(scores, 1, input_ids)
        seen = torch.where(seen < 0, seen * self.penalty_tensor, seen / self.penalty_tensor)
        scores.scatter_(1, input_ids, seen)
        if exclude:
            scores[:, self.id_to_exclude] = saved
        return scores

    def filter(self, indices):
        self.penalty = [self.penalty[i] for i in indices]
        if all(x == 1.0 for x in self.penalty):
            return None
        self.penalty_tensor = self.penalty_tensor[indices]
        return self


class HeterogeneousTemperatureLogitsWarper:
    def __init__(self, temperature: List[float], dtype, device):
        self.temperature = temperature
        self.temperature_tensor = torch.tensor(temperature, dtype=dtype, device=device).unsqueeze(1)

    def __call__(self, input_ids, scores):
        scores.div_(self.temperature_tensor)
        return scores

    def filter(self, indices):
        self.temperature = [self.temperature[i] for i in indices]
        if all(x == 1.0 for x in self.temperature):
            return None
        self.temperature_tensor = self.temperature_tensor[indices]
        return self


class HeterogeneousTopPLogitsWarper:
    def __init__(self, top_p: List[float], dtype, device, filter_value: float = -math.inf, min_tokens_to_keep: int = 1):
        self.top_p = top_p
        self.top_p_opposite = 1 - torch.tensor(top_p, dtype=dtype, device=device).unsqueeze(1)
        self.filter_value = filter_value
        self.min_tokens_to_keep = min_tokens_to_keep

    def __call__(self, input_ids, scores):
        sorted_logits, sorted_indices = torch.sort(scores, descending=False)
        cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        remove = cum <= self.top_p_opposite
        remove[..., -self.min_tokens_to_keep:] = 0
        return scores.masked_fill_(remove.scatter(1, sorted_indices, remove), self.filter_value)

    def filter(self, indices):
        self.top_p = [self.top_p[i] for i in indices]
        if all(x == 1.0 for x in self.top_p):
            return None
        self.top_p_opposite = self.top_p_opposite[indices]
        return self


class HeterogeneousTopKLogitsWarper:
    def __init__(self, top_k: List[int], device, filter_value: float = -math.inf, min_tokens_to_keep: int = 1):
        self.top_k = top_k
        self.max_top_k = max(top_k)
        self.top_k_tensor = torch.tensor([max(x - 1, min_tokens_to_keep - 1) for x in top_k], dtype=torch.int64,
                                         device=device).unsqueeze(1)
        disabled = [x == 0 for x in top_k]  # 0 disables top-k for that request
        self.top_k_disabled_mask = torch.tensor(disabled, dtype=torch.bool, device=device).view(-1, 1) if any(disabled) else None
        self.filter_value = filter_value

    def __call__(self, input_ids, scores):
        if scores.size(-1) < self.max_top_k:
            # the reference clamps the 0-based index to V (logits_process.py:280-282) and then indexes out of bounds for
            # k > V; V - 1 is what its comment intends ("clamp or the warper will fail")
            max_top_k = scores.size(-1)
            top_k = torch.clamp_max(self.top_k_tensor, max_top_k - 1)
        else:
            max_top_k, top_k = self.max_top_k, self.top_k_tensor
        kth = torch.gather(torch.topk(scores, max_top_k).values, 1, top_k)
        if self.top_k_disabled_mask is not None:
            kth.masked_fill_(self.top_k_disabled_mask, self.filter_value)
        scores.masked_fill_(scores < kth, self.filter_value)
        return scores

    def filter(self, indices):
        self.top_k = [self.top_k[i] for i in indices]
        disabled = [x == 0 for x in self.top_k]
        if all(disabled):
            return None
        self.top_k_tensor = self.top_k_tensor[indices]
        self.max_top_k = max(self.top_k)
        if self.top_k_disabled_mask is not None:
            self.top_k_disabled_mask = self.top_k_disabled_mask[indices] if any(disabled) else None
        return self


class HeterogeneousTypicalLogitsWarper:
    def __init__(self, mass: List[float], dtype, device, filter_value: float = -math.inf, min_tokens_to_keep: int = 1):
        self.mass = mass
        self.mass_tensor = torch.tensor(mass, dtype=dtype, device=device).unsqueeze(1)
        disabled = [x == 1.0 for x in mass]  # 1.0 disables typical-p for that request
        self.disabled_mask = torch.tensor(disabled, dtype=torch.bool, device=device) if any(disabled) else None
        self.filter_value = filter_value
        self.min_tokens_to_keep = min_tokens_to_keep

    def __call__(self, input_ids, scores):
        logp = torch.nn.functional.log_softmax(scores, dim=-1)
        ent = -(logp * torch.exp(logp)).nansum(-1, keepdim=True)
        shifted = torch.abs((-logp) - ent)
        sorted_scores, sorted_indices = torch.sort(shifted, descending=False)
        cum = scores.gather(-1, sorted_indices).softmax(dim=-1).cumsum(dim=-1)
        last_ind = (cum < self.mass_tensor).sum(dim=1)
        last_ind.clamp_(max=sorted_scores.shape[-1] - 1)
        if self.disabled_mask is not None:
            last_ind.masked_fill_(self.disabled_mask, scores.shape[-1] - 1)
        remove = sorted_scores > sorted_scores.gather(1, last_ind.view(-1, 1))
        if self.min_tokens_to_keep > 1:
            remove[..., : self.min_tokens_to_keep] = 0
        return scores.masked_fill_(remove.scatter(1, sorted_indices, remove), self.filter_value)

    def filter(self, indices):
        self.mass = [self.mass[i] for i in indices]
        disabled = [x == 1.0 for x in self.mass]
        if all(disabled):
            return None
        self.mass_tensor = self.mass_tensor[indices]
        if self.disabled_mask is not None:
            self.disabled_mask = self.disabled_mask[indices] if any(disabled) else None
        return self
