"""Fused greedy decode for the families that run op by op from Python (GPT-NeoX, Santacoder, Falcon).

`FlashCausalLM._decode_fused_greedy` drives a model through `make_step` / `run_step`: the step writes logits and the arg-max
ids into the caller's buffers, the ids chain on the device, and from the third step of a stable batch the whole step is
replayed as a CUDA graph.  FlashLlama implements the protocol with the C++ step runtime (csrc/llama_step.cu); this mixin
implements it by enqueuing the family's ordinary decode forward, whose torch temporaries then come from the graph's private
pool.  EXPERIMENTAL: off unless B200_PY_FUSED_STEP=1 (or B200_NEOX_FUSED=1, the first spelling), not validated on a GPU yet.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch

from ...utils import _ops
from ...utils.paged import PagedKVState


@dataclass
class PythonStep:
    """What `make_step` hands to FlashCausalLM for these families: the caller's index tensors and output buffers."""
    T: int
    B: int
    max_s: int
    input_ids: torch.Tensor
    position_ids: torch.Tensor
    kv: PagedKVState
    logits: torch.Tensor
    next_ids: Optional[torch.Tensor]
    banned: Optional[torch.Tensor]  # per-row banned id of the arg-max (min_new_tokens EOS mask), set by the caller
    decode_marker: torch.Tensor
    banned_ids: int = 0             # the C struct's field of the same name (FlashCausalLM sets both)


class _NoScratch:
    version = 0  # FlashCausalLM keys its cached step on this; nothing here is ever re-allocated


def _enabled() -> bool:
    return os.environ.get("B200_PY_FUSED_STEP", "0") == "1" or os.environ.get("B200_NEOX_FUSED", "0") == "1"


class PythonFusedGreedy:
    """Mixin for a `...ForCausalLM` module: needs `self.model` (the backbone, called like the reference's forward) and
    `self.lm_head` (a TensorParallelHead)."""
    fused_greedy_enabled = _enabled()
    scratch = _NoScratch()

    def make_step(self, *, T: int, B: int, is_prefill: bool, max_s: int, input_ids, position_ids, kv: PagedKVState,
                  cu_seqlens=None, head_rows=None, logits=None, next_ids=None, inputs_embeds=None) -> PythonStep:
        if is_prefill or head_rows is not None or inputs_embeds is not None:
            raise NotImplementedError("the Python fused step is decode-only")
        # any non-None cu_seqlens_q selects the decode branch of the attention modules; the kernels never read it
        marker = torch.zeros(1, dtype=torch.int32, device=input_ids.device)
        return PythonStep(T=T, B=B, max_s=int(max_s), input_ids=input_ids, position_ids=position_ids, kv=kv, logits=logits,
                          next_ids=next_ids, banned=None, decode_marker=marker)

    def run_step(self, s: PythonStep) -> None:
        hidden, _ = self.model(s.input_ids, s.position_ids, None, s.decode_marker, s.max_s, None, s.kv)
        s.logits.copy_(self.lm_head.linear(hidden))  # this rank's vocab rows; FlashCausalLM gathers them when sharded
        if s.next_ids is not None:
            _ops().argmax(s.logits, s.banned, out=s.next_ids)
