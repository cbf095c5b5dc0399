"""The attention op boundary of the flash models.

Mirrors /root/reference/server/text_generation_server/utils/flash_attn.py:43-127: same positional signature
`attention(q, k, v, cu_seqlens, max_s, softmax_scale, cu_seqlens_q=None, max_s_q=None, causal=True)`.
Prefill (cu_seqlens_q is None): varlen causal attention over fresh q/k/v.  Decode: q holds one token per sequence and
`k` is a `PagedKVLayer` view of the block pool (block table + context lengths) instead of a contiguous cache slice.
There is no sm75/sm8x/sm90 gate (:8-40): this library targets sm_100a only and raises without it.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _ops


@dataclass
class PagedKVLayer:
    k_pool: torch.Tensor  # [num_blocks, n_kv, 16, d] swizzled
    v_pool: torch.Tensor
    block_table: torch.Tensor  # [B, max_blocks] int32
    context_lens: torch.Tensor  # [B] int32
    max_context_len: int


def attention(q, k, v, cu_seqlens, max_s, softmax_scale, cu_seqlens_q=None, max_s_q=None, causal=True,
              out: Optional[torch.Tensor] = None):
    if cu_seqlens_q is None:
        return _ops().attn_prefill_varlen(q, k, v, cu_seqlens, max_s, softmax_scale, causal, out)
    if not isinstance(k, PagedKVLayer):
        raise TypeError("decode attention reads the paged KV pool: pass a PagedKVLayer as `k`")
    assert max_s_q in (None, 1), "decode attention handles one query token per sequence"
    return _ops().attn_decode_paged(q, k.k_pool, k.v_pool, k.block_table, k.context_lens, k.max_context_len,
                                    softmax_scale, k.k_pool.shape[1], out)
