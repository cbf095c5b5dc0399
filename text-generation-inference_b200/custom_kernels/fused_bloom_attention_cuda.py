"""`custom_kernels.fused_bloom_attention_cuda.forward` (fused_bloom_attention_cuda.cu) over `b200_masked_softmax`:
fused-QKV split, ALiBi added through `baddbmm(alibi, q, k^T, beta, inv_norm_factor)`, masked softmax, P.V, head merge."""
from typing import List, Optional, Tuple

import torch

from .. import ops


def forward(fused_qkv: torch.Tensor, layer_past: Optional[List[torch.Tensor]], alibi: torch.Tensor, attention_mask: torch.Tensor,
            head_mask: Optional[torch.Tensor], beta: float, inv_norm_factor: float, num_heads: int, use_cache: bool
            ) -> Tuple[torch.Tensor, Optional[List[torch.Tensor]], torch.Tensor]:
    """fused_qkv [B, q, 3*H] (per head [q | k | v]); alibi [B*h, 1, kv]; layer_past = (key [B*h, d, past], value [B*h, past, d]).
    -> (context [B, q, H], present, attention_probs [B*h, q, kv])"""
    B, q_length, three_h = fused_qkv.shape
    d = three_h // (3 * num_heads)
    bh = B * num_heads
    qkv = fused_qkv.view(B, q_length, num_heads, 3 * d)
    query, key, value = qkv.split(d, dim=-1)
    query = query.transpose(1, 2).reshape(bh, q_length, d)
    key = key.permute(0, 2, 3, 1).reshape(bh, d, q_length)
    value = value.transpose(1, 2).reshape(bh, q_length, d)
    if layer_past is not None:
        key = torch.cat([layer_past[0], key], dim=2)
        value = torch.cat([layer_past[1], value], dim=1)
    present = [key, value] if use_cache else None
    kv_length = key.shape[2]
    scores = alibi.baddbmm(query, key, beta=beta, alpha=inv_norm_factor)
    mask = attention_mask.expand(B, num_heads, q_length, kv_length) if attention_mask.dim() == 4 else attention_mask
    probs = ops.masked_softmax(scores.view(bh * q_length, kv_length), mask.reshape(bh * q_length, kv_length)).view(bh, q_length, kv_length)
    context = torch.bmm(probs, value)
    context = context.view(B, num_heads, q_length, d).permute(0, 2, 1, 3).reshape(B, q_length, num_heads * d)
    return context, present, probs
