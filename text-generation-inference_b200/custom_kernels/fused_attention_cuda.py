"""`custom_kernels.fused_attention_cuda.forward` (fused_attention_cuda.cu:113-242) over `b200_masked_softmax`."""
from typing import List, Optional, Tuple

import torch

from .. import ops


def forward(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, layer_past: Optional[List[torch.Tensor]],
            attention_mask: torch.Tensor, head_mask: Optional[torch.Tensor], inv_norm_factor: float, num_heads: int,
            use_cache: bool) -> Tuple[torch.Tensor, Optional[List[torch.Tensor]], torch.Tensor]:
    """query/key/value [B, h, q|kv, d]; attention_mask bool broadcastable to [B*h, q, kv] (True = masked).
    -> (context [B, q, h*d], present, attention_probs [B*h, q, kv])"""
    if layer_past is not None:  # :128-133
        key = torch.cat([layer_past[0], key], dim=2)
        value = torch.cat([layer_past[1], value], dim=2)
    present = [key, value] if use_cache else None
    B, _, q_length, d = query.shape
    kv_length = key.shape[2]
    bh = B * num_heads
    q = query.reshape(bh, q_length, d) * inv_norm_factor  # :150
    scores = torch.bmm(q, key.reshape(bh, kv_length, d).transpose(1, 2))  # :152
    mask = attention_mask.expand(B, num_heads, q_length, kv_length) if attention_mask.dim() == 4 else attention_mask
    probs = ops.masked_softmax(scores.view(bh * q_length, kv_length), mask.reshape(bh * q_length, kv_length))
    probs = probs.view(bh, q_length, kv_length)
    context = torch.bmm(probs, value.reshape(bh, kv_length, d))  # :234
    context = context.view(B, num_heads, q_length, d).permute(0, 2, 1, 3).reshape(B, q_length, num_heads * d)  # :237-239
    return context, present, probs
