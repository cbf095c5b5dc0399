"""The paged calling convention of the Llama graph.

Mirrors /root/reference/server/text_generation_server/models/custom_modeling/paged_llama_modeling.py:427-462
(`PagedLlamaForCausalLM`: `forward(input_ids, position_ids, cache_data, inputs_embeds=None, return_embeds=False)`,
`get_kv_cache_block_size`, `get_input_embeddings`) on top of the same C++ step runtime and kernels as the flash
convention (flash_llama_modeling.py of this package): the two conventions differ only in how the per-step indices arrive.
`cache_data` is a `PagedAttentionCacheData` (utils/paged.py), fms-extras' structure as the reference uses it:
  * prefill form (`is_filled()` false): `context_lengths` = cumulative prompt lengths [B + 1], one block-table row per sequence;
  * generation form: one row per query TOKEN (`block_mapping` [T, blocks], `context_lengths` [T]), which is also how the n
    tokens of each speculative candidate are verified in one forward (models/paged_causal_lm.py:481-562): the paged decode
    kernel treats every token as its own row over the candidate's blocks, attending its own prefix.
"""
from __future__ import annotations

from typing import Optional

import torch

from ...utils.paged import PagedAttentionCacheData, PagedKVState
from .flash_llama_modeling import FlashLlamaForCausalLM


class PagedLlamaForCausalLM(FlashLlamaForCausalLM):
    def forward(self, input_ids, position_ids, cache_data: Optional[PagedAttentionCacheData] = None,
                inputs_embeds: Optional[torch.Tensor] = None, return_embeds: bool = False, *flash_args, **flash_kwargs):
        if not isinstance(cache_data, PagedAttentionCacheData):
            # the flash convention (cu_seqlens, cu_seqlens_q, max_s, ...): same object serves both model classes
            return super().forward(input_ids, position_ids, cache_data, inputs_embeds, return_embeds, *flash_args, **flash_kwargs)
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        T = position_ids.shape[0]
        slots = cache_data.slot_mapping.reshape(-1)
        if slots.shape[0] != T:
            raise ValueError(f"cache_data.slot_mapping has {slots.shape[0]} entries for {T} tokens (flatten it first, utils/paged.py:100-108)")
        if cache_data.is_filled():
            ctx = cache_data.context_lengths.to(torch.int32).contiguous()
            table = cache_data.block_mapping.to(torch.int32).contiguous()
            if ctx.shape[0] != T or table.shape[0] != T:
                raise ValueError("generation form: one context length and one block-table row per query token")
            kv = PagedKVState(cache_data.sequence_ids, table, ctx, slots, table.shape[1])
            cu, B, prefill = None, T, False
        else:
            cu = cache_data.context_lengths.to(torch.int32).contiguous()
            B = cu.shape[0] - 1
            kv = PagedKVState(cache_data.sequence_ids, cache_data.block_mapping.to(torch.int32).contiguous(),
                              (cu[1:] - cu[:-1]).contiguous(), slots, cache_data.block_mapping.shape[1])
            prefill = True
        V_local = self.lm_head.linear.weight.shape[0]
        logits = torch.empty(T, V_local, dtype=torch.float16, device=self.device)
        s = self.make_step(T=T, B=B, is_prefill=prefill, max_s=int(cache_data.max_sequence_length), input_ids=input_ids,
                           position_ids=position_ids, kv=kv, cu_seqlens=cu, logits=logits, inputs_embeds=inputs_embeds)
        self.run_step(s, embed=inputs_embeds is None)
        logits = self._gather_logits(logits)
        if return_embeds:  # the final-norm output the head consumed (paged_llama_modeling.py:451-460)
            return logits, self.scratch.bufs["normed"][:T].clone()
        return logits
