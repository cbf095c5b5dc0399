"""Ragged (un-padded) continuous-batching model runtime over the paged KV cache.

Mirrors /root/reference/server/text_generation_server/models/flash_causal_lm.py: `FlashCausalLMBatch`
(:28-62) with `from_pb` (:68-194), `concatenate` (:197-285), `prune` (:291-353); `FlashCausalLM.generate_token`
(:405-460) and the token post-processing (:462-588) — same signatures, same in-place batch mutation, same results.
What changes is storage only: KV lives in the model's block pool (`kv_cache_manager`, like the reference's
PagedCausalLM, models/paged_causal_lm.py:338-353), so
  * `concatenate` / `prune` edit block-table rows instead of copying KV (:279, :321-327 in the reference),
  * there is no per-step KV re-pack (:439-447),
  * `batch.sequence_ids` lets the server free a finished batch (server.py:233-249).
"""
from __future__ import annotations

import ctypes
import logging
import time
from dataclasses import dataclass
from typing import Any, List, Optional, Tuple, Type, Union

import torch

from .. import _lib
from .. import pb as generate_pb2
from ..utils.paged import OutOfBlocks, PagedKVCacheManager, PagedKVState
from ..utils.token_types import InputTokens, TokenInfo
from ..utils.tokens import HeterogeneousNextTokenChooser, get_input_tokens_info, get_token_info
from .model import Model
from .types import Batch, GenerateError

USE_CUDA_GRAPHS = __import__("os").getenv("B200_CUDA_GRAPHS", "true").lower() != "false"
FUSED_CHOOSER = __import__("os").getenv("B200_FUSED_CHOOSER", "1") != "0"


@dataclass
class FlashCausalLMBatch(Batch):
    batch_id: int
    requests: List[Any]
    input_ids: Optional[torch.Tensor]       # [sum(seq_lengths)] at prefill, [B] in decode
    position_ids: torch.Tensor
    inputs_embeds: Optional[torch.Tensor]
    cu_seqlens: torch.Tensor                # int32 [B+1]
    cu_seqlens_q: Optional[torch.Tensor]    # int32 arange(B+1), decode only
    past_key_values: Optional[PagedKVState]
    max_seqlen: int
    all_input_ids_tensor: torch.Tensor      # [B, max(total_lengths)] int64
    input_lengths: List[int]
    total_lengths: List[int]
    pad_token_id: int
    next_token_chooser: HeterogeneousNextTokenChooser
    kv_cache_manager: Optional[PagedKVCacheManager] = None

    def get_id(self) -> int:
        return self.batch_id

    @property
    def sequence_ids(self) -> List[int]:
        return self.past_key_values.sequence_ids if self.past_key_values is not None else []

    def __len__(self):
        return len(self.requests)

    @classmethod
    def from_pb(cls, pb, tokenizer, dtype: torch.dtype, device: torch.device, embeddings_lookup: Optional,
                prefix_cache: Optional, use_position_ids: bool = True) -> Tuple[Optional["FlashCausalLMBatch"], List[GenerateError]]:
        errors: List[GenerateError] = []
        batch_inputs, input_lengths, total_lengths, cu_seqlens = [], [], [], [0]
        max_seqlen = 0
        requests = []
        prefixes = {}  # index among the accepted requests -> prefix embeddings [P, H] (prompt tuning, :97-107)
        for r in pb.requests:
            input_length = r.input_length
            if r.prefix_id:
                try:
                    if prefix_cache is None:
                        raise KeyError(r.prefix_id)  # no prompt store configured (prompt_cache.py is outside this library)
                    prefix = prefix_cache.get(r.prefix_id)
                except Exception:
                    message = f"Prefix lookup error for request #{r.id}, prefix id {r.prefix_id}"
                    logging.error(message)
                    errors.append(GenerateError(request_id=r.id, message=message))  # excluded from the batch, reported
                    continue
                prefixes[len(requests)] = prefix
                input_length += prefix.shape[0]  # from here on the request's length includes its prefix
            requests.append(r)
            batch_inputs.append(r.inputs)
            input_lengths.append(input_length)
            max_seqlen = max(max_seqlen, input_length)
            total_lengths.append(input_length + r.max_output_length)
            cu_seqlens.append(cu_seqlens[-1] + input_length)
        if not requests:
            return None, errors

        tokenized = tokenizer(batch_inputs, truncation=True, max_length=max_seqlen, return_token_type_ids=False)["input_ids"]
        all_input_ids_tensor = torch.full((len(requests), max(total_lengths)), tokenizer.pad_token_id, dtype=torch.int64)
        input_ids, position_ids, params, return_logprobs = [], [], [], []
        for i, (r, toks, input_length) in enumerate(zip(requests, tokenized, input_lengths)):
            if r.truncate:
                toks = toks[-r.input_length:]
                if getattr(tokenizer, "add_bos_token", False):
                    toks[0] = tokenizer.bos_token_id
            if len(toks) != r.input_length:
                raise ValueError(f"request #{r.id}: input_length {r.input_length} but {len(toks)} tokens after tokenization")
            # a prefix occupies the leading positions as pad ids (:147-151); its embeddings replace them below
            all_input_ids_tensor[i, input_length - r.input_length:input_length] = torch.tensor(toks, dtype=torch.int64)
            input_ids.append(all_input_ids_tensor[i, :input_length].clone())
            params.append(r.parameters)
            return_logprobs.append(r.details.logprobs)
            position_ids.append(torch.arange(0, input_length))
        flat_ids = torch.cat(input_ids).to(device, non_blocking=True)
        inputs_embeds = None
        if prefixes:  # every request goes in as embeddings as soon as one has a prefix (:157-168)
            if embeddings_lookup is None:
                raise ValueError("requests with a prompt prefix need the model's input embedding")
            inputs_embeds = embeddings_lookup(flat_ids)
            for i, prefix in prefixes.items():
                inputs_embeds[cu_seqlens[i]:cu_seqlens[i] + prefix.shape[0], :] = prefix.to(inputs_embeds)
            flat_ids = None
        chooser = HeterogeneousNextTokenChooser.from_pb(
            pb=params, model_eos_token_id=getattr(tokenizer, "model_eos_token_id", tokenizer.eos_token_id),
            model_pad_token_id=tokenizer.pad_token_id, return_logprobs=return_logprobs, dtype=dtype, device=device)
        return cls(
            batch_id=pb.id, requests=requests,
            input_ids=flat_ids, inputs_embeds=inputs_embeds,
            position_ids=torch.cat(position_ids).to(device, non_blocking=True),
            cu_seqlens=torch.tensor(cu_seqlens, dtype=torch.int32, device=device), cu_seqlens_q=None, max_seqlen=max_seqlen,
            past_key_values=None, input_lengths=input_lengths, total_lengths=total_lengths,
            all_input_ids_tensor=all_input_ids_tensor.to(device, non_blocking=True), next_token_chooser=chooser,
            pad_token_id=tokenizer.pad_token_id), errors

    @classmethod
    def concatenate(cls, batches: List["FlashCausalLMBatch"]) -> "FlashCausalLMBatch":
        """:197-285.  KV is untouched: block tables are stacked row-wise (columns zero-padded)."""
        first = batches[0]
        device = first.cu_seqlens_q.device
        requests, input_lengths, total_lengths, params = [], [], [], []
        cur_tokens, samplings, ret_logprobs = [], [], []
        input_ids, position_ids, seq_ids, tables, ctx = [], [], [], [], []
        new_bs = sum(len(b) for b in batches)
        max_total = max(t for b in batches for t in b.total_lengths)
        all_ids = first.all_input_ids_tensor.new_full((new_bs, max_total), first.pad_token_id)
        start, max_seqlen = 0, 0
        max_cols = max(b.past_key_values.block_table.shape[1] for b in batches)
        for b in batches:
            requests.extend(b.requests)
            input_lengths.extend(b.input_lengths)
            total_lengths.extend(b.total_lengths)
            params.extend(r.parameters for r in b.requests)
            cur_tokens.extend(b.next_token_chooser.current_tokens)
            samplings.extend(b.next_token_chooser.samplings)
            ret_logprobs.extend(b.next_token_chooser.return_logprobs)
            input_ids.append(b.input_ids)
            position_ids.append(b.position_ids)
            kv = b.past_key_values
            seq_ids.extend(kv.sequence_ids)
            bt = kv.block_table
            if bt.shape[1] < max_cols:
                bt = torch.nn.functional.pad(bt, (0, max_cols - bt.shape[1]))
            tables.append(bt)
            ctx.append(kv.context_lens)
            b.past_key_values = None
            end = start + len(b)
            all_ids[start:end, :b.all_input_ids_tensor.shape[1]] = b.all_input_ids_tensor
            start = end
            max_seqlen = max(max_seqlen, b.max_seqlen)
        fc = first.next_token_chooser
        chooser = HeterogeneousNextTokenChooser.from_pb(
            pb=params, model_eos_token_id=fc.eos_token_id, model_pad_token_id=fc.pad_token_id, return_logprobs=ret_logprobs,
            dtype=fc.dtype, device=fc.device, samplings=samplings, current_tokens=cur_tokens)
        lens_t = torch.tensor([0] + input_lengths, dtype=torch.int32, device=device)
        kv = PagedKVState(sequence_ids=seq_ids, block_table=torch.cat(tables).contiguous(), context_lens=torch.cat(ctx),
                          slot_mapping=torch.empty(new_bs, dtype=torch.int64, device=device), max_blocks=max_cols)
        return FlashCausalLMBatch(
            batch_id=first.batch_id, requests=requests, input_ids=torch.cat(input_ids), inputs_embeds=None,
            position_ids=torch.cat(position_ids), cu_seqlens=torch.cumsum(lens_t, 0, dtype=torch.int32),
            cu_seqlens_q=torch.arange(new_bs + 1, device=device, dtype=torch.int32), max_seqlen=max_seqlen, past_key_values=kv,
            input_lengths=input_lengths, total_lengths=total_lengths, all_input_ids_tensor=all_ids, next_token_chooser=chooser,
            pad_token_id=first.pad_token_id, kv_cache_manager=first.kv_cache_manager)

    @classmethod
    def prune(cls, batch: "FlashCausalLMBatch", completed_ids: List[int]) -> Optional["FlashCausalLMBatch"]:
        """:291-353.  Completed sequences give their blocks back; survivors keep theirs in place."""
        if not completed_ids:
            return batch
        keep = Model.get_indices_to_keep(batch.requests, completed_ids)
        kv, mgr = batch.past_key_values, batch.kv_cache_manager
        if kv is not None and mgr is not None:
            keep_set = set(keep)
            mgr.free_sequences([s for i, s in enumerate(kv.sequence_ids) if i not in keep_set])
        if len(keep) == 0:
            batch.past_key_values = None
            return None
        idx = torch.tensor(keep, dtype=torch.int64, device=batch.position_ids.device)
        batch.input_lengths = [batch.input_lengths[i] for i in keep]
        batch.total_lengths = [batch.total_lengths[i] for i in keep]
        batch.requests = [batch.requests[i] for i in keep]
        batch.next_token_chooser = batch.next_token_chooser.filter(keep)
        batch.max_seqlen = max(batch.input_lengths)
        batch.input_ids = batch.input_ids[idx]
        batch.position_ids = batch.position_ids[idx]
        batch.all_input_ids_tensor = batch.all_input_ids_tensor[idx, :max(batch.total_lengths)]
        lens_t = torch.tensor([0] + batch.input_lengths, dtype=torch.int32, device=idx.device)
        batch.cu_seqlens = torch.cumsum(lens_t, 0, dtype=torch.int32)
        batch.cu_seqlens_q = batch.cu_seqlens_q[:len(keep) + 1]
        if kv is not None:
            kv.sequence_ids = [kv.sequence_ids[i] for i in keep]
            kv.block_table = kv.block_table[idx].contiguous()
            kv.context_lens = kv.context_lens[idx].contiguous()
            kv.slot_mapping = kv.slot_mapping[:len(keep)]
        return batch


class FlashCausalLM(Model):
    def __init__(self, model_name: str, revision: Optional[str], deployment_framework: str, dtype: torch.dtype,
                 quantize: Optional[str], model_config: Union[Any] = None, auto_model_class=None,
                 max_sequence_length: Optional[int] = None, engine=None, num_kv_blocks: Optional[int] = None):
        if not torch.cuda.is_available():
            raise NotImplementedError("FlashCausalLM is only available on GPU")
        if engine is None:
            from ..inference_engine import InferenceEngine
            engine = InferenceEngine(model_name, auto_model_class, dtype, quantize, model_config, max_sequence_length)
        super().__init__(engine, dtype, max_sequence_length)
        self.use_position_ids = True
        cfg = self.model.config
        if getattr(cfg, "pad_token_id", None) is not None:
            self.tokenizer.pad_token_id = cfg.pad_token_id
        elif self.tokenizer.pad_token_id is None:
            if getattr(cfg, "eos_token_id", None) is not None:
                self.tokenizer.pad_token_id = cfg.eos_token_id
            elif self.tokenizer.eos_token_id is not None:
                self.tokenizer.pad_token_id = self.tokenizer.eos_token_id
            else:
                self.tokenizer.add_special_tokens({"pad_token": "[PAD]"})
        import os
        if num_kv_blocks is None and os.getenv("KV_CACHE_MANAGER_NUM_GPU_BLOCKS"):
            num_kv_blocks = int(os.environ["KV_CACHE_MANAGER_NUM_GPU_BLOCKS"])  # paged_causal_lm.py:310-311
        world = getattr(engine, "world_size", 1)
        kv_heads, kv_world = getattr(cfg, "num_key_value_heads", None) or cfg.num_attention_heads, world
        layout = getattr(type(self.model), "kv_cache_layout", None)
        if layout is not None:  # multi-query / grouped families say how many KV heads there are and whether they shard
            kv_heads, kv_world = layout(cfg, world)
        self.kv_cache_manager = PagedKVCacheManager(
            cfg.num_hidden_layers, cfg.num_attention_heads, cfg.hidden_size, kv_heads=kv_heads,
            tensor_parallel_size=kv_world, dtype=dtype, device=self.device, total_num_gpu_blocks=num_kv_blocks, block_size=16)
        self.model.kv_cache_manager = self.kv_cache_manager

    @property
    def batch_type(self) -> Type[FlashCausalLMBatch]:
        return FlashCausalLMBatch

    # --------------------------------------------------------------------------------------------- generate_token
    def generate_token(self, batch: FlashCausalLMBatch, first: bool = False, for_concat: bool = False
                       ) -> Tuple[List[TokenInfo], Optional[List[InputTokens]], List[GenerateError], int]:
        mgr = self.kv_cache_manager
        batch.kv_cache_manager = mgr
        B = len(batch)
        if first:
            reserve = [r.max_output_length for r in batch.requests]
            sids = mgr.allocate_tokens(batch.input_lengths, reserve_tokens=reserve)  # OutOfBlocks -> RESOURCE_EXHAUSTED
            kv = PagedKVState(
                sequence_ids=sids, block_table=mgr.block_table_tensor(sids),
                context_lens=torch.tensor(batch.input_lengths, dtype=torch.int32, device=self.device),
                slot_mapping=mgr.slot_mapping_for(sids, [0] * B, batch.input_lengths), max_blocks=0)
            batch.past_key_values = kv
            want_all = any(r.details.input_toks for r in batch.requests)
            head_rows = None if want_all else (batch.cu_seqlens[1:] - 1).to(torch.int64)
            start_time = time.time_ns()
            out, _ = self.model.forward(batch.input_ids, batch.position_ids, batch.cu_seqlens, None, batch.max_seqlen,
                                        batch.inputs_embeds, kv, None, head_rows)
            forward_time_ns = time.time_ns() - start_time
            generated, input_infos, errs = self._process_prefill(batch, out, all_rows=want_all)
        else:
            kv = batch.past_key_values
            if kv.slot_mapping.shape[0] < B:
                kv.slot_mapping = torch.empty(B, dtype=torch.int64, device=self.device)
            if self._can_fuse_greedy(batch):
                generated, errs, forward_time_ns = self._decode_fused_greedy(batch)  # advances cu_seqlens inside the step
            else:
                self._decode_advance(batch)
                start_time = time.time_ns()
                out, _ = self.model.forward(batch.input_ids, batch.position_ids, batch.cu_seqlens, batch.cu_seqlens_q,
                                            batch.max_seqlen, None, kv, None, None)
                forward_time_ns = time.time_ns() - start_time
                generated, errs = self._process_decode(batch, out)
                batch.cu_seqlens.add_(batch.cu_seqlens_q)
            input_infos = None
        if first:
            batch.cu_seqlens.add_(batch.cu_seqlens_q)
        batch.max_seqlen += 1
        return generated, input_infos, errs, forward_time_ns

    def _decode_advance(self, batch) -> None:
        """device bookkeeping at the start of a decode step: position = tokens cached, slot from the block table,
        context += 1 (csrc/kv_alloc.cu)"""
        kv = batch.past_key_values
        _lib.check(_lib.load().b200_decode_advance(
            kv.block_table.data_ptr(), kv.block_table.stride(0), kv.context_lens.data_ptr(), batch.position_ids.data_ptr(),
            kv.slot_mapping.data_ptr(), None, None, len(batch), torch.cuda.current_stream().cuda_stream), "decode_advance")

    # --------------------------------------------------------------------------------------------- fused greedy decode
    def _can_fuse_greedy(self, batch) -> bool:
        """The token is chosen on the device inside the (CUDA-graph replayed) step and the ids chain device-to-device; the host
        reads back B ids (+ log-probabilities / ranks when asked) per step.  Plain greedy batches arg-max inside the C++ step
        (with the min_new_tokens EOS mask); sampling, penalties, top-k / top-p, logprobs and ranks go through the fused chooser
        kernel (csrc/chooser.cu).  Typical-p and top-n details stay on the op-by-op path."""
        # the answer depends only on the requests and their chooser: asked every step, computed once per batch composition
        # (prune installs a new requests list and chooser, concatenate a new batch)
        memo = getattr(batch, "_can_fuse_memo", None)
        if memo is not None and memo[0] is batch.requests and memo[1] is batch.next_token_chooser:
            return memo[2]
        answer = self._can_fuse_greedy_uncached(batch)
        batch._can_fuse_memo = (batch.requests, batch.next_token_chooser, answer)
        return answer

    def _can_fuse_greedy_uncached(self, batch) -> bool:
        # families without the C++ step runtime (flash GPT-NeoX) run op by op unless their Python step is switched on
        if not getattr(self.model, "fused_greedy_enabled", hasattr(self.model, "make_step")):
            return False
        chooser = batch.next_token_chooser
        if chooser.is_plain_greedy and not any(r.details.logprobs or r.details.ranks or r.details.top_n_toks for r in batch.requests):
            return True
        if not FUSED_CHOOSER or any(r.details.top_n_toks for r in batch.requests):
            return False
        tp = getattr(self.engine, "world_size", 1)
        V = self.model.lm_head.linear.weight.shape[0] * (tp if getattr(self.model.lm_head, "should_gather", False) else 1)
        return V <= 131072 and all(x >= 1.0 for x in chooser.typical_p) and self.dtype == torch.float16

    def _decode_fused_greedy(self, batch):
        kv, B = batch.past_key_values, len(batch)
        chooser = batch.next_token_chooser
        key = (batch.input_ids.data_ptr(), batch.position_ids.data_ptr(), kv.block_table.data_ptr(), kv.context_lens.data_ptr(),
               kv.slot_mapping.data_ptr(), batch.all_input_ids_tensor.data_ptr(), batch.cu_seqlens.data_ptr(), B,
               self.model.scratch.version, id(chooser))
        st = getattr(batch, "_fused", None)
        if st is None or st["key"] != key:
            V = self.model.lm_head.linear.weight.shape[0]
            st = dict(key=key, next_ids=torch.empty(B, dtype=torch.int64, device=self.device),
                      logits=torch.empty(B, V, dtype=torch.float16, device=self.device),
                      banned=torch.full((B,), -1, dtype=torch.int64, device=self.device), banned_host=[-1] * B,
                      steps=0, graph=None, max_s_cap=0)
            tp = getattr(self.engine, "world_size", 1)
            plain = chooser.is_plain_greedy and not any(r.details.logprobs or r.details.ranks for r in batch.requests)
            st["device_chooser"] = None if plain else chooser.device_chooser()
            st["want_logprobs"] = any(r.details.logprobs for r in batch.requests)
            st["want_ranks"] = any(r.details.ranks for r in batch.requests)
            in_step = plain and (tp == 1 or getattr(self.model, "greedy_ids_in_step", False))
            st["ids_in_step"] = in_step
            st["step"] = self.model.make_step(T=B, B=B, is_prefill=False, max_s=max(batch.total_lengths), input_ids=batch.input_ids,
                                              position_ids=batch.position_ids, kv=kv, logits=st["logits"],
                                              next_ids=st["next_ids"] if in_step else None)
            st["step"].banned_ids = st["banned"].data_ptr()
            if hasattr(st["step"], "banned"):  # Python step objects (NeoxStep) take the tensor itself
                st["step"].banned = st["banned"]
            if tp > 1 and not in_step and self.model.lm_head.should_gather:
                # vocab-sharded head: all-gather the [B, V/tp] logits (utils/layers.py:249-269), then one arg-max
                st["gathered"] = torch.empty(tp, B, V, dtype=torch.float16, device=self.device)
                st["full"] = torch.empty(B, tp * V, dtype=torch.float16, device=self.device)
            batch._fused = st
        dc = st["device_chooser"]
        if dc is None:
            # min_new_tokens EOS mask (utils/tokens.py:242-246) as a per-row banned id for the in-step arg-max
            eos = chooser.eos_token_id
            banned = [eos if c < m else -1 for c, m in zip(chooser.current_tokens, chooser.min_new_tokens)]
            for i, bnd in enumerate(banned):
                if bnd >= 0:
                    chooser.current_tokens[i] += 1
            if banned != st["banned_host"]:
                st["banned"].copy_(torch.tensor(banned, dtype=torch.int64), non_blocking=True)
                st["banned_host"] = banned
        else:
            dc.set_step(*chooser.step_masks())  # EOS mask / length-penalty factors of this step; draw counters advance on the device
        start_time = time.time_ns()
        self._run_fused_step(batch, st)
        forward_time_ns = time.time_ns() - start_time
        ids = st["next_ids"].tolist()  # the step's D2H read
        lps = dc.logprobs.tolist() if dc is not None and st["want_logprobs"] else None
        rks = dc.ranks.tolist() if dc is not None and st["want_ranks"] else None
        generated = []
        for i, (r, tok) in enumerate(zip(batch.requests, ids)):
            info = TokenInfo(request_id=r.id, token_id=tok)
            if lps is not None and r.details.logprobs:
                info.logprob = lps[i]
            if rks is not None and r.details.ranks:
                info.rank = rks[i]
            generated.append(info)
        batch.input_lengths = [n + 1 for n in batch.input_lengths]
        st["steps"] += 1
        return generated, [], forward_time_ns

    def _run_fused_step(self, batch, st, use_graph: bool = True) -> None:
        """decode_advance, the whole model step and the batch's device-side bookkeeping for the next step (position_ids,
        input_ids <- chosen ids, all_input_ids_tensor, cu_seqlens: flash_causal_lm.py:457-458, 533-535 of the reference);
        replayed as ONE CUDA graph from the third step of a stable batch on."""
        kv, B = batch.past_key_values, len(batch)
        lib = _lib.load()
        s = st["step"]
        if use_graph and st["graph"] is not None:
            st["graph"].replay()
            return

        def enqueue():
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(lib.b200_decode_advance(
                kv.block_table.data_ptr(), kv.block_table.stride(0), kv.context_lens.data_ptr(), batch.position_ids.data_ptr(),
                kv.slot_mapping.data_ptr(), None, None, B, stream), "decode_advance")
            self.model.run_step(s)
            if not st.get("ids_in_step", False):
                from .. import ops
                logits = st["logits"]
                if "gathered" in st:
                    torch.distributed.all_gather_into_tensor(st["gathered"], st["logits"], group=self.model.process_group)
                    st["full"].view(B, -1, st["logits"].shape[1]).copy_(st["gathered"].permute(1, 0, 2))
                    logits = st["full"]
                dc = st.get("device_chooser")
                if dc is not None:
                    # position_ids hold the index of each row's INPUT token here: the history is one longer
                    dc.launch(batch.all_input_ids_tensor, batch.position_ids, logits, st["want_logprobs"], st["want_ranks"],
                              out_ids=st["next_ids"], history_len_bias=1)
                else:
                    ops.argmax(logits, st["banned"], out=st["next_ids"])
            batch.position_ids.add_(1)
            batch.input_ids.copy_(st["next_ids"])
            batch.all_input_ids_tensor.scatter_(dim=1, index=batch.position_ids[:, None], src=st["next_ids"][:, None])
            batch.cu_seqlens.add_(batch.cu_seqlens_q)

        # Captured from the third step of a batch state on.  Measured on the 8B continuous-batching session (a state lives 2-3
        # steps on average there): capturing after 2 steps 11.2 k tokens/s, only after 8 stable steps 9.6 k, after 1 step 9.3 k
        # (same box, profiles/r2_graph_capture_policy.txt) - the eager step is host-bound enough that even short-lived graphs pay.
        if use_graph and st["steps"] >= 2 and USE_CUDA_GRAPHS:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                enqueue()
            st["graph"] = g
            g.replay()
            return
        enqueue()

    def _process_prefill(self, batch, out, all_rows: bool):
        generated: List[TokenInfo] = []
        input_infos: List[InputTokens] = []
        errs: List[GenerateError] = []
        batch.position_ids = batch.position_ids.new_tensor(batch.input_lengths)
        batch.input_ids = self._process_new_tokens(batch, out, generated, errs, input_infos, True, all_rows)
        batch.inputs_embeds = None
        batch.cu_seqlens_q = torch.arange(len(batch) + 1, device=self.device, dtype=torch.int32)
        return generated, input_infos, errs

    def _process_decode(self, batch, out):
        generated: List[TokenInfo] = []
        errs: List[GenerateError] = []
        batch.position_ids += 1
        batch.input_ids = self._process_new_tokens(batch, out, generated, errs, None, False, True)
        return generated, errs

    def _process_new_tokens(self, batch, out, generated, errs, input_infos, prefill: bool, all_rows: bool):
        """:506-588."""
        if prefill and all_rows:
            logits = out[(batch.cu_seqlens[1:] - 1).to(torch.int64), :]
        else:
            logits = out
        chooser = batch.next_token_chooser
        details = [r.details for r in batch.requests]
        if (FUSED_CHOOSER and not chooser.is_plain_greedy and chooser.device_eligible(logits)
                and not any(d.top_n_toks or (prefill and d.input_toks) for d in details)):
            # one fused kernel instead of the warper chain + log_softmax (csrc/chooser.cu); position_ids already point at the slot
            # of the token being chosen, i.e. they equal the history length
            want_lp, want_rk = any(d.logprobs for d in details), any(d.ranks for d in details)
            next_ids, lps, rks = chooser.choose_on_device(batch.all_input_ids_tensor, batch.position_ids, logits.contiguous(),
                                                          want_lp, want_rk)
            next_ids = next_ids.clone()
            batch.all_input_ids_tensor.scatter_(dim=1, index=batch.position_ids[:, None], src=next_ids[:, None])
            ids = next_ids.tolist()
            lps = lps.tolist() if lps is not None else None
            rks = rks.tolist() if rks is not None else None
            for i, (r, tok) in enumerate(zip(batch.requests, ids)):
                info = TokenInfo(request_id=r.id, token_id=tok)
                if lps is not None and r.details.logprobs:
                    info.logprob = lps[i]
                if rks is not None and r.details.ranks:
                    info.rank = rks[i]
                generated.append(info)
                batch.input_lengths[i] += 1
            return next_ids
        next_ids, scores, logprobs = chooser(input_ids=batch.all_input_ids_tensor[:, :batch.max_seqlen], scores=logits)
        batch.all_input_ids_tensor.scatter_(dim=1, index=batch.position_ids[:, None], src=next_ids[:, None])
        plain = not any(r.details.logprobs or r.details.ranks or r.details.top_n_toks or (prefill and r.details.input_toks)
                        for r in batch.requests)
        if plain:
            # one D2H copy for the whole batch instead of a .item() sync per request (tokens.py:394)
            for i, (r, tok) in enumerate(zip(batch.requests, next_ids.tolist())):
                generated.append(TokenInfo(request_id=r.id, token_id=tok))
                batch.input_lengths[i] += 1
            return next_ids
        cu = batch.cu_seqlens.tolist()
        for i, (r, input_length, next_token, sc, lp, all_ids) in enumerate(
                zip(batch.requests, batch.input_lengths, next_ids, scores, logprobs, batch.all_input_ids_tensor)):
            try:
                lp_view = lp.view(-1, lp.shape[-1]) if r.details.logprobs else None
                generated.append(get_token_info(r, sc.view(-1, sc.shape[-1]), next_token.view(-1), lp_view))
                if prefill and r.details.input_toks:
                    input_infos.append(get_input_tokens_info(r, all_ids[:input_length], out[cu[i]:cu[i] + input_length - 1, :]))
            except Exception as e:
                logging.exception(f"token decoding error for request #{r.id}")
                errs.append(GenerateError(request_id=r.id, message=f"Token decoding error: {str(e)}"))
            batch.input_lengths[i] += 1
        return next_ids
