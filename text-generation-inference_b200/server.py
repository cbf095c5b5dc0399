"""Shard server: the router-facing `generate.v1.TextGenerationService` over `unix://{uds_path}-{rank}`.

Thin re-host of /root/reference/server/text_generation_server/server.py:58-249 (servicer: batch cache keyed by batch id,
prune -> concatenate -> `generate_token`, health batch id 2^64-1 never cached, OOM -> RESOURCE_EXHAUSTED) and :251-441
(`serve`: one process per GPU, rank 0 answers ServiceDiscovery with every shard's URL), plus `cache.py:8-33`.
Same message semantics, so the unchanged Rust router (`router/client/src/sharded_client.rs`) drives it.  Model loading
is `FlashCausalLM` only (the hot-path model class); `ModelInfo.batch_padding` is false (ragged batches, server.py:83) and the
memory-scaling coefficients are analytic bytes-per-token of the paged KV pool.
"""
from __future__ import annotations

import asyncio
import logging
import os
from typing import Dict, List, Optional

import torch

from . import pb as generate_pb2
from .models.types import Batch
from .utils.paged import OutOfBlocks

HEALTHCHECK_BATCH_ID = (1 << 64) - 1


class Cache:
    """cache.py:8-33"""

    def __init__(self):
        self.cache: Dict[int, Batch] = {}

    def pop(self, batch_id: int) -> Optional[Batch]:
        return self.cache.pop(batch_id, None)

    def set(self, entry: Optional[Batch]):
        if entry is not None:
            self.cache[entry.batch_id] = entry

    def delete(self, batch_id: int):
        del self.cache[batch_id]

    def clear(self):
        self.cache.clear()

    def __len__(self):
        return len(self.cache)

    def keys(self) -> list:
        return list(self.cache.keys())

    def compact(self):
        for batch in self.cache.values():
            batch.compact()


def memory_scaling_model(model) -> "generate_pb2.MemoryScalingModel":
    """Analytic stand-in for utils/memory_characterizer.py:42-143: with a pre-allocated paged KV pool the marginal
    cost of a token is its KV bytes; the limit is the pool size.  The router multiplies (router/src/batch_types.rs:68-83)."""
    mgr = model.kv_cache_manager
    per_token = mgr.block_bytes() / mgr.block_size
    # The router books a request at input + max_new tokens (FlashBatch.batch_max_weight, router/src/batch_types.rs:68-83), which is
    # what generate_token(first=True) reserves - in whole 16-token blocks per sequence.  The router's model has no per-sequence term,
    # so the up-to-15 slots a sequence's last block wastes come off the limit instead, for as many sequences as a batch may hold
    # (the launcher's --max-batch-size is not visible to the shard: 256, overridable with MAX_BATCH_SIZE).
    max_batch = int(os.getenv("MAX_BATCH_SIZE", "256"))
    usable_tokens = mgr.total_num_gpu_blocks * mgr.block_size - (mgr.block_size - 1) * max_batch
    if usable_tokens <= 0:  # a pool too small for the margin: advertise what there is
        usable_tokens = mgr.total_num_gpu_blocks * mgr.block_size
    return generate_pb2.MemoryScalingModel(
        prefill_linear_coef0=per_token, prefill_quadratic_coef0=0.0, prefill_quadratic_coef1=0.0,
        nexttoken_linear_coef0=0.0, nexttoken_linear_coef1=per_token,
        weight_limit=int(usable_tokens * per_token))


class TextGenerationService:
    def __init__(self, model, cache: Cache, server_urls: List[str]):
        self.cache = cache
        self.model = model
        self.server_urls = server_urls

    async def _guard(self, coro, context):
        try:
            return await coro
        except (torch.cuda.OutOfMemoryError, OutOfBlocks) as e:  # server.py:48-51
            import grpc
            logging.exception("GPU memory exhausted")
            await context.abort(grpc.StatusCode.RESOURCE_EXHAUSTED, str(e))

    async def ServiceDiscovery(self, request, context):
        return generate_pb2.ServiceDiscoveryResponse(urls=self.server_urls)

    async def ClearCache(self, request, context):
        for batch in list(self.cache.cache.values()):
            self._free_paged_sequences(batch, None)
        self.cache.clear()
        return generate_pb2.ClearCacheResponse()

    async def ModelInfo(self, request, context):
        tok = self.model.tokenizer
        return generate_pb2.ModelInfoResponse(
            model_type=generate_pb2.ModelInfoResponse.ModelType.Value("CAUSAL_LM"),
            eos_token=getattr(tok, "model_eos_token_id", tok.eos_token_id),
            batch_padding=False,
            memory_scaling_model=memory_scaling_model(self.model))

    async def Health(self, request, context):
        torch.zeros((2, 2)).cuda()
        return generate_pb2.HealthResponse()

    async def PrefixLookup(self, request, context):
        import grpc
        await context.abort(grpc.StatusCode.NOT_FOUND, f'prefix id "{request.prefix_id}" not found')

    async def PruneBatch(self, request, context):
        import grpc
        await context.abort(grpc.StatusCode.UNIMPLEMENTED, "PruneBatch has no servicer in the reference either")

    async def Prefill(self, request, context):
        return await self._guard(self._prefill(request), context)

    async def NextToken(self, request, context):
        return await self._guard(self._next_token(request), context)

    async def _prefill(self, request):
        with self.model.context_manager():
            for cbatch in request.to_prune:
                batch_to_prune = self.cache.pop(cbatch.batch_id)
                if batch_to_prune is None:
                    raise ValueError(f"Batch ID {cbatch.batch_id} not found in cache.")
                completed_ids = list(cbatch.status.completed_ids) if cbatch.HasField("status") else None
                self._free_paged_sequences(batch_to_prune, completed_ids)
                if completed_ids is not None:
                    self.cache.set(self.model.batch_type.prune(batch_to_prune, completed_ids))
                del batch_to_prune
            is_healthcheck = request.batch.id == HEALTHCHECK_BATCH_ID
            if not is_healthcheck:
                self.cache.compact()
            input_token_info = None
            forward_time_ns = 0
            batch, errors = self.model.batch_type.from_pb(
                request.batch, tokenizer=self.model.tokenizer, dtype=self.model.dtype, device=self.model.device,
                embeddings_lookup=self.model.word_embeddings, prefix_cache=self.model.prefix_cache,
                use_position_ids=self.model.use_position_ids)
            batch_id = 0
            if batch is not None:
                for_concat = len(self.cache) > 0
                try:
                    output_tokens, input_token_info, decode_errors, forward_time_ns = self.model.generate_token(
                        batch, first=True, for_concat=for_concat)
                except BaseException:
                    self._free_paged_sequences(batch, None)
                    raise
                if not is_healthcheck:
                    self.cache.set(batch)
                else:
                    self._free_paged_sequences(batch, None)
                batch_id = batch.get_id()
                errors = (errors + decode_errors) if errors else decode_errors
            else:
                output_tokens = []
            return generate_pb2.PrefillResponse(
                result=generate_pb2.GenerateResult(
                    output_tokens=[t.to_pb() for t in output_tokens], errors=[e.to_pb() for e in errors] if errors else None,
                    batch_id=batch_id, forward_time_ns=forward_time_ns),
                input_tokens=[it.to_pb() for it in input_token_info] if input_token_info is not None else None)

    async def _next_token(self, request):
        if len(request.batches) == 0:
            raise ValueError("Must provide at least one batch")
        with self.model.context_manager():
            batches = []
            for cbatch in request.batches:
                batch = self.cache.pop(cbatch.batch_id)
                completed_ids = list(cbatch.status.completed_ids) if cbatch.HasField("status") else None
                self._free_paged_sequences(batch, completed_ids)
                if completed_ids is not None:
                    if batch is None:
                        raise ValueError(f"Batch ID {cbatch.batch_id} not found in cache.")
                    batch = self.model.batch_type.prune(batch, completed_ids)
                    if batch is not None:
                        batches.append(batch)
            if len(self.cache) > 0:
                print(f"WARN: Clearing additional batches found in cache: {self.cache.keys()}")
                for b in list(self.cache.cache.values()):
                    self._free_paged_sequences(b, None)
                self.cache.clear()
            if len(batches) == 0:
                return generate_pb2.NextTokenResponse()
            batch = batches[0] if len(batches) == 1 else self.model.batch_type.concatenate(batches)
            del batches
            try:
                output_tokens, _, errors, forward_time_ns = self.model.generate_token(batch)
            except BaseException:
                self._free_paged_sequences(batch, None)
                raise
            self.cache.set(batch)
            return generate_pb2.NextTokenResponse(result=generate_pb2.GenerateResult(
                output_tokens=[t.to_pb() for t in output_tokens], errors=[e.to_pb() for e in errors] if errors else None,
                batch_id=batch.get_id(), forward_time_ns=forward_time_ns))

    def _free_paged_sequences(self, batch, completed_ids: Optional[List[int]]):
        """server.py:233-249; completed_ids None = free the whole batch"""
        if batch is None or not hasattr(self.model, "kv_cache_manager"):
            return
        if completed_ids is None:
            ids = list(batch.sequence_ids)
        elif completed_ids:
            done = set(completed_ids)
            ids = [batch.sequence_ids[i] for i, r in enumerate(batch.requests) if r.id in done] if batch.sequence_ids else []
        else:
            return
        if ids:
            self.model.kv_cache_manager.free_sequences(ids, recursive=True)
            if completed_ids is None and batch.past_key_values is not None:
                batch.past_key_values.sequence_ids = []


def serve(model, uds_path: str = "/tmp/text-generation", sharded: bool = False):
    """server.py:251-441 for an already-built model: serve on unix://{uds_path}-{rank}; blocks until cancelled."""
    import grpc

    rank = int(os.getenv("RANK", "0"))
    world = int(os.getenv("WORLD_SIZE", "1"))
    urls = [f"unix://{uds_path}-{r}" for r in range(world)] if sharded or world > 1 else [f"unix://{uds_path}-0"]
    local_url = urls[rank]

    async def _serve():
        server = grpc.aio.server()
        generate_pb2.add_TextGenerationServiceServicer_to_server(TextGenerationService(model, Cache(), urls), server)
        server.add_insecure_port(local_url)
        await server.start()
        print(f"Server started at {local_url}")
        try:
            await server.wait_for_termination()
        finally:
            await server.stop(0)

    asyncio.run(_serve())
