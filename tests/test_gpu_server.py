"""GPU test of the wire surface: a shard serving generate.v1.TextGenerationService on a unix socket is driven the way the
Rust router drives it (ServiceDiscovery, ClearCache, ModelInfo, Prefill, NextToken with completed ids, add-on Prefill
followed by a two-batch NextToken = continuous batching, health batch id 2^64-1 never cached)."""
import asyncio
import os

import pytest
import torch

from oracle import llama as oll
from tests.test_gpu_generate import EOS, _oracle_tokens, _pb_batch, _prompts, _setup

pytestmark = pytest.mark.gpu


def test_router_style_session_over_uds(tmp_path):
    import grpc
    from tgis_b200 import pb
    from tgis_b200.server import HEALTHCHECK_BATCH_ID, Cache, TextGenerationService

    model, oracle, tok = _setup(tmp_path, None)
    mgr = model.kv_cache_manager
    n_new = 6
    pa, pbs = _prompts(11, [9, 3], 512), _prompts(12, [14], 512)
    ref, n_ex = {}, {}
    for i, p in enumerate(pa + pbs):
        t, ne = _oracle_tokens(oracle, [p], n_new)
        ref[i], n_ex[i] = t[0].tolist(), ne[0]
    url = f"unix://{tmp_path}/shard-0"
    got = {0: [], 1: [], 2: []}

    async def session():
        server = grpc.aio.server()
        pb.add_TextGenerationServiceServicer_to_server(TextGenerationService(model, Cache(), [url]), server)
        server.add_insecure_port(url)
        await server.start()
        async with grpc.aio.insecure_channel(url) as ch:
            stub = pb.TextGenerationServiceStub(ch)
            assert list((await stub.ServiceDiscovery(pb.ServiceDiscoveryRequest())).urls) == [url]
            await stub.ClearCache(pb.ClearCacheRequest())
            info = await stub.ModelInfo(pb.ModelInfoRequest())
            assert info.batch_padding is False and info.eos_token == EOS and info.memory_scaling_model.weight_limit > 0
            await stub.Health(pb.HealthRequest())
            # health-check prefill must not be cached and must give its blocks back (health.rs:43-83, server.py:155-158)
            hb = _pb_batch(HEALTHCHECK_BATCH_ID, [[5]], 1)
            await stub.Prefill(pb.PrefillRequest(batch=hb))
            assert mgr.free_blocks == mgr.total_num_gpu_blocks

            def take(res):
                for t in res.output_tokens:
                    got[t.request_id].append(t.token_id)

            r = await stub.Prefill(pb.PrefillRequest(batch=_pb_batch(0, pa, n_new, first_id=0)))
            assert r.result.batch_id == 0 and r.result.forward_time_ns > 0 and not r.input_tokens
            take(r.result)
            for _ in range(2):
                r = await stub.NextToken(pb.NextTokenRequest(batches=[pb.CachedBatch(batch_id=0, status=pb.RequestsStatus())]))
                take(r.result)
            # add-on batch, then NextToken carries both ids -> concatenated server-side
            r = await stub.Prefill(pb.PrefillRequest(batch=_pb_batch(1, pbs, n_new, first_id=2)))
            take(r.result)
            r = await stub.NextToken(pb.NextTokenRequest(batches=[pb.CachedBatch(batch_id=0, status=pb.RequestsStatus()),
                                                                  pb.CachedBatch(batch_id=1, status=pb.RequestsStatus())]))
            assert r.result.batch_id == 0 and len(r.result.output_tokens) == 3
            take(r.result)
            # request 1 completes; the merged batch keeps id 0
            r = await stub.NextToken(pb.NextTokenRequest(batches=[pb.CachedBatch(batch_id=0, status=pb.RequestsStatus(completed_ids=[1]))]))
            assert len(r.result.output_tokens) == 2
            take(r.result)
            # status absent = whole batch finished (server.py:191-199) -> empty response, every block returned
            r = await stub.NextToken(pb.NextTokenRequest(batches=[pb.CachedBatch(batch_id=0)]))
            assert not r.HasField("result")
            assert mgr.free_blocks == mgr.total_num_gpu_blocks
            # unknown batch id with a status is an error surfaced to the router
            with pytest.raises(grpc.aio.AioRpcError):
                await stub.NextToken(pb.NextTokenRequest(batches=[pb.CachedBatch(batch_id=77, status=pb.RequestsStatus())]))
        await server.stop(0)

    asyncio.run(session())
    for i in (0, 1, 2):
        n = min(len(got[i]), n_ex[i])
        assert n >= 2 and got[i][:n] == ref[i][:n], f"request {i}: {got[i]} vs oracle {ref[i]}"


@pytest.mark.parametrize("quantize", [None, "gptq"])
def test_continuous_batching_session_matches_oracle(tmp_path, quantize):
    """BASELINE config[4] in miniature: Poisson arrivals, ragged prompts and output lengths, at most 4 running requests, driven
    through Prefill (with to_prune) / NextToken (several cached batch ids -> concatenate, completed ids -> prune) by the router
    stand-in that bench.py's `*-mixed` workloads use.  Every request must receive exactly the tokens the oracle generates for it
    alone (outside the tie band): batching, pruning and concatenation change no result; every block returns to the pool."""
    from tgis_b200 import pb
    from tgis_b200.server import Cache, TextGenerationService
    from tests.test_gpu_generate import _text
    from tools import router_sim

    model, oracle, tok = _setup(tmp_path, quantize)
    mgr = model.kv_cache_manager
    service = TextGenerationService(model, Cache(), ["unix:///dev/null"])
    reqs = router_sim.make_requests(14, rate_per_s=400.0, prompt_range=(3, 40), new_range=(2, 10), vocab=512, seed=1)
    with torch.inference_mode():
        out = router_sim.run_session(service, pb, reqs, max_batch_size=4, text_of=_text, sync=torch.cuda.synchronize)
    st = out["stats"]
    assert st["prefill_calls"] >= 4 and st["concat_steps"] >= 2 and st["max_batch"] == 4, st
    assert mgr.free_blocks == mgr.total_num_gpu_blocks
    checked = 0
    for r in reqs:
        got = out["tokens"][r.id]
        assert len(got) == r.max_new
        ref, n_ex = _oracle_tokens(oracle, [r.prompt], r.max_new)
        n = min(len(got), n_ex[0])
        assert got[:n] == ref[0].tolist()[:n], f"request {r.id}: {got} vs oracle {ref[0].tolist()}"
        checked += n
    assert checked >= 40
