"""CPU test of tools/router_sim.py (the Rust router's batching loop restated for bench.py's continuous-batching workloads and
the GPU session test) against a fake shard that enforces the wire protocol of server.py:105-231: every cached batch id is named
in every NextToken, completed ids are ascending and belong to the batch, a batch without `status` is dropped, an add-on Prefill
may carry `to_prune`, several ids in one NextToken are merged under the first."""
import asyncio

from tgis_b200 import pb
from tools import router_sim


class FakeShard:
    def __init__(self):
        self.cache = {}   # batch id -> list of request ids (ascending)
        self.count = {}   # request id -> tokens produced
        self.max_cached = 0

    def _apply(self, cb):
        assert cb.batch_id in self.cache, f"unknown batch {cb.batch_id}"
        ids = self.cache.pop(cb.batch_id)
        if not cb.HasField("status"):
            return None
        done = list(cb.status.completed_ids)
        assert done == sorted(done) and set(done) <= set(ids), (done, ids)
        keep = [i for i in ids if i not in set(done)]
        return keep or None

    def _tokens(self, ids):
        out = []
        for i in ids:
            self.count[i] = self.count.get(i, 0) + 1
            out.append(pb.Token(request_id=i, token_id=1000 * i + self.count[i]))
        return out

    async def Prefill(self, request, context):
        for cb in request.to_prune:
            keep = self._apply(cb)
            if keep:
                self.cache[cb.batch_id] = keep
        ids = [r.id for r in request.batch.requests]
        assert ids == sorted(ids) and request.batch.id not in self.cache
        for r in request.batch.requests:
            assert r.input_length == len(r.inputs.split()) and r.max_output_length >= 1
        self.cache[request.batch.id] = ids
        self.max_cached = max(self.max_cached, len(self.cache))
        return pb.PrefillResponse(result=pb.GenerateResult(output_tokens=self._tokens(ids), batch_id=request.batch.id))

    async def NextToken(self, request, context):
        named = [cb.batch_id for cb in request.batches]
        assert sorted(named) == sorted(self.cache), f"router named {named}, cached {sorted(self.cache)}"
        batches = []
        for cb in request.batches:
            keep = self._apply(cb)
            if keep:
                batches.append((cb.batch_id, keep))
        if not batches:
            return pb.NextTokenResponse()
        bid = batches[0][0]
        ids = [i for _, k in batches for i in k]
        assert ids == sorted(ids)
        self.cache[bid] = ids
        return pb.NextTokenResponse(result=pb.GenerateResult(output_tokens=self._tokens(ids), batch_id=bid))


def test_router_sim_speaks_the_shard_protocol():
    shard = FakeShard()
    reqs = router_sim.make_requests(60, rate_per_s=5000.0, prompt_range=(1, 9), new_range=(1, 7), vocab=50, seed=3)
    out = router_sim.run_session(shard, pb, reqs, max_batch_size=5, text_of=lambda p: " ".join("x" for _ in p), max_batch_tokens=70)
    st = out["stats"]
    for r in reqs:
        assert out["tokens"][r.id] == [1000 * r.id + k for k in range(1, r.max_new + 1)], r.id
    assert not shard.cache, "every batch must have been dropped at the end"
    assert st["max_batch"] <= 5 and st["decode_tokens"] == sum(r.max_new - 1 for r in reqs)
    assert st["prefill_tokens"] == sum(len(r.prompt) for r in reqs) and st["concat_steps"] > 0 and shard.max_cached == 2


def test_router_sim_respects_arrival_times():
    shard = FakeShard()
    reqs = router_sim.make_requests(5, rate_per_s=1.0, prompt_range=(2, 2), new_range=(3, 3), vocab=50, seed=0)
    out = router_sim.run_session(shard, pb, reqs, max_batch_size=8, text_of=lambda p: "x x")
    # arrivals are ~1 s apart and a fake RPC takes microseconds: no two requests ever run together
    assert out["stats"]["max_batch"] == 1 and out["stats"]["prefill_calls"] == 5
    assert out["stats"]["sim_time_s"] >= reqs[-1].arrival_s


def test_router_sim_ranks_agree_on_the_clock():
    """With several shards every rank runs the loop itself; `agree` replaces each RPC's measured time by one all ranks share, so
    two ranks whose own clocks differ wildly still admit the same requests at the same steps."""
    def session(own_speed):
        shard = FakeShard()
        reqs = router_sim.make_requests(40, rate_per_s=20.0, prompt_range=(1, 9), new_range=(2, 9), vocab=50, seed=11)
        calls = []
        ticks = iter(range(10 ** 6))

        def agree(dt):
            assert dt >= 0.0
            calls.append(dt * own_speed)       # what this rank measured does not matter ...
            return 0.013 + 0.001 * (next(ticks) % 7)  # ... the agreed (max over ranks) value drives the session clock

        out = router_sim.run_session(shard, pb, reqs, max_batch_size=6, text_of=lambda p: " ".join("x" for _ in p), agree=agree)
        return out, len(calls)

    a, n_a = session(1.0)
    b, n_b = session(1000.0)
    assert n_a == n_b == a["stats"]["prefill_calls"] + a["stats"]["decode_steps"]
    assert a["tokens"] == b["tokens"]
    for k in ("prefill_calls", "decode_steps", "decode_tokens", "batch_size_sum", "concat_steps", "max_batch", "sim_time_s"):
        assert a["stats"][k] == b["stats"][k], k
