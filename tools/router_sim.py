"""Stand-in for the Rust router's batching loop (measurement + test harness; the router itself is out of scope and would
drive the shard unchanged).

Drives a `TextGenerationService` exactly through its wire messages, in process, the way `router/src/batcher.rs:399-570`
does over gRPC: while requests are waiting and the running batch has room, an add-on `Prefill` of the new requests (the
`to_prune` field carries the completed ids of cached batches, server.py:105-122), then `NextToken` naming every cached batch
id with its `completed_ids` - the shard prunes and concatenates (server.py:183-231).  Requests arrive by a Poisson process
in measured time (the clock advances by the wall time of each RPC; idle gaps are skipped), prompt and output lengths are
uniform in the given ranges, seeded.

`run_session` returns per-request token ids plus timing / traffic totals for decode steps and prefills separately.
"""
from __future__ import annotations

import asyncio
import random
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional


@dataclass
class SimRequest:
    id: int
    arrival_s: float
    prompt: List[int]
    max_new: int
    tokens: List[int] = field(default_factory=list)
    done: bool = False


def make_requests(n: int, rate_per_s: float, prompt_range, new_range, vocab: int, seed: int = 0, filler: Optional[int] = None
                  ) -> List[SimRequest]:
    """Poisson arrivals (exponential gaps), prompt length U[prompt_range], output length U[new_range].  `filler`: every prompt
    token is this id (the reference's "test " * N recipe); otherwise random ids >= 4."""
    rng = random.Random(seed)
    t = 0.0
    out = []
    for i in range(n):
        t += rng.expovariate(rate_per_s) if rate_per_s > 0 else 0.0
        L = rng.randint(*prompt_range)
        prompt = [filler] * L if filler is not None else [rng.randrange(4, vocab) for _ in range(L)]
        out.append(SimRequest(id=i, arrival_s=t, prompt=prompt, max_new=rng.randint(*new_range)))
    return out


class _Ctx:
    async def abort(self, code, details):
        raise RuntimeError(f"gRPC abort {code}: {details}")


def run_session(service, pb, requests: List[SimRequest], max_batch_size: int, text_of: Callable[[List[int]], str],
                max_batch_tokens: Optional[int] = None, sync: Optional[Callable[[], None]] = None,
                on_decode_step: Optional[Callable[[int, int], None]] = None,
                agree: Optional[Callable[[float], float]] = None) -> Dict:
    """max_batch_tokens: admission budget in tokens (prompt + max_new summed over the running requests), the role of the
    router's weight limit (queue.rs:266-345).  on_decode_step(batch_size, sum of contexts) lets the caller count bytes.
    agree(seconds) -> seconds: with several shards every rank runs this loop itself (the real router broadcasts each RPC to all
    shards, sharded_client.rs:38-48); admission depends on the clock, so the ranks must advance it by the SAME amount - pass a
    max-over-ranks reduction."""
    ctx = _Ctx()
    loop = asyncio.new_event_loop()
    waiting = sorted(requests, key=lambda r: r.arrival_s)
    by_id = {r.id: r for r in requests}
    running: Dict[int, List[int]] = {}   # cached batch id -> request ids still running in it
    completed_since: Dict[int, List[int]] = {}  # batch id -> ids completed since the shard last heard
    now = 0.0
    next_batch_id = 0
    stats = dict(prefill_calls=0, prefill_tokens=0, prefill_s=0.0, decode_steps=0, decode_tokens=0, decode_s=0.0,
                 batch_size_sum=0, concat_steps=0, max_batch=0)

    def timed(coro):
        if sync:
            sync()
        t0 = time.perf_counter()
        res = loop.run_until_complete(coro)
        if sync:
            sync()
        dt = time.perf_counter() - t0
        return res, (agree(dt) if agree else dt)

    def n_running():
        return sum(len(v) for v in running.values())

    def tokens_booked():
        return sum(len(by_id[i].prompt) + by_id[i].max_new for v in running.values() for i in v)

    def take(result):
        for t in result.output_tokens:
            r = by_id[t.request_id]
            r.tokens.append(t.token_id)
            if len(r.tokens) >= r.max_new and not r.done:
                r.done = True
                for bid, ids in running.items():
                    if r.id in ids:
                        ids.remove(r.id)
                        completed_since.setdefault(bid, []).append(r.id)

    def cached_batches():
        out = []
        for bid in list(running.keys()):
            done = sorted(completed_since.pop(bid, []))
            if running[bid]:
                out.append(pb.CachedBatch(batch_id=bid, status=pb.RequestsStatus(completed_ids=done)))
            else:
                out.append(pb.CachedBatch(batch_id=bid))  # status absent: the whole batch finished (server.py:191-199)
                del running[bid]
        return out

    try:
        while waiting or running:
            if not running and waiting and waiting[0].arrival_s > now:
                now = waiting[0].arrival_s  # idle: skip to the next arrival
            # ---- add-on prefill of everything that has arrived and fits (batcher.rs:459-474, simplified: no waiting-token heuristic)
            new = []
            while waiting and waiting[0].arrival_s <= now and n_running() + len(new) < max_batch_size:
                r = waiting[0]
                need = len(r.prompt) + r.max_new
                if max_batch_tokens is not None and tokens_booked() + sum(len(x.prompt) + x.max_new for x in new) + need > max_batch_tokens:
                    break
                new.append(waiting.pop(0))
            if new:
                reqs = [pb.Request(id=r.id, inputs=text_of(r.prompt), input_length=len(r.prompt), truncate=False, max_output_length=r.max_new,
                                   parameters=pb.NextTokenChooserParameters(temperature=0.0, top_p=1.0, min_new_tokens=r.max_new))
                        for r in new]
                bid = next_batch_id
                next_batch_id += 1
                to_prune = [cb for cb in cached_batches()] if running else []
                res, dt = timed(service.Prefill(pb.PrefillRequest(batch=pb.Batch(id=bid, requests=reqs), to_prune=to_prune), ctx))
                now += dt
                stats["prefill_calls"] += 1
                stats["prefill_tokens"] += sum(len(r.prompt) for r in new)
                stats["prefill_s"] += dt
                running[bid] = [r.id for r in new]
                take(res.result)
            if not running:
                continue
            # ---- one decode step over every cached batch (concatenated by the shard)
            cbs = cached_batches()
            if not running:  # everything finished on its first token: tell the shard to drop the batches, no step
                loop.run_until_complete(service.NextToken(pb.NextTokenRequest(batches=cbs), ctx))
                continue
            bs = n_running()
            if on_decode_step:
                on_decode_step(bs, sum(len(by_id[i].prompt) + len(by_id[i].tokens) for v in running.values() for i in v))
            res, dt = timed(service.NextToken(pb.NextTokenRequest(batches=cbs), ctx))
            now += dt
            stats["decode_steps"] += 1
            stats["decode_tokens"] += bs
            stats["decode_s"] += dt
            stats["batch_size_sum"] += bs
            stats["max_batch"] = max(stats["max_batch"], bs)
            if len(cbs) > 1:
                stats["concat_steps"] += 1
            if res.HasField("result"):
                merged = res.result.batch_id
                if len(running) > 1:  # the shard concatenated: one cached batch remains, under the first id
                    ids = [i for v in running.values() for i in v]
                    pend = [i for v in completed_since.values() for i in v]
                    running.clear()
                    completed_since.clear()
                    running[merged] = ids
                    if pend:
                        completed_since[merged] = pend
                take(res.result)
        stats["sim_time_s"] = now
    finally:
        loop.close()
    return {"tokens": {r.id: r.tokens for r in requests}, "stats": stats}
