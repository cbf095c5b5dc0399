"""CPU test of the shard servicer's batch-cache semantics (reference server.py:58-249) with a stub model: what the Rust
router relies on - Prefill caches under the batch id, the health-check batch (id 2^64-1) is never cached,
NextToken prunes by completed ids, concatenates several cached batches, returns an empty response when everything
finished, frees KV sequences of finished requests, and rejects unknown batch ids."""
import asyncio
import contextlib
import types

import pytest
import torch

import tgis_b200  # noqa: F401
from tgis_b200 import pb
from tgis_b200.server import HEALTHCHECK_BATCH_ID, Cache, TextGenerationService


class _Tok:
    def __init__(self, request_id, token_id):
        self.request_id, self.token_id = request_id, token_id

    def to_pb(self):
        return pb.Token(request_id=self.request_id, token_id=self.token_id)


class _Batch:
    def __init__(self, batch_id, request_ids, seq_ids):
        self.batch_id, self.requests = batch_id, [types.SimpleNamespace(id=i) for i in request_ids]
        self.sequence_ids, self.past_key_values = list(seq_ids), types.SimpleNamespace(sequence_ids=list(seq_ids))

    def get_id(self):
        return self.batch_id

    def __len__(self):
        return len(self.requests)

    def compact(self):
        pass

    _next_seq = 100

    @classmethod
    def from_pb(cls, batch_pb, tokenizer, dtype, device, embeddings_lookup, prefix_cache, use_position_ids):
        ids = [r.id for r in batch_pb.requests]
        seqs = list(range(cls._next_seq, cls._next_seq + len(ids)))
        cls._next_seq += len(ids)
        return cls(batch_pb.id, ids, seqs), []

    @classmethod
    def prune(cls, batch, completed_ids):
        keep = [(r.id, s) for r, s in zip(batch.requests, batch.sequence_ids) if r.id not in set(completed_ids)]
        if not keep:
            return None
        return cls(batch.batch_id, [k[0] for k in keep], [k[1] for k in keep])

    @classmethod
    def concatenate(cls, batches):
        cls.concatenated = [b.batch_id for b in batches]
        return cls(batches[0].batch_id, [r.id for b in batches for r in b.requests], [s for b in batches for s in b.sequence_ids])


class _Model:
    batch_type = _Batch
    dtype, device, word_embeddings, prefix_cache, use_position_ids = torch.float16, torch.device("cpu"), None, None, True
    tokenizer = types.SimpleNamespace(eos_token_id=2)

    def __init__(self):
        self.freed = []
        self.kv_cache_manager = types.SimpleNamespace(free_sequences=lambda ids, recursive=True: self.freed.extend(ids))
        self.calls = []

    def context_manager(self):
        return contextlib.nullcontext()

    def generate_token(self, batch, first=False, for_concat=False):
        self.calls.append((batch.batch_id, first, for_concat, [r.id for r in batch.requests]))
        return [_Tok(r.id, 7) for r in batch.requests], ([] if first else None), [], 123


def _pb_batch(batch_id, request_ids):
    return pb.Batch(id=batch_id, requests=[pb.Request(id=i, inputs="test", input_length=1, max_output_length=4) for i in request_ids])


def _cached(batch_id, completed=None):
    if completed is None:
        return pb.CachedBatch(batch_id=batch_id)
    return pb.CachedBatch(batch_id=batch_id, status=pb.RequestsStatus(completed_ids=completed))


def test_servicer_batch_cache_semantics():
    model = _Model()
    svc = TextGenerationService(model, Cache(), ["unix:///tmp/x-0", "unix:///tmp/x-1"])
    run = asyncio.run
    assert list(run(svc.ServiceDiscovery(pb.ServiceDiscoveryRequest(), None)).urls) == ["unix:///tmp/x-0", "unix:///tmp/x-1"]

    r = run(svc._prefill(pb.PrefillRequest(batch=_pb_batch(1, [10, 11, 12]))))
    assert r.result.batch_id == 1 and [t.request_id for t in r.result.output_tokens] == [10, 11, 12] and r.result.forward_time_ns == 123
    assert svc.cache.keys() == [1] and model.calls[-1] == (1, True, False, [10, 11, 12])

    # health check: served, never cached, its KV released (server.py:37,124,155-158)
    n_freed = len(model.freed)
    run(svc._prefill(pb.PrefillRequest(batch=_pb_batch(HEALTHCHECK_BATCH_ID, [99]))))
    assert svc.cache.keys() == [1] and len(model.freed) == n_freed + 1

    # decode with one finished request: pruned, its sequence freed
    seq_of_11 = svc.cache.cache[1].sequence_ids[1]
    r = run(svc._next_token(pb.NextTokenRequest(batches=[_cached(1, [11])])))
    assert [t.request_id for t in r.result.output_tokens] == [10, 12] and seq_of_11 in model.freed

    # add-on prefill while batch 1 is cached (for_concat), then NextToken with both ids -> concatenate
    run(svc._prefill(pb.PrefillRequest(batch=_pb_batch(2, [20]))))
    assert model.calls[-1][:3] == (2, True, True) and sorted(svc.cache.keys()) == [1, 2]
    r = run(svc._next_token(pb.NextTokenRequest(batches=[_cached(1, []), _cached(2, [])])))
    assert _Batch.concatenated == [1, 2] and [t.request_id for t in r.result.output_tokens] == [10, 12, 20]
    assert svc.cache.keys() == [1]

    # a batch named without status is finished as a whole (server.py:191-199); nothing left -> empty response
    r = run(svc._next_token(pb.NextTokenRequest(batches=[_cached(1)])))
    assert not r.HasField("result") and len(svc.cache) == 0

    with pytest.raises(ValueError):
        run(svc._next_token(pb.NextTokenRequest(batches=[_cached(77, [1])])))
    with pytest.raises(ValueError):
        run(svc._next_token(pb.NextTokenRequest(batches=[])))
    with pytest.raises(ValueError):
        run(svc._prefill(pb.PrefillRequest(batch=_pb_batch(3, [1]), to_prune=[_cached(55, [1])])))


def test_get_model_rejects_what_is_out_of_scope(tmp_path):
    """models/__init__.py: get_model - only local directories of flash decoder families; no CPU form of the model itself"""
    import json
    import torch
    import tgis_b200  # noqa: F401
    from tgis_b200.models import get_model
    with pytest.raises(ValueError):
        get_model("no/such/dir", None, "tgis_native", "float16", None, 2048)
    t5 = tmp_path / "t5"
    t5.mkdir()
    (t5 / "config.json").write_text(json.dumps({"model_type": "t5"}))
    with pytest.raises(NotImplementedError):
        get_model(str(t5), None, "hf_transformers", "float16", None, 2048)
    llama = tmp_path / "llama"
    llama.mkdir()
    (llama / "config.json").write_text(json.dumps({"model_type": "llama", "hidden_size": 64, "num_attention_heads": 4,
                                                   "num_hidden_layers": 1, "intermediate_size": 128, "vocab_size": 100}))
    with pytest.raises(ValueError):
        get_model(str(llama), None, "tgis_native", "float16", "bitsandbytes", 2048)
    with pytest.raises(ValueError):
        get_model(str(llama), None, "tgis_native", "float17", None, 2048)
    if not torch.cuda.is_available():
        with pytest.raises(NotImplementedError):  # "FlashCausalLM is only available on GPU": no CPU fallback
            get_model(str(llama), None, "tgis_native", "float16", None, 2048)
