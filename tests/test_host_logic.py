"""CPU tests of the host-side runtime around the hot path: batch construction from the wire message, the paged KV
manager, prune / concatenate bookkeeping (block-table edits only), `get_indices_to_keep`, and the world-size-2
tensor-parallel host logic over gloo (shard loading + collective placement)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import llama as oll


@pytest.fixture(scope="module")
def mods():
    import __graft_entry__ as ge
    ge._load_build_module().build()
    import tgis_b200  # noqa: F401
    from tgis_b200 import pb
    from tgis_b200.models.flash_causal_lm import FlashCausalLMBatch
    from tgis_b200.models.model import Model
    from tgis_b200.utils.paged import OutOfBlocks, PagedKVCacheManager, PagedKVState
    from tgis_b200.utils.synthetic import make_tokenizer
    return dict(pb=pb, Batch=FlashCausalLMBatch, Model=Model, Mgr=PagedKVCacheManager, State=PagedKVState,
                OutOfBlocks=OutOfBlocks, tok=make_tokenizer(64))


def _req(pb, i, text, n_in, n_out, truncate=False, **params):
    return pb.Request(id=i, inputs=text, input_length=n_in, truncate=truncate, max_output_length=n_out,
                      parameters=pb.NextTokenChooserParameters(temperature=0.0, top_p=1.0, **params))


def test_from_pb_builds_the_ragged_batch(mods):
    pb, Batch, tok = mods["pb"], mods["Batch"], mods["tok"]
    msg = pb.Batch(id=11, requests=[_req(pb, 0, "test <tok5> test", 3, 4), _req(pb, 1, "test " * 100, 5, 2, truncate=True),
                                    _req(pb, 2, "<tok9>", 1, 6)])
    b, errs = Batch.from_pb(msg, tok, torch.float16, torch.device("cpu"), None, None, True)
    assert not errs and len(b) == 3 and b.get_id() == 11
    assert b.input_ids.tolist() == [3, 5, 3, 3, 3, 3, 3, 3, 9]
    assert b.position_ids.tolist() == [0, 1, 2, 0, 1, 2, 3, 4, 0]
    assert b.cu_seqlens.tolist() == [0, 3, 8, 9] and b.cu_seqlens.dtype == torch.int32
    assert b.input_lengths == [3, 5, 1] and b.total_lengths == [7, 7, 7] and b.max_seqlen == 5
    assert b.all_input_ids_tensor.shape == (3, 7)
    assert b.all_input_ids_tensor[1].tolist() == [3, 3, 3, 3, 3, 0, 0]  # padded with pad_token_id
    assert b.past_key_values is None and b.sequence_ids == []
    assert b.next_token_chooser.is_plain_greedy


def test_from_pb_reports_prefix_requests_as_errors(mods):
    pb, Batch, tok = mods["pb"], mods["Batch"], mods["tok"]
    r = _req(pb, 4, "test", 1, 2)
    r.prefix_id = "some-prefix"
    b, errs = Batch.from_pb(pb.Batch(id=1, requests=[r]), tok, torch.float16, torch.device("cpu"), None, None, True)
    assert b is None and len(errs) == 1 and errs[0].request_id == 4


def test_from_pb_prepends_prompt_prefixes_as_embeddings(mods):
    """flash_causal_lm.py:97-107, :147-168: a request with a prefix_id is lengthened by the prefix, its leading positions hold pad
    ids in all_input_ids_tensor, and the whole batch goes in as embeddings with the prefix rows filled in; a failed lookup
    only drops that request."""
    pb, Batch, tok = mods["pb"], mods["Batch"], mods["tok"]
    H = 8
    table = torch.arange(16 * H, dtype=torch.float32).view(16, H)
    lookup = lambda ids: table[ids].clone()  # noqa: E731
    prefix = -torch.ones(2, H)

    class Store:
        def get(self, prefix_id):
            if prefix_id != "tuned":
                raise KeyError(prefix_id)
            return prefix

    plain, tuned, missing = _req(pb, 1, "test test", 2, 3), _req(pb, 2, "test test test", 3, 2), _req(pb, 3, "test", 1, 2)
    tuned.prefix_id, missing.prefix_id = "tuned", "nope"
    b, errs = Batch.from_pb(pb.Batch(id=5, requests=[plain, tuned, missing]), tok, torch.float32, torch.device("cpu"), lookup, Store(), True)
    assert [e.request_id for e in errs] == [3] and len(b) == 2
    assert b.input_ids is None and b.inputs_embeds.shape == (2 + 5, H)
    assert b.input_lengths == [2, 5] and b.total_lengths == [5, 7] and b.max_seqlen == 5
    assert b.cu_seqlens.tolist() == [0, 2, 7] and b.position_ids.tolist() == [0, 1, 0, 1, 2, 3, 4]
    assert b.all_input_ids_tensor[1].tolist() == [0, 0, 3, 3, 3, 0, 0]            # pad ids where the prefix sits
    assert torch.equal(b.inputs_embeds[2:4], prefix) and torch.equal(b.inputs_embeds[4:7], table[[3, 3, 3]])
    assert torch.equal(b.inputs_embeds[0:2], table[[3, 3]])
    with pytest.raises(ValueError):
        Batch.from_pb(pb.Batch(id=6, requests=[tuned]), tok, torch.float32, torch.device("cpu"), None, Store(), True)


def test_get_indices_to_keep_matches_reference_semantics(mods):
    Model, pb = mods["Model"], mods["pb"]
    reqs = [pb.Request(id=i) for i in (2, 5, 7, 9, 12)]
    assert Model.get_indices_to_keep(reqs, [5, 9]) == [0, 2, 4]
    assert Model.get_indices_to_keep(reqs, [2, 5, 7, 9, 12]) == []
    assert Model.get_indices_to_keep(reqs, [1, 3, 12]) == [0, 1, 2, 3]  # ids not in the batch are skipped


def test_paged_kv_manager_allocation_and_exhaustion(mods):
    Mgr, OutOfBlocks = mods["Mgr"], mods["OutOfBlocks"]
    m = Mgr(num_layers=2, num_heads=4, emb_dim=256, kv_heads=2, device="cpu", total_num_gpu_blocks=10)
    assert m.pool.shape == (2, 2, 10, 2, 16, 64) and m.free_blocks == 10
    assert m.block_bytes() == 2 * 2 * 2 * 16 * 64 * 2
    s = m.allocate_tokens([17, 1], reserve_tokens=[15, 15])      # 32 -> 2 blocks, 16 -> 1 block
    assert [len(m.sequence_blocks(i)) for i in s] == [2, 1] and m.free_blocks == 7
    assert m.slot_mapping_for(s, [0, 0], [17, 1]).tolist() == [b * 16 + o for b, o in
                                                              [(m.sequence_blocks(s[0])[i // 16], i % 16) for i in range(17)]] + \
        [m.sequence_blocks(s[1])[0] * 16]
    bt = m.block_table_tensor(s)
    assert bt.shape == (2, 2) and bt.dtype == torch.int32 and bt[1, 1].item() == 0
    with pytest.raises(OutOfBlocks):
        m.allocate_tokens([16 * 8])                               # all or nothing
    assert m.free_blocks == 7
    for _ in range(15):
        assert not m.note_decode_step([s[1]])                     # inside the reservation
    assert m.note_decode_step([s[1]]) and m.free_blocks == 6      # 17th token needs a new block
    m.free_sequences(s)
    assert m.free_blocks == 10
    with pytest.raises(ValueError):
        Mgr(num_layers=1, num_heads=4, emb_dim=256, kv_heads=2, tensor_parallel_size=4, device="cpu", total_num_gpu_blocks=2)


def _decoding_batch(mods, batch_id, first_id, lens, n_out, mgr):
    """a batch as generate_token(first=True) leaves it, without running a model"""
    pb, Batch, tok, State = mods["pb"], mods["Batch"], mods["tok"], mods["State"]
    msg = pb.Batch(id=batch_id, requests=[_req(pb, first_id + i, "test " * L, L, n_out, min_new_tokens=n_out) for i, L in enumerate(lens)])
    b, _ = Batch.from_pb(msg, tok, torch.float16, torch.device("cpu"), None, None, True)
    sids = mgr.allocate_tokens(lens, reserve_tokens=[n_out] * len(lens))
    b.past_key_values = State(sids, mgr.block_table_tensor(sids), torch.tensor(lens, dtype=torch.int32),
                              torch.empty(len(lens), dtype=torch.int64), 0)
    b.kv_cache_manager = mgr
    b.position_ids = torch.tensor(lens)
    b.input_ids = torch.full((len(lens),), 3)
    b.cu_seqlens_q = torch.arange(len(lens) + 1, dtype=torch.int32)
    b.input_lengths = [L + 1 for L in lens]
    b.max_seqlen += 1
    b.cu_seqlens = b.cu_seqlens + b.cu_seqlens_q
    return b


def test_concatenate_and_prune_edit_block_tables_only(mods):
    Batch, Mgr = mods["Batch"], mods["Mgr"]
    mgr = Mgr(num_layers=1, num_heads=2, emb_dim=128, kv_heads=2, device="cpu", total_num_gpu_blocks=64)
    a = _decoding_batch(mods, 0, 0, [5, 40], 8, mgr)
    b = _decoding_batch(mods, 1, 2, [17], 30, mgr)
    used = mgr.total_num_gpu_blocks - mgr.free_blocks
    pool_before = mgr.pool.clone()
    rows_a, rows_b = a.past_key_values.block_table.clone(), b.past_key_values.block_table.clone()
    c = Batch.concatenate([a, b])
    assert torch.equal(mgr.pool, pool_before) and mgr.total_num_gpu_blocks - mgr.free_blocks == used
    assert len(c) == 3 and c.batch_id == 0 and [r.id for r in c.requests] == [0, 1, 2]
    kv = c.past_key_values
    assert kv.block_table.shape == (3, 3) and kv.context_lens.tolist() == [5, 40, 17]
    assert torch.equal(kv.block_table[:2, :rows_a.shape[1]], rows_a) and torch.equal(kv.block_table[2, :rows_b.shape[1]], rows_b[0])
    assert c.cu_seqlens.tolist() == [0, 6, 47, 65] and c.cu_seqlens_q.tolist() == [0, 1, 2, 3]
    assert c.input_lengths == [6, 41, 18] and c.max_seqlen == 41 and c.all_input_ids_tensor.shape == (3, 48)
    assert a.past_key_values is None and b.past_key_values is None  # inputs released (flash_causal_lm.py:246)
    lens_before = list(c.input_lengths)
    c = Batch.concatenate([c])                                        # concatenate([single]) is legal
    assert c.input_lengths == lens_before and c.past_key_values.context_lens.tolist() == [5, 40, 17]
    # prune the middle request: its blocks come back, the survivors' rows are untouched
    seq1_blocks = list(mgr.sequence_blocks(c.past_key_values.sequence_ids[1]))
    free_before = mgr.free_blocks
    p = Batch.prune(c, [1])
    assert p is c and len(p) == 2 and [r.id for r in p.requests] == [0, 2]
    assert mgr.free_blocks == free_before + len(seq1_blocks)
    assert p.past_key_values.context_lens.tolist() == [5, 17] and p.input_lengths == [6, 18]
    assert p.cu_seqlens.tolist() == [0, 6, 24] and p.cu_seqlens_q.tolist() == [0, 1, 2]
    assert p.next_token_chooser.min_new_tokens == [8, 30]
    assert Batch.prune(p, []) is p
    assert Batch.prune(p, [0, 2]) is None and mgr.free_blocks == mgr.total_num_gpu_blocks


# ---------------------------------------------------------------------------------------------- world size 2 over gloo
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _tp_worker(rank, world, port, path, quant, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import tgis_b200  # noqa: F401
    from tgis_b200.utils.dist import initialize_torch_distributed
    from tgis_b200.utils.weights import Weights
    pg = initialize_torch_distributed(world, rank)        # gloo on CPU (utils/dist.py:81-83)
    w = Weights([path], device="cpu", dtype=torch.float16, process_group=pg)
    if quant:
        w.gptq_bits, w.gptq_groupsize = 4, 128
    cfg = oll.LlamaConfig(256, 512, 1, 4, 2, 512)
    p = "model.layers.0"
    col = w.get_multi_weights_col([f"{p}.mlp.gate_proj", f"{p}.mlp.up_proj"], quant, 0)
    row = w.get_multi_weights_row(f"{p}.mlp.down_proj", quant)
    if quant:
        gate_up = oll.Linear(qweight=col[0], qzeros=col[1], scales=col[2], g_idx=col[3], groupsize=128)
        down = oll.Linear(qweight=row[0], qzeros=row[1], scales=row[2], g_idx=row[3], groupsize=128)
    else:
        gate_up, down = oll.Linear(weight=col), oll.Linear(weight=row)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, cfg.hidden_size, generator=g).half()
    part = down(oll.silu_mul(gate_up(x), cfg.intermediate_size // world)).float()
    dist.all_reduce(part)                                  # row-parallel all-reduce (utils/layers.py:318-322)
    # vocab-parallel embedding: out-of-shard ids contribute zeros, then all-reduce (utils/layers.py:346-357)
    emb = w.get_partial_sharded("model.embed_tokens.weight", dim=0)
    ids = torch.tensor([0, 255, 256, 511])
    block = cfg.vocab_size // world
    local = ids - rank * block
    ok = (local >= 0) & (local < block)
    e = torch.zeros(4, cfg.hidden_size)
    e[ok] = emb[local[ok]].float()
    from tgis_b200.utils.p2p import LayerBoundaryAllReduce
    reduce = LayerBoundaryAllReduce(pg)                    # what FlashLlama's step calls at the layer boundary
    assert not reduce.uses_peer_memory                     # the peer-memory kernel is opt-in and GPU-only: gloo here
    assert reduce(e) is e
    if rank == 0:
        out_q.put((part, e))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("quant", [None, "gptq"])
def test_world_size_2_sharded_mlp_and_embedding_over_gloo(tmp_path, quant):
    from safetensors.torch import save_file
    cfg = oll.LlamaConfig(256, 512, 1, 4, 2, 512)
    sd = oll.make_state_dict(cfg, seed=8, quantize=quant)
    path = os.path.join(str(tmp_path), "m.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tp_worker, args=(r, 2, port, path, quant, q)) for r in range(2)]
    for p in procs:
        p.start()
    part, e = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = oll.build_shards(cfg, sd, 1)[0].layers[0]
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, cfg.hidden_size, generator=g).half()
    ref = full.down(oll.silu_mul(full.gate_up(x), cfg.intermediate_size)).float()
    assert (part - ref).abs().max().item() <= 4e-3 * ref.abs().max().item() + 1e-3   # fp16 partial sums vs one fp32 accumulation
    assert torch.equal(e, sd["model.embed_tokens.weight"][[0, 255, 256, 511]].float())


# ---------------------------------------------------------------------------------------------- reference bookkeeping
def test_batch_bookkeeping_matches_reference_flash_causal_lm(mods):
    """tests/golden/batch_bookkeeping.npz was produced by the REFERENCE's FlashCausalLMBatch.from_pb /
    FlashCausalLM.generate_token / concatenate / prune (CPU, stand-in model.forward).  The same session through this
    repo's classes (same stand-in forward; the device kernels of the step replaced by their torch equivalents) must
    leave identical ids, positions, cu_seqlens, all_input_ids_tensor, lengths, chooser counters and emitted tokens."""
    import numpy as np
    import types
    from tgis_b200.models.flash_causal_lm import FlashCausalLM
    pb, Batch, tok, Mgr = mods["pb"], mods["Batch"], mods["tok"], mods["Mgr"]
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "batch_bookkeeping.npz"))
    V = 64

    def fake_logits(input_ids, position_ids):
        g = (input_ids.to(torch.int64) * 7919 + position_ids.to(torch.int64) * 104729) % 1000003
        base = torch.arange(V, dtype=torch.int64)[None, :]
        return (((g[:, None] * (base + 3)) % 97).float() / 9.7 - 5.0).to(torch.float16)

    class FakeModel:
        def forward(self, input_ids, position_ids, cu_seqlens, cu_seqlens_q, max_s, inputs_embeds, past_key_values, prealloc,
                    lm_head_indices=None):
            lg = fake_logits(input_ids, position_ids)
            return (lg if lm_head_indices is None else lg[lm_head_indices]), past_key_values

    class HostOnlyLM(FlashCausalLM):
        def __init__(self):
            self.model = FakeModel()
            self.device = torch.device("cpu")
            self.tokenizer = tok
            self.engine = types.SimpleNamespace(world_size=1)
            self.kv_cache_manager = Mgr(num_layers=1, num_heads=1, emb_dim=64, kv_heads=1, device="cpu", total_num_gpu_blocks=32)

        def _can_fuse_greedy(self, batch):
            return False

        def _decode_advance(self, batch):  # torch equivalent of b200_decode_advance
            kv = batch.past_key_values
            pos = kv.context_lens.to(torch.int64)
            blk = kv.block_table.to(torch.int64).gather(1, (pos // 16)[:, None])[:, 0]
            kv.slot_mapping[:len(batch)] = blk * 16 + pos % 16
            batch.position_ids.copy_(pos)
            kv.context_lens += 1

    lm = HostOnlyLM()

    def check(tag, b, toks=None):
        assert np.array_equal(b.input_ids.numpy(), z[f"{tag}_input_ids"]), tag
        assert np.array_equal(b.position_ids.numpy(), z[f"{tag}_position_ids"]), tag
        assert np.array_equal(b.cu_seqlens.numpy(), z[f"{tag}_cu_seqlens"]), tag
        ref_all = z[f"{tag}_all_input_ids"]
        assert np.array_equal(b.all_input_ids_tensor.numpy()[:, :ref_all.shape[1]], ref_all), tag
        assert list(b.input_lengths) == z[f"{tag}_input_lengths"].tolist(), tag
        assert list(b.total_lengths) == z[f"{tag}_total_lengths"].tolist(), tag
        assert b.max_seqlen == int(z[f"{tag}_max_seqlen"]), tag
        assert [r.id for r in b.requests] == z[f"{tag}_request_ids"].tolist(), tag
        assert list(b.next_token_chooser.current_tokens) == z[f"{tag}_current_tokens"].tolist(), tag
        if toks is not None:
            assert [[t.request_id, t.token_id] for t in toks] == z[f"{tag}_tokens"].tolist(), tag

    msg_a = pb.Batch.FromString(z["msg_a"].tobytes())
    msg_b = pb.Batch.FromString(z["msg_b"].tobytes())
    A, errs = Batch.from_pb(msg_a, tok, torch.float16, torch.device("cpu"), None, None, True)
    check("a0", A)
    toks = lm.generate_token(A, first=True)[0]
    check("a1", A, toks)
    for s in range(2):
        toks = lm.generate_token(A)[0]
        check(f"a{2 + s}", A, toks)
    Bb, _ = Batch.from_pb(msg_b, tok, torch.float16, torch.device("cpu"), None, None, True)
    toks = lm.generate_token(Bb, first=True, for_concat=True)[0]
    check("b1", Bb, toks)
    C = Batch.concatenate([A, Bb])
    check("c0", C)
    toks = lm.generate_token(C)[0]
    check("c1", C, toks)
    C = Batch.prune(C, [1])
    check("p0", C)
    toks = lm.generate_token(C)[0]
    check("p1", C, toks)
    # the paged state agrees with the bookkeeping: context = tokens in cache, one slot per sequence for the next step
    assert C.past_key_values.context_lens.tolist() == [L - 1 for L in C.input_lengths]


def test_speculation_hooks_of_the_kv_manager(mods):
    """add_child_sequences / remove_tokens / recursive free (models/paged_causal_lm.py:481-562, utils/paged.py:185-203, 309-315 of the
    reference; fms-extras semantics): candidates share their parent's blocks, the first write into a shared partial block copies it,
    losers give everything back, the winner forgets its rejected tail, freeing the survivor recursively frees its ancestors."""
    from tgis_b200.utils import paged
    m = mods["Mgr"](num_layers=1, num_heads=2, emb_dim=128, kv_heads=2, device="cpu", total_num_gpu_blocks=16)
    [parent] = m.allocate_tokens([20])
    m.pool[:, :, m.sequence_blocks(parent)[1]] = 7.0
    pos, cd, children = paged.prepare_candidates([parent], n_candidates=3, n_tokens=3, kv_cache_manager=m)
    kids = children[0]
    assert m.free_blocks == 11  # 2 parent blocks + one private copy of the partial block per child
    assert all(m.sequence_blocks(k)[0] == m.sequence_blocks(parent)[0] for k in kids)
    assert len({m.sequence_blocks(k)[1] for k in kids} | {m.sequence_blocks(parent)[1]}) == 4
    assert float(m.pool[0, 0, m.sequence_blocks(kids[2])[1]].min()) == 7.0  # copy-on-write kept the parent's tokens
    # generation form: one row per query token, contexts 21, 22, 23 for every candidate, slots 4..6 of its own block
    assert pos.tolist() == [20, 21, 22] * 3 and cd.context_lengths.tolist() == [21, 22, 23] * 3
    assert cd.block_mapping.shape == (9, 2) and cd.is_filled()
    assert cd.slot_mapping.tolist() == [m.sequence_blocks(k)[1] * 16 + o for k in kids for o in (4, 5, 6)]
    # candidate inputs: [last accepted token, 2 speculated]; candidate 1 speculated both right, candidate 0 one, candidate 2 none
    fed = torch.tensor([[[5, 8, 3], [5, 8, 9], [5, 1, 1]]])
    nxt = torch.tensor([[[8, 9, 4], [8, 9, 6], [8, 2, 2]]])
    survivors, accepted = paged.accept_candidates(fed, nxt, children, m)
    assert survivors == [kids[1]] and accepted == [[8, 9, 6]]
    assert m.sequence_length(kids[1]) == 23 and m.free_blocks == 13
    # a winner with a wrong tail forgets it
    pos, cd, children = paged.prepare_candidates(survivors, n_candidates=2, n_tokens=4, kv_cache_manager=m)
    fed = torch.tensor([[[6, 1, 1, 1], [6, 7, 1, 1]]])
    nxt = torch.tensor([[[7, 2, 2, 2], [7, 2, 2, 2]]])
    survivors, accepted = paged.accept_candidates(fed, nxt, children, m)
    assert accepted == [[7, 2]] and m.sequence_length(survivors[0]) == 23 + 2
    m.free_sequences(survivors, recursive=True)  # the survivor, its parent candidate and the original sequence (server.py:249)
    assert m.free_blocks == 16
    # prefill form
    pos, cd = paged.prepare_inputs_for_prefill([3, 18], m)
    assert not cd.is_filled() and cd.context_lengths.tolist() == [0, 3, 21] and pos.tolist() == [0, 1, 2] + list(range(18))
    assert cd.slot_mapping.shape == (21,) and cd.block_mapping.shape == (2, 2)
    pos, cd = paged.prepare_inputs_without_speculation(cd.sequence_ids, m)
    assert pos.tolist() == [3, 18] and cd.context_lengths.tolist() == [4, 19]
