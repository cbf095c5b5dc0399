"""GPU parity tests: every C-ABI kernel against the CPU oracle on the same seeded inputs.

Tolerances (written here as the task demands): fp16 outputs must match the fp32-arithmetic oracle within
1e-3 relative of the tensor's magnitude scale plus one fp16 ulp; integer outputs (arg-max ids, KV pool bytes
written by copy) bit-exact.
"""
import math

import pytest
import torch

from oracle import gptq as ogptq
from oracle import llama as oll

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    import tgis_b200  # noqa: F401
    from tgis_b200 import ops as _ops
    return _ops


def _close(got: torch.Tensor, ref: torch.Tensor, rel=1e-3, what=""):
    got = got.float().cpu()
    ref = ref.float().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs()
    tol = rel * scale + 2.0 ** -10 * ref.abs()  # 1e-3 of scale + 1 fp16 ulp of the element
    bad = err > tol
    if bad.any():
        idx = torch.nonzero(bad)[0].tolist()
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} elements off; max err {err.max().item():.4e} "
                             f"(scale {scale:.3e}); first bad at {idx}: got {got[tuple(idx)].item()} ref {ref[tuple(idx)].item()}")


# ------------------------------------------------------------------------------------------ elementwise
@pytest.mark.parametrize("T,H", [(1, 2048), (64, 4096), (7, 8192), (3, 256)])
@pytest.mark.parametrize("with_res", [False, True])
def test_rmsnorm_residual(ops, T, H, with_res):
    g = torch.Generator().manual_seed(T * 1000 + H)
    h = torch.randn(T, H, generator=g).half()
    r = torch.randn(T, H, generator=g).half() if with_res else None
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).half()
    ref_n, ref_r = oll.rmsnorm_residual(h, r, gamma, 1e-5)
    n, ro = ops.rmsnorm_residual(h.to(DEV), None if r is None else r.to(DEV), gamma.to(DEV), 1e-5)
    _close(n, ref_n, what="normed")
    if with_res:
        assert torch.equal(ro.cpu(), ref_r), "residual_out must be bit-exact (fp32 add, one rounding)"
    else:
        assert torch.equal(ro.cpu(), h)


@pytest.mark.parametrize("T,I", [(1, 5632), (64, 11008), (5, 128)])
def test_silu_mul(ops, T, I):
    g = torch.Generator().manual_seed(I)
    gu = (torch.randn(T, 2 * I, generator=g) * 2).half()
    ref = oll.silu_mul(gu, I)
    got = ops.silu_mul(gu.to(DEV))
    _close(got, ref, what="silu_mul")


def test_embedding_and_argmax(ops):
    g = torch.Generator().manual_seed(3)
    table = torch.randn(1000, 256, generator=g).half()
    ids = torch.tensor([0, 999, 5, 500, 250], dtype=torch.int64)
    out = ops.embedding(table.to(DEV), ids.to(DEV))
    assert torch.equal(out.cpu(), table[ids])
    # vocab-parallel shard [250, 500): out-of-shard ids -> zero rows (utils/layers.py:344-353)
    out = ops.embedding(table[250:500].contiguous().to(DEV), ids.to(DEV), vocab_start=250)
    exp = torch.zeros(5, 256, dtype=torch.float16)
    exp[4] = table[250]
    assert torch.equal(out.cpu(), exp)
    logits = torch.randn(9, 32000, generator=g).half()
    logits[3, 17] = 100.0
    logits[4, :] = -float("inf")
    logits[5, 31999] = 50.0
    logits[6, 100] = logits[6, 200] = 60.0  # tie -> first index
    got = ops.argmax(logits.to(DEV)).cpu()
    assert torch.equal(got, logits.float().argmax(-1)), (got, logits.float().argmax(-1))
    assert got[6].item() == 100


@pytest.mark.parametrize("h,kv,d", [(8, 8, 128), (32, 4, 64), (8, 2, 128)])
@pytest.mark.parametrize("T", [37, 700])  # 700: the vectorised prefill-sized form (T >= 256)
def test_rope_kv_write_paged(ops, h, kv, d, T):
    g = torch.Generator().manual_seed(h * d)
    nblocks = (T + 15) // 16 + 9
    qkv = torch.randn(T, (h + 2 * kv) * d, generator=g).half()
    pos = torch.randint(0, 300, (T,), generator=g)
    cos_t, sin_t = oll.rope_tables(d, 10000.0, 300)
    slots = torch.randperm(nblocks * 16, generator=g)[:T].to(torch.int64)
    slots[5] = -1  # padding token: rotated in place but not stored
    k_pool, v_pool = ops.kv_pool_alloc(nblocks, kv, d, DEV)
    qkv_d = qkv.to(DEV)
    ops.rope_kv_write_paged(qkv_d, cos_t.to(DEV), sin_t.to(DEV), pos.to(DEV), slots.to(DEV), k_pool, v_pool, h, kv, d)
    q, k, v = qkv.split([h * d, kv * d, kv * d], dim=1)
    q_ref = oll.apply_rotary(q.reshape(T, h, d), cos_t[pos], sin_t[pos])
    k_ref = oll.apply_rotary(k.reshape(T, kv, d), cos_t[pos], sin_t[pos])
    got = qkv_d.cpu()
    gq, gk, gv = got.split([h * d, kv * d, kv * d], dim=1)
    assert torch.equal(gq.reshape(T, h, d), q_ref), "rotated q must be bit-exact (fp32 math, one rounding)"
    assert torch.equal(gk.reshape(T, kv, d), k_ref)
    assert torch.equal(gv, v)
    kl = ops.kv_pool_unswizzle(k_pool).cpu()
    vl = ops.kv_pool_unswizzle(v_pool).cpu()
    exp_k = torch.zeros_like(kl)
    exp_v = torch.zeros_like(vl)
    for t in range(T):
        s = int(slots[t])
        if s < 0:
            continue
        exp_k[s // 16, :, s % 16] = k_ref[t]
        exp_v[s // 16, :, s % 16] = v.reshape(T, kv, d)[t]
    assert torch.equal(kl, exp_k) and torch.equal(vl, exp_v)


# ------------------------------------------------------------------------------------------ attention
def _paged_case(ops, lens, h, kv, d, seed):
    g = torch.Generator().manual_seed(seed)
    B = len(lens)
    pages = [(L + 15) // 16 for L in lens]
    nblocks = sum(pages) + 3
    perm = torch.randperm(nblocks, generator=g).tolist()
    max_pages = max(pages)
    bt = torch.zeros(B, max_pages + 2, dtype=torch.int32)
    k_log = torch.zeros(nblocks, kv, 16, d, dtype=torch.float16)
    v_log = torch.zeros_like(k_log)
    ks, vs = [], []
    cur = 0
    for b, L in enumerate(lens):
        k = torch.randn(L, kv, d, generator=g).half()
        v = torch.randn(L, kv, d, generator=g).half()
        ks.append(k)
        vs.append(v)
        for p in range(pages[b]):
            blk = perm[cur]
            cur += 1
            bt[b, p] = blk
            n = min(16, L - p * 16)
            k_log[blk, :, :n] = k[p * 16:p * 16 + n].transpose(0, 1)
            v_log[blk, :, :n] = v[p * 16:p * 16 + n].transpose(0, 1)
    q = torch.randn(B, h, d, generator=g).half()
    k_pool = ops.kv_pool_unswizzle(k_log.to(DEV))  # XOR swizzle is an involution
    v_pool = ops.kv_pool_unswizzle(v_log.to(DEV))
    return q, ks, vs, k_pool, v_pool, bt


@pytest.mark.parametrize("h,kv,d", [(8, 8, 128), (32, 4, 64), (16, 2, 128), (16, 1, 64)])
def test_attn_decode_paged(ops, h, kv, d):
    lens = [1, 15, 16, 17, 64, 511, 512, 513, 1029, 2047]
    q, ks, vs, k_pool, v_pool, bt = _paged_case(ops, lens, h, kv, d, seed=h + d)
    scale = d ** -0.5
    ref = oll.attention_decode(q, ks, vs, scale)
    ctx = torch.tensor(lens, dtype=torch.int32)
    # q as a strided view of a fused qkv activation, like the model passes it
    qkv = torch.zeros(len(lens), (h + 2 * kv) * d, dtype=torch.float16)
    qkv[:, :h * d] = q.reshape(len(lens), -1)
    qkv_d = qkv.to(DEV)
    q_view = qkv_d[:, :h * d].view(len(lens), h, d)
    got = ops.attn_decode_paged(q_view, k_pool, v_pool, bt.to(DEV), ctx.to(DEV), max(lens), scale, kv)
    torch.cuda.synchronize()
    _close(got, ref, rel=2e-3, what=f"attn_decode h{h} kv{kv} d{d}")


@pytest.mark.parametrize("h,kv,d", [(4, 4, 128), (8, 2, 64), (8, 1, 128)])
@pytest.mark.parametrize("causal", [True, False])
def test_attn_prefill_varlen(ops, h, kv, d, causal):
    g = torch.Generator().manual_seed(h * 7 + d)
    lens = [1, 63, 64, 65, 200, 5]
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    T = cu[-1]
    qkv = torch.randn(T, (h + 2 * kv) * d, generator=g).half()
    q, k, v = qkv.split([h * d, kv * d, kv * d], dim=1)
    q, k, v = q.reshape(T, h, d), k.reshape(T, kv, d), v.reshape(T, kv, d)
    scale = d ** -0.5
    if causal:
        ref = oll.attention_prefill(q, k, v, cu, scale)
    else:
        ref = torch.cat([oll.attention_decode(q[cu[b]:cu[b + 1]], [k[cu[b]:cu[b + 1]]] * lens[b], [v[cu[b]:cu[b + 1]]] * lens[b], scale)
                         for b in range(len(lens))])
    qkv_d = qkv.to(DEV)
    qd, kd, vd = qkv_d.split([h * d, kv * d, kv * d], dim=1)
    got = ops.attn_prefill_varlen(qd.view(T, h, d), kd.view(T, kv, d), vd.view(T, kv, d),
                                  torch.tensor(cu, dtype=torch.int32, device=DEV), max(lens), scale, causal)
    torch.cuda.synchronize()
    _close(got, ref, rel=2e-3, what=f"attn_prefill h{h} kv{kv} d{d} causal={causal}")


@pytest.mark.parametrize("h,kv,d", [(4, 4, 128), (8, 2, 128), (8, 2, 64), (4, 1, 128)])
def test_attn_prefill_paged_with_cached_context(ops, h, kv, d):
    """tcgen05 / TMA prefill attention through the block pool: every sequence has `past` cached tokens in front of `n_q` new
    query tokens (past = 0: a plain prefill; n_q = 1: a decode row; ragged mixes of both, tile boundaries 127 / 128 / 129, a
    context that ends in the middle of a page).  Oracle: causal attention over the whole sequence, rows of the new tokens."""
    cases = [(0, 1), (0, 127), (0, 128), (0, 129), (0, 300), (5, 60), (100, 1), (128, 128), (200, 57), (1000, 40), (17, 260)]
    lens = [p + n for p, n in cases]
    _, ks, vs, k_pool, v_pool, bt = _paged_case(ops, lens, h, kv, d, seed=3 * h + d)
    g = torch.Generator().manual_seed(99)
    scale = d ** -0.5
    qs, refs, cu = [], [], [0]
    for (past, n_q), k, v in zip(cases, ks, vs):
        L = past + n_q
        q_full = torch.randn(L, h, d, generator=g).half()
        ref_full = oll.attention_prefill(q_full, k, v, [0, L], scale)
        qs.append(q_full[past:])
        refs.append(ref_full[past:])
        cu.append(cu[-1] + n_q)
    q = torch.cat(qs)
    T = q.shape[0]
    qkv = torch.zeros(T, (h + 2 * kv) * d, dtype=torch.float16)
    qkv[:, :h * d] = q.reshape(T, -1)
    qkv_d = qkv.to(DEV)
    got = ops.attn_prefill_paged(qkv_d[:, :h * d].view(T, h, d), k_pool, v_pool, bt.to(DEV), torch.tensor(lens, dtype=torch.int32, device=DEV),
                                 torch.tensor(cu, dtype=torch.int32, device=DEV), max(n for _, n in cases), scale)
    torch.cuda.synchronize()
    ref = torch.cat(refs)
    for b, (past, n_q) in enumerate(cases):  # per sequence, so a failure names the case
        _close(got[cu[b]:cu[b + 1]], ref[cu[b]:cu[b + 1]], rel=2e-3, what=f"attn_prefill_paged h{h} kv{kv} d{d} past={past} n_q={n_q}")


def test_masked_softmax_and_fused_attention(ops):
    """b200_masked_softmax vs the restated semantics of forward_masked_softmax_kernel (custom_kernels/fused_attention_cuda.cu:28-107):
    fp32 softmax over unmasked positions, masked -> 0, all-masked row -> zeros; kv beyond the reference's 4096 limit; and the
    `fused_attention_cuda.forward` wrapper vs eager attention."""
    from tgis_b200.custom_kernels import fused_attention_cuda
    g = torch.Generator().manual_seed(4)
    for dtype, kv in ((torch.float16, 37), (torch.float32, 300), (torch.float16, 5000)):
        rows = 19
        sc = (3 * torch.randn(rows, kv, generator=g)).to(dtype)
        mask = torch.rand(rows, kv, generator=g) < 0.3
        mask[3] = True  # an all-masked row
        got = ops.masked_softmax(sc.to(DEV), mask.to(DEV)).cpu()
        ref = torch.softmax(sc.float().masked_fill(mask, float("-inf")), -1)
        ref = torch.nan_to_num(ref, nan=0.0).masked_fill(mask, 0.0).to(dtype)
        assert torch.equal(got[3], torch.zeros(kv, dtype=dtype))
        assert (got.float() - ref.float()).abs().max().item() <= (1e-3 if dtype == torch.float16 else 1e-6)
    B, h, q, past, d = 2, 3, 4, 6, 64
    query = torch.randn(B, h, q, d, generator=g).half()
    key = torch.randn(B, h, q, d, generator=g).half()
    value = torch.randn(B, h, q, d, generator=g).half()
    pk = torch.randn(B, h, past, d, generator=g).half()
    pv = torch.randn(B, h, past, d, generator=g).half()
    kvl = past + q
    causal = torch.ones(q, kvl, dtype=torch.bool).triu(past + 1)[None, None].expand(B, 1, q, kvl)
    ctx, present, probs = fused_attention_cuda.forward(query.to(DEV), key.to(DEV), value.to(DEV), [pk.to(DEV), pv.to(DEV)], causal.to(DEV),
                                                       None, d ** -0.5, h, True)
    k_all, v_all = torch.cat([pk, key], 2).float(), torch.cat([pv, value], 2).float()
    s = (query.float() * d ** -0.5) @ k_all.transpose(-1, -2)
    p = torch.softmax(s.masked_fill(causal, float("-inf")), -1)
    ref_ctx = (p @ v_all).permute(0, 2, 1, 3).reshape(B, q, h * d)
    assert present[0].shape == (B, h, kvl, d)
    assert (ctx.float().cpu() - ref_ctx).abs().max().item() <= 2e-2
    assert (probs.float().cpu().view(B, h, q, kvl) - p).abs().max().item() <= 2e-3
