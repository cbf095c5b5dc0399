"""Test-only stand-in for `tgis_b200.ops` on a CPU box: every op the host-side modeling code calls, restated with the oracle's
arithmetic on CPU tensors (the KV pool keeps the logical [block, kv head, 16, d] layout, un-swizzled).

Purpose: run the PRODUCT's host wiring (weight loading and slicing, fused-projection layouts, views and strides handed to
the attention ops, residual forms, KV placement through the block table) end to end without a GPU and compare it with the
family's oracle.  It is never imported by the package; the kernels themselves are covered by the `-m gpu` parity tests."""
from typing import Optional

import torch

from oracle import llama as oll
from oracle import neox as onx

F16 = torch.float16
PAGE = 16


class DryRunOps:
    def __init__(self):
        self.calls = []

    # ---- linears / row ops
    def gemm_f16(self, x, w, bias=None, out=None, workspace=None):
        self.calls.append(("gemm_f16", tuple(x.shape), tuple(w.shape)))
        assert x.dtype == F16 and w.dtype == F16 and x.is_contiguous() and w.is_contiguous() and x.shape[1] == w.shape[1]
        y = onx.linear(x, w, bias)
        return y if out is None else out.copy_(y)

    def layernorm_residual(self, h, residual, gamma, beta, eps):
        return onx.layernorm_residual(h, residual, gamma, beta, eps)  # no residual in -> residual out is h, like the op

    def rmsnorm_residual(self, h, residual, gamma, eps):
        return oll.rmsnorm_residual(h, residual, gamma, eps)

    def gelu(self, x, approximate_tanh=False):
        return onx.gelu(x, approximate_tanh)

    def embedding(self, table, ids, vocab_start=0):
        assert ids.dtype == torch.int64 and ids.dim() == 1
        local = ids - vocab_start
        ok = (local >= 0) & (local < table.shape[0])
        out = torch.zeros(ids.shape[0], table.shape[1], dtype=F16)
        out[ok] = table[local[ok]]
        return out

    def argmax(self, logits, banned_ids=None, out=None):
        scores = logits.float().clone()
        if banned_ids is not None:
            for r, b in enumerate(banned_ids.tolist()):
                if b >= 0:
                    scores[r, b] = float("-inf")
        ids = scores.argmax(-1)
        return ids if out is None else out.copy_(ids)

    # ---- rotary + KV append (in place on qkv, like the kernel)
    def rope_kv_write_paged(self, qkv, cos, sin, position_ids, slot_mapping, k_pool, v_pool, n_heads, n_kv_heads, head_dim,
                            rotary_dim: Optional[int] = None):
        self.calls.append(("rope_kv_write_paged", n_heads, n_kv_heads, head_dim, rotary_dim))
        T, d = qkv.shape[0], head_dim
        assert qkv.is_contiguous() and qkv.shape[1] == (n_heads + 2 * n_kv_heads) * d
        rd = d if rotary_dim is None else rotary_dim
        assert cos.shape[-1] * 2 == rd and position_ids.dtype == torch.int64 and slot_mapping.dtype == torch.int64
        assert int(position_ids.max()) < cos.shape[0], "position beyond the rotary table"
        c, s = cos[position_ids], sin[position_ids]
        q = qkv[:, :n_heads * d].view(T, n_heads, d)
        k = qkv[:, n_heads * d:(n_heads + n_kv_heads) * d].view(T, n_kv_heads, d)
        v = qkv[:, (n_heads + n_kv_heads) * d:].view(T, n_kv_heads, d)
        q.copy_(oll.apply_rotary(q, c, s))
        k.copy_(oll.apply_rotary(k, c, s))
        assert k_pool.shape[1:] == (n_kv_heads, PAGE, d) and v_pool.shape == k_pool.shape
        for t, slot in enumerate(slot_mapping.tolist()):
            if slot >= 0:
                k_pool[slot // PAGE, :, slot % PAGE] = k[t]
                v_pool[slot // PAGE, :, slot % PAGE] = v[t]

    # ---- attention (strided views allowed exactly where the real ops allow them)
    @staticmethod
    def _check_view(t, d):
        assert t.dtype == F16 and t.stride(2) == 1 and t.stride(1) == d and t.stride(0) % 8 == 0, (t.shape, t.stride())

    def attn_prefill_varlen(self, q, k, v, cu_seqlens, max_s, softmax_scale, causal=True, out=None):
        self.calls.append(("attn_prefill_varlen", tuple(q.shape), tuple(k.shape)))
        assert causal and cu_seqlens.dtype == torch.int32
        for t in (q, k, v):
            self._check_view(t, q.shape[2])
        o = oll.attention_prefill(q, k, v, cu_seqlens.tolist(), softmax_scale)
        return o if out is None else out.copy_(o)

    def attn_decode_paged(self, q, k_pool, v_pool, block_table, context_lens, max_context_len, softmax_scale, n_kv_heads, out=None,
                          workspace=None):
        self.calls.append(("attn_decode_paged", tuple(q.shape), n_kv_heads))
        B, h, d = q.shape
        self._check_view(q, d)
        assert h % n_kv_heads == 0 and h // n_kv_heads <= 16, "the decode kernel shares a KV head among at most 16 query heads"
        assert block_table.dtype == torch.int32 and context_lens.dtype == torch.int32 and k_pool.shape[1] == n_kv_heads
        assert int(context_lens.max()) <= max_context_len
        ks, vs = [], []
        for b in range(B):
            L = int(context_lens[b])
            blocks = block_table[b, :(L + PAGE - 1) // PAGE].long()
            ks.append(k_pool[blocks].permute(0, 2, 1, 3).reshape(-1, n_kv_heads, d)[:L])   # [pages, kv, 16, d] -> [L, kv, d]
            vs.append(v_pool[blocks].permute(0, 2, 1, 3).reshape(-1, n_kv_heads, d)[:L])
        o = oll.attention_decode(q, ks, vs, softmax_scale)
        if out is None:
            return o
        assert out.stride(2) == 1 and out.stride(1) == d
        return out.copy_(o)


def patch_ops(monkeypatch, *modules):
    """Points the `_ops` accessor of the given modules (and of the shared layers / attention wrappers) at a DryRunOps."""
    import tgis_b200  # noqa: F401
    from tgis_b200.models.custom_modeling import python_step
    from tgis_b200.utils import flash_attn, layers
    fake = DryRunOps()
    for m in (layers, flash_attn, python_step, *modules):
        if hasattr(m, "_ops"):
            monkeypatch.setattr(m, "_ops", lambda: fake)
    return fake
