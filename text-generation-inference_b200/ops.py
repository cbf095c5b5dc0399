"""Tensor-level wrappers over the C ABI.  PyTorch only supplies device memory and the current stream.

Every function requires CUDA tensors and raises if the library is missing — no eager fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

PAGE = 16


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.B200Error(f"{name}: expected a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.B200Error(f"{name}: expected {dtype}, got {t.dtype}")


def rmsnorm_residual(h: torch.Tensor, residual: Optional[torch.Tensor], gamma: torch.Tensor, eps: float
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(normed, residual_out) — LlamaRMSNorm.forward semantics (flash_llama_modeling.py:113-152)."""
    _req(h, torch.float16, "h")
    h = h.contiguous()
    T, H = h.shape
    normed = torch.empty_like(h)
    if residual is not None:
        residual = residual.contiguous()
        res_out = torch.empty_like(h)
    else:
        res_out = None
    _lib.check(_lib.load().b200_rmsnorm_residual(_ptr(h), _ptr(residual), _ptr(gamma), _ptr(normed), _ptr(res_out),
                                                 T, H, float(eps), _stream()), "rmsnorm_residual")
    return normed, (res_out if res_out is not None else h)


def layernorm_residual(h: torch.Tensor, residual: Optional[torch.Tensor], gamma: torch.Tensor, beta: Optional[torch.Tensor],
                       eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """(normed, residual_out) — FastLayerNorm.forward semantics (utils/layers.py:360-392)."""
    _req(h, torch.float16, "h")
    h = h.contiguous()
    T, H = h.shape
    normed = torch.empty_like(h)
    if residual is not None:
        residual = residual.contiguous()
        res_out = torch.empty_like(h)
    else:
        res_out = None
    _lib.check(_lib.load().b200_layernorm_residual(_ptr(h), _ptr(residual), _ptr(gamma), _ptr(beta), _ptr(normed), _ptr(res_out),
                                                   T, H, float(eps), _stream()), "layernorm_residual")
    return normed, (res_out if res_out is not None else h)


def masked_softmax(scores: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Row softmax of `scores` [..., kv] (fp16 / fp32) over the positions where `mask` (bool, same shape) is False; the
    forward_masked_softmax_kernel of server/custom_kernels (fp32 math, masked -> 0, all-masked row -> zeros)."""
    if scores.dtype not in (torch.float16, torch.float32) or scores.device.type != "cuda":
        raise _lib.B200Error("masked_softmax: need a CUDA fp16 or fp32 tensor")
    assert mask.dtype == torch.bool and mask.shape == scores.shape
    s2 = scores.contiguous()
    m2 = mask.contiguous()
    out = torch.empty_like(s2)
    kv = s2.shape[-1]
    _lib.check(_lib.load().b200_masked_softmax(_ptr(s2), _ptr(m2), _ptr(out), s2.numel() // max(kv, 1), kv,
                                               int(scores.dtype == torch.float32), _stream()), "masked_softmax")
    return out


def gelu(x: torch.Tensor, approximate_tanh: bool = False) -> torch.Tensor:
    _req(x, torch.float16, "x")
    x = x.contiguous()
    out = torch.empty_like(x)
    _lib.check(_lib.load().b200_gelu(_ptr(x), _ptr(out), x.numel(), int(approximate_tanh), _stream()), "gelu")
    return out


def rope_kv_write_paged(qkv: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, position_ids: torch.Tensor,
                        slot_mapping: torch.Tensor, k_pool: torch.Tensor, v_pool: torch.Tensor, n_heads: int,
                        n_kv_heads: int, head_dim: int, rotary_dim: Optional[int] = None) -> None:
    """cos/sin: fp16 tables [max_pos, rotary_dim/2]; rotary_dim defaults to head_dim (Llama), smaller = partial (NeoX)."""
    _req(qkv, torch.float16, "qkv")
    _req(position_ids, torch.int64, "position_ids")
    _req(slot_mapping, torch.int64, "slot_mapping")
    assert qkv.is_contiguous() and cos.is_contiguous() and sin.is_contiguous()
    T = qkv.shape[0]
    rd = head_dim if rotary_dim is None else rotary_dim
    assert cos.shape[-1] * 2 == rd
    _lib.check(_lib.load().b200_rope_kv_write_paged_ex(_ptr(qkv), _ptr(cos), _ptr(sin), _ptr(position_ids), _ptr(slot_mapping),
                                                       _ptr(k_pool), _ptr(v_pool), T, n_heads, n_kv_heads, head_dim, rd, _stream()),
               "rope_kv_write_paged")


def silu_mul(gate_up: torch.Tensor) -> torch.Tensor:
    _req(gate_up, torch.float16, "gate_up")
    T, I2 = gate_up.shape
    out = torch.empty(T, I2 // 2, dtype=torch.float16, device=gate_up.device)
    _lib.check(_lib.load().b200_silu_mul(_ptr(gate_up.contiguous()), _ptr(out), T, I2 // 2, _stream()), "silu_mul")
    return out


def embedding(table: torch.Tensor, ids: torch.Tensor, vocab_start: int = 0) -> torch.Tensor:
    _req(table, torch.float16, "table")
    _req(ids, torch.int64, "ids")
    out = torch.empty(ids.shape[0], table.shape[1], dtype=torch.float16, device=table.device)
    _lib.check(_lib.load().b200_embedding(_ptr(table), _ptr(ids), _ptr(out), ids.shape[0], table.shape[1], vocab_start,
                                          table.shape[0], _stream()), "embedding")
    return out


def argmax(logits: torch.Tensor, banned_ids: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(logits, torch.float16, "logits")
    B, V = logits.shape
    assert logits.stride(1) == 1
    if out is None:
        out = torch.empty(B, dtype=torch.int64, device=logits.device)
    _lib.check(_lib.load().b200_argmax(_ptr(logits), _ptr(out), B, V, logits.stride(0), _ptr(banned_ids), _stream()), "argmax")
    return out


_attn_ws = {}


def attn_decode_paged(q: torch.Tensor, k_pool: torch.Tensor, v_pool: torch.Tensor, block_table: torch.Tensor,
                      context_lens: torch.Tensor, max_context_len: int, softmax_scale: float, n_kv_heads: int,
                      out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [B, n_heads, d] (may be a strided view of the fused qkv activation: stride(0) arbitrary, inner contiguous)."""
    _req(q, torch.float16, "q")
    _req(block_table, torch.int32, "block_table")
    _req(context_lens, torch.int32, "context_lens")
    B, h, d = q.shape
    assert q.stride(2) == 1 and q.stride(1) == d
    if out is None:
        out = torch.empty(B, h, d, dtype=torch.float16, device=q.device)
    lib = _lib.load()
    need = lib.b200_attn_decode_workspace_bytes(B, h, d, max_context_len)
    if workspace is None:  # per-device grow-only workspace
        workspace = _attn_ws.get(q.device)
        if workspace is None or workspace.numel() < need:
            workspace = torch.zeros(max(need, 1 << 22), dtype=torch.uint8, device=q.device)
            _attn_ws[q.device] = workspace
    elif workspace.numel() * workspace.element_size() < need:
        raise _lib.B200Error("attn_decode_paged: workspace too small")
    _lib.check(lib.b200_attn_decode_paged(_ptr(q), q.stride(0), _ptr(k_pool), _ptr(v_pool), _ptr(block_table),
                                          block_table.stride(0), _ptr(context_lens), _ptr(out), out.stride(0), _ptr(workspace),
                                          workspace.numel() * workspace.element_size(), B, h, n_kv_heads, d, max_context_len,
                                          float(softmax_scale), _stream()), "attn_decode_paged")
    return out


def attn_prefill_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_seqlens: torch.Tensor, max_s: int,
                        softmax_scale: float, causal: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [T,h,d], k/v [T,h_kv,d]: strided views allowed (inner two dims contiguous)."""
    _req(q, torch.float16, "q")
    _req(cu_seqlens, torch.int32, "cu_seqlens")
    T, h, d = q.shape
    h_kv = k.shape[1]
    for t in (q, k, v):
        assert t.stride(2) == 1 and t.stride(1) == d
    if out is None:
        out = torch.empty(T, h, d, dtype=torch.float16, device=q.device)
    _lib.check(_lib.load().b200_attn_prefill_varlen(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0),
                                                    _ptr(cu_seqlens), _ptr(out), out.stride(0), cu_seqlens.shape[0] - 1, max_s,
                                                    h, h_kv, d, float(softmax_scale), int(causal), _stream()),
               "attn_prefill_varlen")
    return out


def attn_prefill_paged(q: torch.Tensor, k_pool: torch.Tensor, v_pool: torch.Tensor, block_table: torch.Tensor,
                       context_lens: torch.Tensor, cu_seqlens_q: torch.Tensor, max_q: int, softmax_scale: float,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [T_q, h, d] (a strided view of the fused qkv activation is fine); pools [num_blocks, n_kv, 16, d] already holding this
    step's K / V; context_lens [B] include them.  Causal over absolute positions (queries may follow a cached context)."""
    _req(q, torch.float16, "q")
    _req(block_table, torch.int32, "block_table")
    _req(context_lens, torch.int32, "context_lens")
    _req(cu_seqlens_q, torch.int32, "cu_seqlens_q")
    T, h, d = q.shape
    assert q.stride(2) == 1 and q.stride(1) == d and k_pool.is_contiguous() and v_pool.is_contiguous()
    if out is None:
        out = torch.empty(T, h, d, dtype=torch.float16, device=q.device)
    _lib.check(_lib.load().b200_attn_prefill_paged(_ptr(q), q.stride(0), T, _ptr(k_pool), _ptr(v_pool), k_pool.shape[0], _ptr(block_table),
                                                   block_table.stride(0), _ptr(context_lens), _ptr(cu_seqlens_q), _ptr(out),
                                                   out.stride(0), cu_seqlens_q.shape[0] - 1, int(max_q), h, k_pool.shape[1], d,
                                                   float(softmax_scale), _stream()), "attn_prefill_paged")
    return out


_gemm_ws = {}


def gemm_workspace(device, T: int, N: int, K: int) -> torch.Tensor:
    """Per-device split-K workspace (counters zeroed once; the kernels re-arm them)."""
    need = _lib.load().b200_gemm_workspace_bytes_max(N, K)
    ws = _gemm_ws.get(device)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=device)
        _gemm_ws[device] = ws
    return ws


def gemm_f16(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
             workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [T,K] @ w[N,K]^T -> [T,N] fp16."""
    _req(x, torch.float16, "x")
    _req(w, torch.float16, "w")
    assert x.is_contiguous() and w.is_contiguous()
    T, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(T, N, dtype=torch.float16, device=x.device)
    if workspace is None:
        workspace = gemm_workspace(x.device, T, N, K)
    _lib.check(_lib.load().b200_gemm_f16(_ptr(x), _ptr(w), _ptr(bias), _ptr(out), T, N, K, _ptr(workspace), _stream()), "gemm_f16")
    return out


W4_LAYOUT_PLAIN, W4_LAYOUT_GATE_UP = 0, 1


def permute_columns(x: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """out[t, k'] = x[t, perm[k']] (act-order GPTQ: activations follow the packed weight's row order)."""
    _req(x, torch.float16, "x")
    _req(perm, torch.int32, "perm")
    assert x.is_contiguous() and perm.is_contiguous() and perm.shape[0] == x.shape[1]
    out = torch.empty_like(x)
    _lib.check(_lib.load().b200_permute_columns(_ptr(x), _ptr(perm), _ptr(out), x.shape[0], x.shape[1], _stream()), "permute_columns")
    return out


def gptq_pack(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor, groupsize: int,
              layout: int = W4_LAYOUT_PLAIN, row_perm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Checkpoint GPTQ tensors of one linear -> the kernel's unit-record stream (uint8 tensor).
    layout = W4_LAYOUT_GATE_UP pairs gate tile s with up tile s (fused [gate; up] projection, N = 2 I, I % 128 == 0)."""
    _req(qweight, torch.int32, "qweight")
    _req(qzeros, torch.int32, "qzeros")
    _req(scales, torch.float16, "scales")
    assert qweight.is_contiguous() and qzeros.is_contiguous() and scales.is_contiguous()
    Kw, N = qweight.shape
    K = Kw * 8
    lib = _lib.load()
    nbytes = lib.b200_gptq_packed_bytes(K, N, groupsize)
    if nbytes < 0:
        raise _lib.B200Error(f"gptq_pack: {lib.b200_last_error().decode()}")
    packed = torch.empty(nbytes, dtype=torch.uint8, device=qweight.device)
    if row_perm is not None:
        _req(row_perm, torch.int32, "row_perm")
        assert row_perm.is_contiguous() and row_perm.shape[0] == K
    _lib.check(lib.b200_gptq_pack_ex(_ptr(qweight), _ptr(qzeros), _ptr(scales), _ptr(row_perm), _ptr(packed), K, N, groupsize,
                                     layout, _stream()), "gptq_pack")
    return packed


def gemm_w4a16(x: torch.Tensor, packed: torch.Tensor, N: int, groupsize: int, bias: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None, layout: int = W4_LAYOUT_PLAIN,
               silu_mul: bool = False) -> torch.Tensor:
    """x [T,K] fp16 @ dequant(packed) -> [T,N] fp16; `packed` from gptq_pack for the same (K, N, groupsize, layout).
    silu_mul (gate|up layout only): returns [T, N/2] = SiLU(x Wgate) * (x Wup)."""
    _req(x, torch.float16, "x")
    _req(packed, torch.uint8, "packed")
    assert x.is_contiguous() and packed.is_contiguous()
    T, K = x.shape
    if out is None:
        out = torch.empty(T, N // 2 if silu_mul else N, dtype=torch.float16, device=x.device)
    if workspace is None:
        workspace = gemm_workspace(x.device, T, N, K)
    _lib.check(_lib.load().b200_gemm_w4a16_ex(_ptr(x), _ptr(packed), _ptr(bias), _ptr(out), T, N, K, groupsize, layout,
                                              int(silu_mul), _ptr(workspace), _stream()), "gemm_w4a16")
    return out


# ------------------------------------------------------------------------------------------------------
# Deferred split-K reduction (include/b200_tgis.h "B200SplitK"): the GEMM leaves fp32 partials in the workspace and the
# next kernel on the stream sums them.  The returned SplitK keeps the workspace alive; it is valid until the next GEMM
# that uses the same workspace.
# ------------------------------------------------------------------------------------------------------
class SplitK:
    def __init__(self, desc: "_lib.B200SplitK", workspace: torch.Tensor, bias: Optional[torch.Tensor]):
        self.desc, self.workspace, self.bias = desc, workspace, bias

    @property
    def ref(self):
        import ctypes
        return ctypes.byref(self.desc)

    @property
    def shape(self):
        return self.desc.T, self.desc.N


def gemm_w4a16_deferred(x: torch.Tensor, packed: torch.Tensor, N: int, groupsize: int, bias: Optional[torch.Tensor] = None,
                        workspace: Optional[torch.Tensor] = None, layout: int = W4_LAYOUT_PLAIN) -> SplitK:
    _req(x, torch.float16, "x")
    _req(packed, torch.uint8, "packed")
    assert x.is_contiguous() and packed.is_contiguous()
    T, K = x.shape
    if workspace is None:
        workspace = gemm_workspace(x.device, T, N, K)
    desc = _lib.B200SplitK()
    import ctypes
    _lib.check(_lib.load().b200_gemm_w4a16_deferred(_ptr(x), _ptr(packed), _ptr(bias), T, N, K, groupsize, layout, _ptr(workspace),
                                                    ctypes.byref(desc), _stream()), "gemm_w4a16_deferred")
    return SplitK(desc, workspace, bias)


def gemm_f16_deferred(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
                      workspace: Optional[torch.Tensor] = None) -> SplitK:
    _req(x, torch.float16, "x")
    _req(w, torch.float16, "w")
    assert x.is_contiguous() and w.is_contiguous()
    T, K = x.shape
    N = w.shape[0]
    if workspace is None:
        workspace = gemm_workspace(x.device, T, N, K)
    desc = _lib.B200SplitK()
    import ctypes
    _lib.check(_lib.load().b200_gemm_f16_deferred(_ptr(x), _ptr(w), _ptr(bias), T, N, K, _ptr(workspace), ctypes.byref(desc),
                                                  _stream()), "gemm_f16_deferred")
    return SplitK(desc, workspace, bias)


def splitk_reduce(parts: SplitK) -> torch.Tensor:
    T, N = parts.shape
    y = torch.empty(T, N, dtype=torch.float16, device=parts.workspace.device)
    _lib.check(_lib.load().b200_splitk_reduce(parts.ref, _ptr(y), _stream()), "splitk_reduce")
    return y


def rmsnorm_residual_splitk(parts: SplitK, residual: torch.Tensor, gamma: torch.Tensor, eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    _req(residual, torch.float16, "residual")
    residual = residual.contiguous()
    normed, res_out = torch.empty_like(residual), torch.empty_like(residual)
    _lib.check(_lib.load().b200_rmsnorm_residual_splitk(parts.ref, _ptr(residual), _ptr(gamma), _ptr(normed), _ptr(res_out),
                                                        float(eps), _stream()), "rmsnorm_residual_splitk")
    return normed, res_out


def rope_kv_write_paged_splitk(parts: SplitK, cos: torch.Tensor, sin: torch.Tensor, position_ids: torch.Tensor,
                               slot_mapping: torch.Tensor, k_pool: torch.Tensor, v_pool: torch.Tensor, n_heads: int,
                               n_kv_heads: int, head_dim: int) -> torch.Tensor:
    """-> qkv [T, (n_heads + 2 n_kv) d] fp16, q and k rotated; k and v also scattered into the pools."""
    T, N = parts.shape
    qkv = torch.empty(T, N, dtype=torch.float16, device=parts.workspace.device)
    _lib.check(_lib.load().b200_rope_kv_write_paged_splitk(parts.ref, _ptr(qkv), _ptr(cos), _ptr(sin), _ptr(position_ids),
                                                           _ptr(slot_mapping), _ptr(k_pool), _ptr(v_pool), n_heads, n_kv_heads,
                                                           head_dim, _stream()), "rope_kv_write_paged_splitk")
    return qkv


def splitk_silu_mul(parts: SplitK) -> torch.Tensor:
    T, N = parts.shape
    out = torch.empty(T, N // 2, dtype=torch.float16, device=parts.workspace.device)
    _lib.check(_lib.load().b200_splitk_silu_mul(parts.ref, _ptr(out), _stream()), "splitk_silu_mul")
    return out


# ------------------------------------------------------------------------------------------------------
# KV pool helpers (layout documented in DESIGN.md; used by tests and the block manager)
# ------------------------------------------------------------------------------------------------------
def kv_pool_alloc(num_blocks: int, n_kv_heads: int, head_dim: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    shape = (num_blocks, n_kv_heads, PAGE, head_dim)
    return (torch.zeros(shape, dtype=torch.float16, device=device), torch.zeros(shape, dtype=torch.float16, device=device))


def kv_pool_unswizzle(pool: torch.Tensor) -> torch.Tensor:
    """Logical [num_blocks, n_kv, 16, d] view of a swizzled pool (test/debug helper; not on the hot path)."""
    nb, hk, pg, d = pool.shape
    chunks = pool.view(nb, hk, pg, d // 8, 8)
    t = torch.arange(pg, device=pool.device)
    c = torch.arange(d // 8, device=pool.device)
    phys = (c[None, :] ^ (t[:, None] & 7))  # [16, d/8] physical chunk of logical chunk c at token t
    idx = phys[None, None, :, :, None].expand(nb, hk, pg, d // 8, 8)
    return torch.gather(chunks, 3, idx).reshape(nb, hk, pg, d)
