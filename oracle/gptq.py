"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the reference's GPTQ int4 arithmetic.

Follows, line by line:
  * pack format ......... server/text_generation_server/utils/gptq/quant_linear.py:290-345
                          (`QuantLinear.pack`: 8 nibbles per int32 along K for qweight,
                          along N for qzeros, zeros stored minus one)
  * dequant formula ..... quant_linear.py:184-192 (`zeros + 1`, `(b - zeros) * scales`,
                          fp16 product, fp32 `tl.dot` accumulate, fp16 store)
  * exllamav2 contract .. utils/gptq/exllamav2.py:14-20,100-144 (y = x @ W (+ bias), fp16 out)

Parity status: PINNED for the pack layout / `+1` zero convention — tests/golden/gptq_pack.npz is
produced by running the reference's own `QuantLinear.pack` (tests/golden/make_golden.py).
The exllamav2 CUDA kernel itself (auto-gptq 0.7.1, un-vendored) has no golden vectors in the
reference: its numerical contract is restated here from the formula of record above.
"""
from __future__ import annotations

import numpy as np
import torch


def pack_rows_int4(intweight: np.ndarray) -> np.ndarray:
    """[K, N] values in 0..15 -> qweight int32 [K/8, N]; nibble i of word r is k = 8r+i.
    quant_linear.py:313-327."""
    K, N = intweight.shape
    assert K % 8 == 0
    iw = intweight.astype(np.uint32).reshape(K // 8, 8, N)
    q = np.zeros((K // 8, N), dtype=np.uint32)
    for i in range(8):
        q |= iw[:, i, :] << np.uint32(4 * i)
    return q.view(np.int32)


def pack_cols_int4(zeros_minus_one: np.ndarray) -> np.ndarray:
    """[G, N] values in 0..15 -> qzeros int32 [G, N/8]; nibble i of word c is n = 8c+i.
    quant_linear.py:329-345 (caller already subtracted 1, :329)."""
    G, N = zeros_minus_one.shape
    assert N % 8 == 0
    z = zeros_minus_one.astype(np.uint32).reshape(G, N // 8, 8)
    q = np.zeros((G, N // 8), dtype=np.uint32)
    for i in range(8):
        q |= z[:, :, i] << np.uint32(4 * i)
    return q.view(np.int32)


def unpack_rows_int4(qweight: np.ndarray) -> np.ndarray:
    """qweight int32 [K/8, N] -> uint8 [K, N]. quant_linear.py:160-166,188."""
    q = np.ascontiguousarray(qweight).view(np.uint32)
    Kw, N = q.shape
    out = np.empty((Kw, 8, N), dtype=np.uint8)
    for i in range(8):
        out[:, i, :] = (q >> np.uint32(4 * i)) & np.uint32(15)
    return out.reshape(Kw * 8, N)


def unpack_cols_int4(qzeros: np.ndarray) -> np.ndarray:
    """qzeros int32 [G, N/8] -> uint8 [G, N] (stored value, i.e. zero-1). quant_linear.py:168,184."""
    q = np.ascontiguousarray(qzeros).view(np.uint32)
    G, Nw = q.shape
    out = np.empty((G, Nw, 8), dtype=np.uint8)
    for i in range(8):
        out[:, :, i] = (q >> np.uint32(4 * i)) & np.uint32(15)
    return out.reshape(G, Nw * 8)


def dequantize(qweight, qzeros, scales, g_idx=None, groupsize: int = 128) -> torch.Tensor:
    """-> fp16 W[K, N] = fp16( scales[g,n] * (q[k,n] - (z[g,n] + 1)) ), one fp16 rounding.
    quant_linear.py:184-192: int32 difference times an fp16 scale."""
    qweight = _np(qweight)
    qzeros = _np(qzeros)
    q = unpack_rows_int4(qweight).astype(np.int32)
    z = unpack_cols_int4(qzeros).astype(np.int32) + 1
    K, N = q.shape
    if g_idx is None:
        g = np.arange(K) // (groupsize if groupsize > 0 else K)
    else:
        g = _np(g_idx).astype(np.int64)
    diff = torch.from_numpy((q - z[g]).astype(np.float32))
    s = scales.detach().cpu().to(torch.float16).to(torch.float32)[torch.from_numpy(g)]
    # exact integer (|d| <= 16) times fp16 scale, rounded once to fp16
    return (diff * s).to(torch.float16)


def gemm_half_q_half(x: torch.Tensor, qweight, qzeros, scales, g_idx=None, groupsize: int = 128,
                     bias: torch.Tensor | None = None) -> torch.Tensor:
    """y[M,N] = fp16( fp32-accumulate( x fp16 @ dequant(W) fp16 ) ) (+ bias).
    exllamav2.py:14-20 / 139-144; accumulate precision per quant_linear.py:171,193."""
    w = dequantize(qweight, qzeros, scales, g_idx, groupsize)
    y = (x.detach().cpu().to(torch.float16).float() @ w.float()).to(torch.float16)
    if bias is not None:
        y = y + bias.cpu().to(torch.float16)
    return y


def quantize_rtn(weight_nk: torch.Tensor, groupsize: int = 128, seed: int | None = None):
    """Synthetic-checkpoint helper: round-to-nearest group quantization of an
    fp W[N,K] (nn.Linear layout) into reference-format tensors
    (qweight [K/8,N], qzeros [K/g,N/8], scales [K/g,N] fp16, g_idx [K]).
    Packing identical to QuantLinear.pack (quant_linear.py:290-345)."""
    w = weight_nk.detach().cpu().float().t().contiguous()  # [K, N]
    K, N = w.shape
    g = groupsize if groupsize > 0 else K
    G = K // g
    wg = w.reshape(G, g, N)
    wmax = wg.amax(1)
    wmin = wg.amin(1)
    scale = ((wmax - wmin).clamp(min=1e-5) / 15.0).to(torch.float16)
    zero = torch.round(-wmin / scale.float()).clamp(1, 16)  # true zero in 1..16, stored zero-1
    q = torch.round(wg / scale.float()[:, None, :] + zero[:, None, :]).clamp(0, 15)
    intweight = q.reshape(K, N).numpy().astype(np.uint8)
    qweight = torch.from_numpy(pack_rows_int4(intweight))
    qzeros = torch.from_numpy(pack_cols_int4((zero - 1).numpy().astype(np.uint8)))
    g_idx = torch.arange(K, dtype=torch.int32) // g
    return qweight, qzeros, scale, g_idx


def _np(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


# --------------------------------------------------------------------------------------
# The product's packed weight stream, restated with plain numpy so that tests can
# (a) compare what b200_gptq_pack writes and (b) check, on a CPU, that the kernel's register
# arithmetic on those records is bit-identical to `dequantize` above.
# --------------------------------------------------------------------------------------
def packed_nkb(K: int) -> int:
    """128-wide k-blocks per tile, rounded up to a multiple of 8 when that adds at most 5 % (csrc/gemm_w4a16.cu w4_nkb)."""
    nkb = (K + 127) // 128
    up = (nkb + 7) // 8 * 8
    return up if (up - nkb) * 20 <= nkb else nkb


def unit_records(qweight, qzeros, scales, groupsize: int = 128, layout: int = 0, row_perm=None) -> np.ndarray:
    """-> uint32 [n_super, nkb, 2, 2048 + 128 * group_rows]: the unit records in stream order.
    words [4 chunks][128 features][4]: chunk c, word j = the 8 weights k = 128 kb + 32 c + 8 j .. + 7 of that feature with
    the nibbles ordered k0 k2 k4 k6 | k1 k3 k5 k7; meta [group_rows][128]: fp16 scale bits | (zero + 1) << 16.
    layout 0: super-tile s = feature tiles (2s, 2s + 1); layout 1: (s, s + n_tiles / 2).  row_perm: act-order row order."""
    q = unpack_rows_int4(_np(qweight)).astype(np.uint32)  # [K, N]
    if row_perm is not None:
        q = q[_np(row_perm).astype(np.int64)]
    K, N = q.shape
    if groupsize <= 0:
        groupsize = (K + 127) // 128 * 128
    z = unpack_cols_int4(_np(qzeros)).astype(np.uint32) + 1  # [G, N]
    sbits = _np(scales).view(np.uint16).astype(np.uint32)
    G = z.shape[0]
    gr = 1 if groupsize >= 128 else 128 // groupsize
    nkb = packed_nkb(K)
    n_tiles = (N + 127) // 128
    n_super = (n_tiles + 1) // 2
    half = n_tiles // 2 if layout == 1 else 0
    out = np.zeros((n_super, nkb, 2, 2048 + 128 * gr), dtype=np.uint32)
    out[..., 2048:] = 1 << 16  # padding: scale 0, zero 1
    pos = [(k & 1) * 4 + (k >> 1) for k in range(8)]
    for s in range(n_super):
        for r in range(2):
            tile = s + r * half if layout == 1 else 2 * s + r
            n0, n1 = tile * 128, min(N, tile * 128 + 128)
            if n0 >= N:
                continue
            for kb in range(nkb):
                words = np.zeros((4, 128, 4), dtype=np.uint32)
                for c in range(4):
                    for j in range(4):
                        k0 = kb * 128 + c * 32 + j * 8
                        if k0 >= K:
                            continue
                        w = np.zeros(n1 - n0, dtype=np.uint32)
                        for k in range(8):
                            w |= q[k0 + k, n0:n1] << np.uint32(4 * pos[k])
                        words[c, :n1 - n0, j] = w
                out[s, kb, r, :2048] = words.reshape(-1)
                for row in range(gr):
                    k0 = kb * 128 + row * (128 // gr)
                    if k0 < K:
                        g = min(k0 // groupsize, G - 1)
                        out[s, kb, r, 2048 + row * 128: 2048 + row * 128 + (n1 - n0)] = sbits[g, n0:n1] | (z[g, n0:n1] << 16)
    return out


def dequant_records_like_the_kernel(records: np.ndarray, K: int, N: int, groupsize: int = 128, layout: int = 0) -> torch.Tensor:
    """fp16 W[K, N] (rows in packed order) computed from unit records with the kernel's instruction sequence
    (csrc/gemm_w4a16.cu dequant_word): low nibbles 0x6400 | q = 1024 + q, high nibbles 0x5400 | q << 4 = 64 + q, an fp16
    add of -(1024 + zero) / -(64 + zero) (exact) and one fp16 multiply by the scale."""
    n_super, nkb, _, rec = records.shape
    gr = (rec - 2048) // 128
    n_tiles = (N + 127) // 128
    half = n_tiles // 2 if layout == 1 else 0
    W = np.zeros((nkb * 128, n_super * 256), dtype=np.float16)
    for s in range(n_super):
        for r in range(2):
            tile = s + r * half if layout == 1 else 2 * s + r
            for kb in range(nkb):
                words = records[s, kb, r, :2048].reshape(4, 128, 4)
                meta = records[s, kb, r, 2048:].reshape(gr, 128)
                for c in range(4):
                    m = meta[c * gr // 4]
                    scale = (m & 0xFFFF).astype(np.uint16).view(np.float16)
                    zp = (m >> 16).astype(np.uint16)
                    nz1024 = (np.uint16(0xE400) + zp).view(np.float16)          # -(1024 + zero)
                    nz64 = (np.uint16(0xD400) + (zp << 4)).view(np.float16)     # -(64 + zero)
                    for j in range(4):
                        w = words[c, :, j]
                        for half_word, shift in ((w, 0), (w >> 8, 4)):
                            lo_pair = [((half_word >> sh) & 0xF).astype(np.uint16) for sh in (0, 16)]          # (k, k+1)
                            hi_pair = [((half_word >> sh) & 0xF0).astype(np.uint16) for sh in (0, 16)]         # (k+2, k+3), q << 4
                            for e, qq in enumerate(lo_pair):
                                v = (np.uint16(0x6400) | qq).view(np.float16)
                                k = kb * 128 + c * 32 + j * 8 + shift + e
                                W[k, tile * 128:(tile + 1) * 128] = (v + nz1024) * scale
                            for e, qq in enumerate(hi_pair):
                                v = (np.uint16(0x5400) | qq).view(np.float16)
                                k = kb * 128 + c * 32 + j * 8 + shift + 2 + e
                                W[k, tile * 128:(tile + 1) * 128] = (v + nz64) * scale
    return torch.from_numpy(W[:K, :N].copy())
