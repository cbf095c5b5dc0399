"""GPU parity tests for the tcgen05 linears (fp16 and GPTQ int4) against the CPU oracle.

Tolerance: both sides accumulate exact fp16 x fp16 products in fp32 and round once to fp16, so they may differ
only by accumulation order: |err| <= 1e-3 * max|y| + 1 fp16 ulp.  The int4 dequantised weights are bit-identical
to the oracle's (one fp16 rounding of scale * (q - zero)), checked through a one-hot activation probe.
"""
import pytest
import torch

from oracle import gptq as ogptq

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    import tgis_b200  # noqa: F401
    from tgis_b200 import ops as _ops
    return _ops


def _close(got, ref, rel=1e-3, what=""):
    got = got.float().cpu()
    ref = ref.float().cpu()
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs()
    tol = rel * scale + 2.0 ** -10 * ref.abs()
    bad = err > tol
    if bad.any():
        idx = torch.nonzero(bad)[0].tolist()
        rows = sorted(set(torch.nonzero(bad)[:, 0].tolist()))[:8]
        cols = sorted(set(torch.nonzero(bad)[:, 1].tolist()))[:16]
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} off; max err {err.max().item():.4e} (scale {scale:.3e}); "
                             f"first bad {idx}: got {got[tuple(idx)].item()} ref {ref[tuple(idx)].item()}; bad rows {rows} cols {cols}")


F16_SHAPES = [(64, 4096, 4096), (1, 256, 128), (7, 384, 192), (16, 128, 64), (33, 2560, 2048), (64, 32000, 2048),
              (128, 512, 1024), (300, 768, 512), (64, 4096, 11008)]


@pytest.mark.parametrize("T,N,K", F16_SHAPES)
def test_gemm_f16(ops, T, N, K):
    g = torch.Generator().manual_seed(T + N + K)
    x = torch.randn(T, K, generator=g).half()
    w = (torch.randn(N, K, generator=g) * 0.05).half()
    ref = (x.float() @ w.float().t()).half()
    got = ops.gemm_f16(x.to(DEV), w.to(DEV))
    torch.cuda.synchronize()
    _close(got, ref, what=f"gemm_f16 {T}x{N}x{K}")
    # replay: split-K tile counters must have re-armed themselves
    got2 = ops.gemm_f16(x.to(DEV), w.to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(got.cpu(), got2.cpu()), "gemm_f16 must be deterministic across launches"


def test_gemm_f16_bias(ops):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(9, 256, generator=g).half()
    w = (torch.randn(384, 256, generator=g) * 0.05).half()
    b = torch.randn(384, generator=g).half()
    ref = (x.float() @ w.float().t() + b.float()).half()
    _close(ops.gemm_f16(x.to(DEV), w.to(DEV), b.to(DEV)), ref, what="gemm_f16 bias")


def test_gptq_pack_layout(ops):
    """b200_gptq_pack output == the documented unit-record layout, built here with plain indexing."""
    g = torch.Generator().manual_seed(11)
    K, N, gs = 320, 288, 64  # ragged: 3 k-blocks (last half empty), 3 feature tiles (third has 32 features) padded to 2 super-tiles, 2 meta rows
    qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 8, N), generator=g, dtype=torch.int64).to(torch.int32)
    qz = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // gs, N // 8), generator=g, dtype=torch.int64).to(torch.int32)
    sc = torch.randn(K // gs, N, generator=g).half()
    packed = ops.gptq_pack(qw.to(DEV), qz.to(DEV), sc.to(DEV), gs).cpu()
    nkb, nt, gr = 3, 4, 2
    # records are ordered (super-tile = pair of feature tiles, k-block, tile within the pair)
    rec = (packed.view(torch.int32).to(torch.int64).view(nt // 2, nkb, 2, (8192 + gr * 512) // 4) & 0xFFFFFFFF).permute(0, 2, 1, 3)
    rec = rec.reshape(nt, nkb, -1)
    q = ogptq.unpack_rows_int4(qw.numpy())                 # [K, N] nibbles
    z = ogptq.unpack_cols_int4(qz.numpy()).astype("int64") + 1  # [G, N]
    sbits = sc.view(torch.int16).to(torch.int64) & 0xFFFF
    for tile in range(nt):
        for kb in range(nkb):
            words = rec[tile, kb, :2048].view(4, 128, 4)
            meta = rec[tile, kb, 2048:].view(gr, 128)
            for m in (0, 1, 31, 32, 127):
                n = tile * 128 + m
                for c in range(4):
                    for j in range(4):
                        k0 = kb * 128 + c * 32 + j * 8
                        exp = 0
                        if n < N and k0 < K:
                            for k in range(8):
                                exp |= int(q[k0 + k, n]) << (4 * ((k & 1) * 4 + (k >> 1)))
                        assert int(words[c, m, j]) == exp, (tile, kb, c, m, j)
                for r in range(gr):
                    k0 = kb * 128 + r * 64
                    exp = 1 << 16
                    if n < N and k0 < K:
                        exp = int(sbits[k0 // gs, n]) | (int(z[k0 // gs, n]) << 16)
                    assert int(meta[r, m]) == exp, (tile, kb, r, m)


W4_SHAPES = [(64, 4096, 4096, 128), (1, 256, 128, 128), (7, 384, 256, 64), (16, 128, 64, 32), (33, 2560, 2048, 128),
             (64, 4096, 11008, 128), (128, 512, 1024, 128), (300, 768, 512, -1),
             # prefill sizes: persistent CTAs over (token tile, super-tile) items - fewer items than SMs, and several items per CTA
             (2100, 1280, 1024, 128), (4096, 2048, 512, 64), (5000, 4096, 256, 128)]


@pytest.mark.parametrize("T,N,K,gs", W4_SHAPES)
def test_gemm_w4a16(ops, T, N, K, gs):
    g = torch.Generator().manual_seed(T + N + K)
    w = torch.randn(N, K, generator=g) * 0.05
    qweight, qzeros, scales, g_idx = ogptq.quantize_rtn(w, gs)
    # random zeros, exercising the whole 0..15 stored range incl. the +1 -> 16 case
    qzeros = torch.randint(-2 ** 31, 2 ** 31 - 1, qzeros.shape, generator=g, dtype=torch.int64).to(torch.int32)
    x = torch.randn(T, K, generator=g).half()
    ref = ogptq.gemm_half_q_half(x, qweight, qzeros, scales, None, gs)
    packed = ops.gptq_pack(qweight.to(DEV), qzeros.to(DEV), scales.to(DEV), gs)
    got = ops.gemm_w4a16(x.to(DEV), packed, N, gs)
    torch.cuda.synchronize()
    _close(got, ref, what=f"gemm_w4a16 {T}x{N}x{K} g{gs}")


@pytest.mark.parametrize("T,I,K", [(64, 1024, 512), (5, 256, 256), (64, 11008, 4096)])
def test_gemm_w4a16_gate_up_layout_and_fused_silu(ops, T, I, K):
    """gate|up record order: the plain product is unchanged (up to the fp32 summation grouping of a different stream-K
    split), and the fused epilogue equals linear -> b200_silu_mul bit for bit (same arithmetic,
    flash_llama_modeling.py:332-335) and the oracle's SiLU * up."""
    g = torch.Generator().manual_seed(T + I + K)
    N, gs = 2 * I, 128
    w = torch.randn(N, K, generator=g) * 0.05
    qweight, qzeros, scales, _ = ogptq.quantize_rtn(w, gs)
    x = torch.randn(T, K, generator=g).half()
    qw, qz, sc, xd = qweight.to(DEV), qzeros.to(DEV), scales.to(DEV), x.to(DEV)
    plain = ops.gemm_w4a16(xd, ops.gptq_pack(qw, qz, sc, gs), N, gs)
    paired = ops.gptq_pack(qw, qz, sc, gs, layout=ops.W4_LAYOUT_GATE_UP)
    same = ops.gemm_w4a16(xd, paired, N, gs, layout=ops.W4_LAYOUT_GATE_UP)
    _close(same, plain.cpu(), what="gate|up layout vs plain layout")
    assert (same != plain).float().mean().item() < 0.05  # only fp16 rounding flips from the regrouped fp32 sums
    fused = ops.gemm_w4a16(xd, paired, N, gs, layout=ops.W4_LAYOUT_GATE_UP, silu_mul=True)
    unfused = ops.silu_mul(same)
    torch.cuda.synchronize()
    assert fused.shape == (T, I)
    assert torch.equal(fused, unfused), f"{(fused != unfused).sum().item()} elements differ"
    ref = ogptq.gemm_half_q_half(x, qweight, qzeros, scales, None, gs).float()
    ref = (torch.nn.functional.silu(ref[:, :I].half().float()).half() * ref[:, I:].half()).half()
    _close(fused, ref, what=f"fused silu*up {T}x{I}x{K}")


def test_gemm_w4a16_act_order(ops):
    """Act-order checkpoint (shuffled g_idx, exllamav2.py:31-48): Ex4bitLinearV2 packs the rows in group order and gathers
    the activations; result == oracle dequant with g_idx (quant_linear.py:184-192 `g = g_idx[k]`)."""
    from tgis_b200.utils.gptq.exllamav2 import Ex4bitLinearV2
    g = torch.Generator().manual_seed(5)
    T, N, K, gs = 17, 384, 512, 128
    w = torch.randn(N, K, generator=g) * 0.05
    qweight, qzeros, scales, _ = ogptq.quantize_rtn(w, gs)
    # a random assignment of rows to groups with exactly gs rows each; the row's nibbles keep their values
    g_idx = (torch.randperm(K, generator=g) // gs).to(torch.int32)
    x = torch.randn(T, K, generator=g).half()
    ref = ogptq.gemm_half_q_half(x, qweight, qzeros, scales, g_idx, gs)
    lin = Ex4bitLinearV2(qweight.to(DEV), qzeros.to(DEV), scales.to(DEV), g_idx.to(DEV), None, 4, gs)
    assert lin.q_perm is not None
    got = lin(x.to(DEV))
    torch.cuda.synchronize()
    _close(got, ref, what="act-order gemm")
    # one-hot probes: row k of the dequantised matrix, bit for bit
    wd = ogptq.dequantize(qweight, qzeros, scales, g_idx, gs)
    ks = [0, 1, 127, 128, 300, 511]
    xo = torch.zeros(len(ks), K, dtype=torch.float16)
    for i, k in enumerate(ks):
        xo[i, k] = 1.0
    assert torch.equal(lin(xo.to(DEV)).cpu(), wd[ks])


def test_gemm_w4a16_dequant_bit_exact(ops):
    """x = one-hot rows -> y[t] is row k_t of the dequantised matrix: must equal the oracle's fp16 W bit for bit."""
    g = torch.Generator().manual_seed(77)
    N, K, gs = 256, 256, 128
    w = torch.randn(N, K, generator=g) * 0.05
    qweight, qzeros, scales, _ = ogptq.quantize_rtn(w, gs)
    qzeros = torch.randint(-2 ** 31, 2 ** 31 - 1, qzeros.shape, generator=g, dtype=torch.int64).to(torch.int32)
    wd = ogptq.dequantize(qweight, qzeros, scales, None, gs)  # [K, N] fp16
    ks = list(range(0, 64)) + [64, 127, 128, 200, 255]
    x = torch.zeros(len(ks), K, dtype=torch.float16)
    for i, k in enumerate(ks):
        x[i, k] = 1.0
    packed = ops.gptq_pack(qweight.to(DEV), qzeros.to(DEV), scales.to(DEV), gs)
    got = ops.gemm_w4a16(x.to(DEV), packed, N, gs).cpu()
    exp = wd[ks]
    if not torch.equal(got, exp):
        bad = torch.nonzero(got != exp)
        raise AssertionError(f"dequant not bit exact: {bad.shape[0]} mismatches, first {bad[0].tolist()} "
                             f"(k={ks[bad[0][0]]}): got {got[tuple(bad[0])].item()} exp {exp[tuple(bad[0])].item()}; "
                             f"bad k rows {sorted(set(ks[i] for i in bad[:, 0].tolist()))[:20]}")


# per-rank linears of BASELINE configs not covered above: Llama-3-70B int4 at NUM_SHARD = 8 (hidden 8192, 8 + 2 heads of 128,
# intermediate 28672 / 8) and Llama-3-8B at bs = 64 / 128.
W4_MODEL_SHAPES = [(128, 1280, 8192, 128), (128, 8192, 1024, 128), (128, 7168, 8192, 128), (128, 8192, 3584, 128),
                   (64, 6144, 4096, 128), (64, 28672, 4096, 128), (64, 4096, 14336, 128), (1, 8192, 3584, 128)]


@pytest.mark.parametrize("T,N,K,gs", W4_MODEL_SHAPES)
def test_gemm_w4a16_model_shapes(ops, T, N, K, gs):
    test_gemm_w4a16(ops, T, N, K, gs)
