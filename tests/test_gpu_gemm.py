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


def test_gptq_repack_roundtrip(ops):
    g = torch.Generator().manual_seed(11)
    q = torch.randint(-2 ** 31, 2 ** 31 - 1, (64, 256), generator=g, dtype=torch.int64).to(torch.int32)
    d = q.clone().to(DEV)
    ops.gptq_repack(d)
    rp = d.cpu()
    # nibble k of the original sits at position (k & 1) * 4 + (k >> 1)
    qu = q.to(torch.int64) & 0xFFFFFFFF
    ru = rp.to(torch.int64) & 0xFFFFFFFF
    for k in range(8):
        pos = (k & 1) * 4 + (k >> 1)
        assert torch.equal((qu >> (4 * k)) & 15, (ru >> (4 * pos)) & 15)
    ops.gptq_repack(d, inverse=True)
    assert torch.equal(d.cpu(), q)


W4_SHAPES = [(64, 4096, 4096, 128), (1, 256, 128, 128), (7, 384, 256, 64), (16, 128, 64, 32), (33, 2560, 2048, 128),
             (64, 4096, 11008, 128), (128, 512, 1024, 128), (300, 768, 512, -1)]


@pytest.mark.parametrize("T,N,K,gs", W4_SHAPES)
def test_gemm_w4a16(ops, T, N, K, gs):
    g = torch.Generator().manual_seed(T + N + K)
    w = torch.randn(N, K, generator=g) * 0.05
    qweight, qzeros, scales, g_idx = ogptq.quantize_rtn(w, gs)
    # random zeros, exercising the whole 0..15 stored range incl. the +1 -> 16 case
    qzeros = torch.randint(-2 ** 31, 2 ** 31 - 1, qzeros.shape, generator=g, dtype=torch.int64).to(torch.int32)
    x = torch.randn(T, K, generator=g).half()
    ref = ogptq.gemm_half_q_half(x, qweight, qzeros, scales, None, gs)
    qw_d = qweight.clone().to(DEV)
    ops.gptq_repack(qw_d)
    got = ops.gemm_w4a16(x.to(DEV), qw_d, qzeros.to(DEV), scales.to(DEV), gs)
    torch.cuda.synchronize()
    _close(got, ref, what=f"gemm_w4a16 {T}x{N}x{K} g{gs}")


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
    qw_d = qweight.clone().to(DEV)
    ops.gptq_repack(qw_d)
    got = ops.gemm_w4a16(x.to(DEV), qw_d, qzeros.to(DEV), scales.to(DEV), gs).cpu()
    exp = wd[ks]
    if not torch.equal(got, exp):
        bad = torch.nonzero(got != exp)
        raise AssertionError(f"dequant not bit exact: {bad.shape[0]} mismatches, first {bad[0].tolist()} "
                             f"(k={ks[bad[0][0]]}): got {got[tuple(bad[0])].item()} exp {exp[tuple(bad[0])].item()}; "
                             f"bad k rows {sorted(set(ks[i] for i in bad[:, 0].tolist()))[:20]}")
