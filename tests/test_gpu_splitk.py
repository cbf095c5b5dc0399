"""GPU parity tests of the deferred split-K reduction (include/b200_tgis.h "B200SplitK"): a *_deferred GEMM leaves fp32
partials in the workspace and the consumer kernel that follows sums them.

Two kinds of checks:
  * the summed output (b200_splitk_reduce) against the CPU oracle, same tolerance as the fused GEMMs (fp32 accumulation in a
    different grouping, one fp16 rounding): |err| <= 1e-3 * max|y| + 1 fp16 ulp;
  * every fused consumer against "materialise, then the plain kernel": bit for bit, because both add the same partials in the
    same order and round once (residual + RMSNorm, RoPE + KV write, SiLU * up).
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
    got, ref = got.float().cpu(), ref.float().cpu()
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs()
    bad = err > rel * scale + 2.0 ** -10 * ref.abs()
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} off, max err {err.max().item():.3e} (scale {scale:.3e})"


def _w4(T, N, K, gs, seed, bias=False):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(N, K, generator=g) * 0.05
    qweight, qzeros, scales, _ = ogptq.quantize_rtn(w, gs)
    qzeros = torch.randint(-2 ** 31, 2 ** 31 - 1, qzeros.shape, generator=g, dtype=torch.int64).to(torch.int32)
    x = torch.randn(T, K, generator=g).half()
    b = torch.randn(N, generator=g).half() if bias else None
    ref = ogptq.gemm_half_q_half(x, qweight, qzeros, scales, None, gs)
    if bias:
        ref = (ref.float() + b.float()).half()
    return x, qweight, qzeros, scales, b, ref


# decode shapes of the served models (per rank) + ragged ones: T, N, K, groupsize
W4_DEFERRED = [(64, 4096, 4096, 128), (64, 6144, 4096, 128), (64, 4096, 14336, 128), (64, 12288, 4096, 128), (64, 4096, 11008, 128),
               (1, 256, 128, 128), (7, 384, 256, 64), (16, 128, 64, 32), (33, 2560, 2048, 128), (128, 512, 1024, 128),
               (128, 8192, 1024, 128), (100, 1280, 8192, 128)]


@pytest.mark.parametrize("T,N,K,gs", W4_DEFERRED)
def test_w4_deferred_reduce_matches_oracle_and_is_deterministic(ops, T, N, K, gs):
    x, qw, qz, sc, b, ref = _w4(T, N, K, gs, T + N + K, bias=(N == 384))
    packed = ops.gptq_pack(qw.to(DEV), qz.to(DEV), sc.to(DEV), gs)
    bd = b.to(DEV) if b is not None else None
    parts = ops.gemm_w4a16_deferred(x.to(DEV), packed, N, gs, bias=bd)
    y = ops.splitk_reduce(parts)
    torch.cuda.synchronize()
    _close(y, ref, what=f"w4 deferred {T}x{N}x{K} g{gs}")
    y2 = ops.splitk_reduce(ops.gemm_w4a16_deferred(x.to(DEV), packed, N, gs, bias=bd))
    torch.cuda.synchronize()
    assert torch.equal(y, y2)
    # the non-deferred kernel interleaved on the same workspace still works (counters untouched by the deferred launches)
    _close(ops.gemm_w4a16(x.to(DEV), packed, N, gs, bias=bd), ref, what="plain after deferred")


@pytest.mark.parametrize("T,N,K", [(64, 4096, 4096), (64, 1536, 4096), (64, 4096, 512), (64, 4096, 1376), (1, 256, 128), (7, 384, 192),
                                   (33, 2560, 2048), (128, 512, 1024), (200, 768, 512)])
def test_f16_deferred_reduce_matches_reference(ops, T, N, K):
    g = torch.Generator().manual_seed(T + N + K)
    x = torch.randn(T, K, generator=g).half()
    w = (torch.randn(N, K, generator=g) * 0.05).half()
    b = torch.randn(N, generator=g).half() if N == 384 else None
    ref = x.float() @ w.float().t()
    if b is not None:
        ref = ref + b.float()
    parts = ops.gemm_f16_deferred(x.to(DEV), w.to(DEV), bias=b.to(DEV) if b is not None else None)
    y = ops.splitk_reduce(parts)
    torch.cuda.synchronize()
    _close(y, ref.half(), what=f"f16 deferred {T}x{N}x{K}")
    _close(ops.gemm_f16(x.to(DEV), w.to(DEV), b.to(DEV) if b is not None else None), ref.half(), what="plain after deferred")


@pytest.mark.parametrize("quant", [True, False])
@pytest.mark.parametrize("T,H,K", [(64, 4096, 4096), (64, 4096, 1376), (5, 256, 512), (128, 2048, 5632)])
def test_rmsnorm_consumes_partials_bit_exactly(ops, quant, T, H, K):
    g = torch.Generator().manual_seed(T + H + K)
    res = torch.randn(T, H, generator=g).half().to(DEV)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).half().to(DEV)
    if quant:
        x, qw, qz, sc, _, _ = _w4(T, H, K, 128 if K % 128 == 0 else 32, 3)
        packed = ops.gptq_pack(qw.to(DEV), qz.to(DEV), sc.to(DEV), 128 if K % 128 == 0 else 32)
        parts = ops.gemm_w4a16_deferred(x.to(DEV), packed, H, 128 if K % 128 == 0 else 32)
    else:
        x = torch.randn(T, K, generator=g).half()
        w = (torch.randn(H, K, generator=g) * 0.05).half()
        parts = ops.gemm_f16_deferred(x.to(DEV), w.to(DEV))
    h = ops.splitk_reduce(parts)
    n_ref, r_ref = ops.rmsnorm_residual(h, res, gamma, 1e-5)
    n_got, r_got = ops.rmsnorm_residual_splitk(parts, res, gamma, 1e-5)
    torch.cuda.synchronize()
    assert torch.equal(r_got, r_ref), f"residual_out differs: {(r_got != r_ref).sum().item()} elements"
    assert torch.equal(n_got, n_ref), f"normed differs: {(n_got != n_ref).sum().item()} elements"


@pytest.mark.parametrize("h,kv,d", [(32, 32, 128), (32, 8, 128), (4, 1, 128), (32, 4, 64)])
def test_rope_consumes_partials_bit_exactly(ops, h, kv, d):
    T, K = 37, 512
    N = (h + 2 * kv) * d
    g = torch.Generator().manual_seed(h + kv + d)
    x, qw, qz, sc, b, _ = _w4(T, N, K, 128, 9, bias=True)
    packed = ops.gptq_pack(qw.to(DEV), qz.to(DEV), sc.to(DEV), 128)
    cos = torch.randn(64, d // 2, generator=g).half().to(DEV)
    sin = torch.randn(64, d // 2, generator=g).half().to(DEV)
    pos = torch.randint(0, 64, (T,), generator=g).to(DEV)
    slots = torch.randperm(8 * 16, generator=g)[:T].to(torch.int64)
    slots[3] = -1  # a padding row: rotated in qkv, not written to the pool
    slots = slots.to(DEV)
    parts = ops.gemm_w4a16_deferred(x.to(DEV), packed, N, 128, bias=b.to(DEV))
    qkv_ref = ops.splitk_reduce(parts)
    k1, v1 = ops.kv_pool_alloc(8, kv, d, DEV)
    k2, v2 = ops.kv_pool_alloc(8, kv, d, DEV)
    ops.rope_kv_write_paged(qkv_ref, cos, sin, pos, slots, k1, v1, h, kv, d)
    qkv_got = ops.rope_kv_write_paged_splitk(parts, cos, sin, pos, slots, k2, v2, h, kv, d)
    torch.cuda.synchronize()
    assert torch.equal(qkv_got, qkv_ref)
    assert torch.equal(k1, k2) and torch.equal(v1, v2)


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("T,I,K", [(64, 1024, 512), (5, 256, 256), (64, 14336, 4096), (64, 11008, 4096)])
def test_silu_consumes_partials_bit_exactly(ops, layout, T, I, K):
    N = 2 * I
    x, qw, qz, sc, b, ref = _w4(T, N, K, 128, 4, bias=(I == 256))
    packed = ops.gptq_pack(qw.to(DEV), qz.to(DEV), sc.to(DEV), 128, layout=layout)
    bd = b.to(DEV) if b is not None else None
    parts = ops.gemm_w4a16_deferred(x.to(DEV), packed, N, 128, bias=bd, layout=layout)
    gu = ops.splitk_reduce(parts)  # natural [gate | up] column order for either record layout
    _close(gu, ref, what="gate_up deferred")
    got = ops.splitk_silu_mul(parts)
    torch.cuda.synchronize()
    assert torch.equal(got, ops.silu_mul(gu))


def test_silu_consumes_f16_partials(ops):
    T, I, K = 64, 2752, 4096
    g = torch.Generator().manual_seed(1)
    x = torch.randn(T, K, generator=g).half().to(DEV)
    w = (torch.randn(2 * I, K, generator=g) * 0.05).half().to(DEV)
    parts = ops.gemm_f16_deferred(x, w)
    assert torch.equal(ops.splitk_silu_mul(parts), ops.silu_mul(ops.splitk_reduce(parts)))
