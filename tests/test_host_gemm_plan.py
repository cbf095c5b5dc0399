"""Host-side stream-K planner of the two GEMMs (csrc/gemm_f16.cu plan_f16, csrc/gemm_w4a16.cu plan_w4), no GPU needed.

The kernels' fix-up spins on the other contributors of a tile, so a plan is only safe when every CTA is resident at once,
every (tile, k-block) unit belongs to exactly one CTA, and the workspace the caller was told to allocate
(`b200_gemm_workspace_bytes`) holds one fp32 partial per contributor slot.  Swept over every linear of the model families
this path serves (Llama-2/3, TinyLlama, GPT-NeoX) at tensor-parallel degrees 1, 2, 4, 8 and decode batch sizes 1 ... 256.
"""
import ctypes

import pytest

SMS = 148
TILE = 128  # feature rows per tile; k-blocks are 64 wide for fp16 weights and 128 for int4 (checked via nkb below)

# (hidden, intermediate, heads, kv heads, vocab)
FAMILIES = {
    "tinyllama-1.1b": (2048, 5632, 32, 4, 32000),
    "llama-2-7b": (4096, 11008, 32, 32, 32000),
    "llama-2-13b": (5120, 13824, 40, 40, 32000),
    "llama-3-8b": (4096, 14336, 32, 8, 128256),
    "llama-3-70b": (8192, 28672, 64, 8, 128256),
    "gpt-neox-20b": (6144, 24576, 64, 64, 50432),
}


def _linears(name, world):
    h, inter, heads, kv, vocab = FAMILIES[name]
    d = h // heads
    if heads % world or (kv % world and kv >= world):
        return []
    kv_local = max(kv // world, 1)
    shapes = {
        "qkv": ((heads // world + 2 * kv_local) * d, h),
        "o": (h, h // world),
        "gate_up": (2 * inter // world, h),
        "down": (h, inter // world),
        "head": (vocab // world, h),
    }
    return [(k, n, kk) for k, (n, kk) in shapes.items()]


@pytest.fixture(scope="module")
def lib():
    import tgis_b200  # noqa: F401
    from tgis_b200 import _lib
    return _lib.load()


def _plan(lib, kind, T, N, K, sms=SMS):
    out = (ctypes.c_int32 * 8)()
    assert lib.b200_debug_gemm_plan(kind, T, N, K, sms, out) == 0, lib.b200_last_error()
    keys = ("TN", "nkb", "tiles", "tiles_t", "units_per_cta", "ctas", "contrib", "r")
    return dict(zip(keys, out))


def _contributors(nkb, units_per_cta, tiles):
    """Brute force: how many CTAs touch each tile when consecutive CTAs take `units_per_cta` consecutive (tile, k-block) units."""
    total = tiles * nkb
    worst = 0
    for t in range(tiles):
        first_cta = (t * nkb) // units_per_cta
        last_cta = ((t + 1) * nkb - 1) // units_per_cta
        worst = max(worst, last_cta - first_cta + 1)
    return worst, (total + units_per_cta - 1) // units_per_cta


@pytest.mark.parametrize("kind", [0, 1], ids=["fp16", "int4"])
@pytest.mark.parametrize("family", list(FAMILIES))
def test_plans_are_safe(lib, kind, family):
    checked = 0
    for world in (1, 2, 4, 8):
        for which, N, K in _linears(family, world):
            if kind == 1 and (which == "head" or N % 32 or K % 32):
                continue  # GPTQ never quantizes the head (layers.py:236-237)
            for T in (1, 7, 16, 17, 32, 48, 64, 65, 128, 129, 256, 300):
                p = _plan(lib, kind, T, N, K)
                what = f"{family} tp{world} {which} N={N} K={K} T={T}: {p}"
                assert p["TN"] in (16, 32, 64, 128, 256) and p["TN"] * p["tiles_t"] >= T > p["TN"] * (p["tiles_t"] - 1), what
                assert p["tiles"] == -(-N // (TILE * p["r"])), what
                assert p["nkb"] * (64 if kind == 0 else 128) >= K, what
                worst, ctas = _contributors(p["nkb"], p["units_per_cta"], p["tiles"])
                assert p["ctas"] == ctas, what
                if p["units_per_cta"] % p["nkb"] == 0:
                    assert p["contrib"] == 1 and worst == 1, what  # whole tiles: no fix-up, any grid size
                else:
                    assert p["tiles_t"] == 1, what                  # stream-K only with a single token tile
                    assert p["ctas"] <= SMS, what                   # contributors spin on each other: all resident
                    assert worst <= p["contrib"], what              # a partial slot for every contributor
                    need = p["tiles"] * p["contrib"] * p["r"] * p["TN"] * TILE * 4
                    assert lib.b200_gemm_workspace_bytes(T, N, K) >= need, what
                assert lib.b200_gemm_workspace_bytes_max(N, K) >= lib.b200_gemm_workspace_bytes(T, N, K), what
                checked += 1
    assert checked > 0


def test_int4_decode_plans_fill_the_gpu(lib):
    """bs = 64 (the measured configuration): every linear of Llama-2-7B / Llama-3-8B keeps at least 60 % of the SMs busy, and
    a short-K projection (o_proj) gets an aligned cut: no CTA straddles two super-tiles."""
    for N, K in [(4096, 4096), (12288, 4096), (22016, 4096), (4096, 11008), (6144, 4096), (28672, 4096), (4096, 14336)]:
        p = _plan(lib, 1, 64, N, K)
        assert 0.6 * SMS <= p["ctas"] <= SMS, (N, K, p)
    p = _plan(lib, 1, 64, 4096, 4096)
    assert p["nkb"] % p["units_per_cta"] == 0, p


def test_bad_arguments(lib):
    out = (ctypes.c_int32 * 8)()
    assert lib.b200_debug_gemm_plan(2, 1, 128, 128, SMS, out) != 0
    assert lib.b200_debug_gemm_plan(0, 0, 128, 128, SMS, out) != 0
