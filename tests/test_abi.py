"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/b200_tgis.h declares; the ctypes table matches the header; host-only entry points work."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200_tgis.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge._load_build_module().build()  # no-op when up to date; nvcc cross-compiles sm_100a without a GPU
    import tgis_b200  # noqa: F401
    from tgis_b200 import _lib
    return _lib


def _header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = _header_symbols()
    for s in ["b200_rmsnorm_residual", "b200_rope_kv_write_paged", "b200_attn_prefill_varlen", "b200_attn_decode_paged",
              "b200_gptq_packed_bytes", "b200_gptq_pack", "b200_gemm_w4a16", "b200_gemm_f16", "b200_silu_mul", "b200_argmax", "b200_embedding",
              "b200_kv_alloc_create", "b200_kv_alloc_take", "b200_kv_alloc_release", "b200_llama_step",
              "b200_gemm_w4a16_deferred", "b200_gemm_f16_deferred", "b200_rmsnorm_residual_splitk", "b200_rope_kv_write_paged_splitk",
              "b200_splitk_silu_mul", "b200_splitk_reduce", "b200_p2p_allreduce_rmsnorm", "b200_p2p_argmax"]:
        assert s in syms, s


def test_library_exports_every_header_symbol(lib):
    handle = lib.load()
    missing = [s for s in _header_symbols() if not hasattr(handle, s)]
    assert not missing, f"libb200_tgis.so lacks {missing}"


def test_ctypes_table_covers_the_header(lib):
    syms = set(_header_symbols())
    table = set(lib.SIGNATURES)
    assert syms <= table, f"_lib.SIGNATURES lacks {sorted(syms - table)}"
    assert table <= syms, f"_lib.SIGNATURES binds undeclared {sorted(table - syms)}"


def test_struct_layouts_match_the_header(lib, tmp_path):
    """compile the header with gcc and compare sizeof / offsetof of every struct field with the ctypes mirror"""
    import subprocess
    structs = {"B200Linear": lib.B200Linear, "B200LlamaLayer": lib.B200LlamaLayer, "B200LlamaWeights": lib.B200LlamaWeights,
               "B200LlamaStep": lib.B200LlamaStep, "B200SplitK": lib.B200SplitK, "B200ChooserParams": lib.B200ChooserParams}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{name}.{fname} %zu\\n", offsetof({name}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(out[name]) == ctypes.sizeof(cls), name
        for fname, _ in cls._fields_:
            assert int(out[f"{name}.{fname}"]) == getattr(cls, fname).offset, f"{name}.{fname}"


def test_host_only_entry_points_work_without_a_gpu(lib):
    h = lib.load()
    assert h.b200_abi_version() == 1
    assert h.b200_launch_count() >= 0
    assert h.b200_attn_decode_workspace_bytes(64, 32, 128, 2048) == 64 * 32 * 16 * 130 * 4  # sized for 128-token chunks
    assert h.b200_gemm_workspace_bytes(64, 4096, 4096) >= 64 * 1024
    assert h.b200_gemm_workspace_bytes_max(4096, 4096) >= h.b200_gemm_workspace_bytes(64, 4096, 4096)
    a = h.b200_kv_alloc_create(8)
    buf = (ctypes.c_int32 * 8)()
    assert h.b200_kv_alloc_take(a, 5, buf) == 0 and list(buf)[:5] == [0, 1, 2, 3, 4]
    assert h.b200_kv_alloc_num_free(a) == 3
    assert h.b200_kv_alloc_take(a, 4, buf) == -4 and b"out of KV cache blocks" in h.b200_last_error()
    rel = (ctypes.c_int32 * 2)(1, 3)
    assert h.b200_kv_alloc_release(a, rel, 2) == 0 and h.b200_kv_alloc_num_free(a) == 5
    bad = (ctypes.c_int32 * 1)(99)
    assert h.b200_kv_alloc_release(a, bad, 1) == -1
    # a double free is refused while other blocks are still out, and leaves the free list untouched (block 1 is free, 0 is not)
    twice = (ctypes.c_int32 * 2)(0, 1)
    assert h.b200_kv_alloc_release(a, twice, 2) == -1 and b"double free" in h.b200_last_error()
    assert h.b200_kv_alloc_num_free(a) == 5
    same = (ctypes.c_int32 * 2)(2, 2)
    assert h.b200_kv_alloc_release(a, same, 2) == -1 and h.b200_kv_alloc_num_free(a) == 5
    assert h.b200_kv_alloc_release(a, (ctypes.c_int32 * 3)(0, 2, 4), 3) == 0 and h.b200_kv_alloc_num_free(a) == 8
    h.b200_kv_alloc_destroy(a)


def test_ops_fail_loudly_without_cuda(lib):
    """no CPU fallback: a CPU tensor is rejected before anything runs"""
    import torch
    from tgis_b200 import ops
    with pytest.raises(lib.B200Error):
        ops.rmsnorm_residual(torch.zeros(2, 64, dtype=torch.float16), None, torch.ones(64, dtype=torch.float16), 1e-5)
    with pytest.raises(lib.B200Error):
        ops.gemm_f16(torch.zeros(2, 64, dtype=torch.float16), torch.zeros(8, 64, dtype=torch.float16))


def test_gptq_packed_format_sizes(lib):
    """host-side size rules of the int4 weight stream: 8 KB of words + 512 B of meta per group row for every
    (128-feature tile, 128-wide k-block); feature tiles padded to pairs; k-blocks padded to a multiple of 8 when <= 5 %"""
    h = lib.load()
    rec = 8192 + 512
    assert h.b200_gptq_packed_bytes(4096, 4096, 128) == 32 * 32 * rec
    assert h.b200_gptq_packed_bytes(4096, 22016, 128) == 172 * 32 * rec
    assert h.b200_gptq_packed_bytes(11008, 4096, 128) == 32 * 88 * rec          # 86 k-blocks -> 88
    assert h.b200_gptq_packed_bytes(320, 288, 64) == 4 * 3 * (8192 + 2 * 512)   # 3 tiles -> 2 pairs; 2 meta rows
    assert h.b200_gptq_packed_bytes(4096, 4096, -1) == 32 * 32 * rec            # one group: one meta row per record
    assert h.b200_gptq_packed_bytes(4096, 4100, 128) < 0                        # N % 32 != 0 (exllamav2.py:118-119)
    assert h.b200_gptq_packed_bytes(4096, 4096, 48) < 0                         # groupsize must divide / be divided by 128
    assert b"groupsize" in h.b200_last_error()


def test_p2p_allreduce_argument_checks(lib):
    """the peer-memory all-reduce (experimental) validates before it touches CUDA, and has no CPU form"""
    import torch
    h = lib.load()
    assert h.b200_p2p_handle_bytes() == 64  # cudaIpcMemHandle_t
    ctx = ctypes.c_void_p()
    handle = (ctypes.c_ubyte * 64)()
    assert h.b200_p2p_create(1 << 20, 1, 0, ctypes.byref(ctx), handle) != 0   # a group of one has nothing to reduce
    assert h.b200_p2p_create(1 << 20, 9, 0, ctypes.byref(ctx), handle) != 0   # one NVSwitch domain: <= 8 ranks
    assert h.b200_p2p_create(1 << 20, 2, 2, ctypes.byref(ctx), handle) != 0
    assert h.b200_p2p_allreduce_f16(None, None, 8, None) != 0
    assert h.b200_p2p_max_bytes(None) == 0
    if not torch.cuda.is_available():
        assert h.b200_p2p_create(1 << 20, 2, 0, ctypes.byref(ctx), handle) != 0  # no device: fails loudly
    # default: NCCL / gloo stays the transport unless B200_P2P_ALLREDUCE=1
    from tgis_b200.utils.dist import FakeGroup
    from tgis_b200.utils.p2p import LayerBoundaryAllReduce
    reduce = LayerBoundaryAllReduce(FakeGroup(0, 1))
    x = torch.ones(8, dtype=torch.float16)
    assert not reduce.uses_peer_memory and reduce(x) is x
