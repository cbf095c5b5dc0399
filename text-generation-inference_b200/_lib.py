"""ctypes binding of libb200_tgis.so (the C ABI declared in include/b200_tgis.h)."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_void_p, c_char_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200_tgis.so")

_P, _I, _L, _F = c_void_p, c_int, c_int64, c_float

# name -> (restype, argtypes); must list every symbol of include/b200_tgis.h (tests/test_abi.py checks)
SIGNATURES = {
    "b200_abi_version": (_I, []),
    "b200_last_error": (c_char_p, []),
    "b200_rmsnorm_residual": (_I, [_P, _P, _P, _P, _P, _L, _L, _F, _P]),
    "b200_rope_kv_write_paged": (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P]),
    "b200_silu_mul": (_I, [_P, _P, _L, _L, _P]),
    "b200_embedding": (_I, [_P, _P, _P, _L, _L, _L, _L, _P]),
    "b200_argmax": (_I, [_P, _P, _L, _L, _L, _P]),
    "b200_attn_decode_workspace_bytes": (_L, [_I, _I, _I, _I]),
    "b200_attn_decode_paged": (_I, [_P, _L, _P, _P, _P, _L, _P, _P, _L, _P, _L, _I, _I, _I, _I, _I, _F, _P]),
    "b200_attn_prefill_varlen": (_I, [_P, _L, _P, _L, _P, _L, _P, _P, _L, _I, _I, _I, _I, _I, _F, _I, _P]),
    "b200_gemm_workspace_bytes": (_L, [_L, _L, _L]),
    "b200_gemm_f16": (_I, [_P, _P, _P, _P, _L, _L, _L, _P, _P]),
    "b200_gptq_repack": (_I, [_P, _L, _L, _I, _P]),
    "b200_gemm_w4a16": (_I, [_P, _P, _P, _P, _P, _P, _L, _L, _L, _I, _P, _P]),
}

_lib = None


class B200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Loads the library (building is `__graft_entry__.build()` / build.py's job).  Fails loudly when absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} is missing: build it with `python text-generation-inference_b200/build.py` "
                            "(there is no CPU fallback for the B200 hot path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().b200_last_error().decode()
        raise B200Error(f"{what} failed with status {status}: {msg}")
