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
    "b200_cuda_peek_error": (c_char_p, []),
    "b200_debug_w4_trace": (None, [_P]),
    "b200_debug_w4_flags": (None, [_I]),
    "b200_debug_step_trace": (None, [_P, _I]),
    "b200_debug_step_trace_begin": (None, [_P]),
    "b200_debug_step_trace_names": (_I, [c_char_p, _L]),
    "b200_debug_gemm_plan": (_I, [_I, _L, _L, _L, _I, _P]),
    "b200_rmsnorm_residual": (_I, [_P, _P, _P, _P, _P, _L, _L, _F, _P]),
    "b200_rope_kv_write_paged": (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P]),
    "b200_rope_kv_write_paged_ex": (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P]),
    "b200_layernorm_residual": (_I, [_P, _P, _P, _P, _P, _P, _L, _L, _F, _P]),
    "b200_gelu": (_I, [_P, _P, _L, _I, _P]),
    "b200_masked_softmax": (_I, [_P, _P, _P, _L, _L, _I, _P]),
    "b200_silu_mul": (_I, [_P, _P, _L, _L, _P]),
    "b200_embedding": (_I, [_P, _P, _P, _L, _L, _L, _L, _P]),
    "b200_argmax": (_I, [_P, _P, _L, _L, _L, _P, _P]),
    "b200_launch_count": (_L, []),
    "b200_timing_create": (_P, [_I]),
    "b200_timing_destroy": (None, [_P]),
    "b200_timing_attach": (None, [_P, _I]),
    "b200_timing_collect": (_I, [_P, ctypes.POINTER(c_float)]),
    "b200_attn_decode_workspace_bytes": (_L, [_I, _I, _I, _I]),
    "b200_attn_decode_paged": (_I, [_P, _L, _P, _P, _P, _L, _P, _P, _L, _P, _L, _I, _I, _I, _I, _I, _F, _P]),
    "b200_attn_prefill_varlen": (_I, [_P, _L, _P, _L, _P, _L, _P, _P, _L, _I, _I, _I, _I, _I, _F, _I, _P]),
    "b200_attn_prefill_paged": (_I, [_P, _L, _L, _P, _P, _L, _P, _L, _P, _P, _P, _L, _I, _I, _I, _I, _I, _F, _P]),
    "b200_gemm_workspace_bytes": (_L, [_L, _L, _L]),
    "b200_gemm_workspace_bytes_max": (_L, [_L, _L]),
    "b200_gemm_f16": (_I, [_P, _P, _P, _P, _L, _L, _L, _P, _P]),
    "b200_p2p_handle_bytes": (_I, []),
    "b200_p2p_create": (_I, [_L, _I, _I, ctypes.POINTER(_P), _P]),
    "b200_p2p_connect": (_I, [_P, _P]),
    "b200_p2p_max_bytes": (_L, [_P]),
    "b200_p2p_allreduce_f16": (_I, [_P, _P, _L, _P]),
    "b200_p2p_destroy": (None, [_P]),
    "b200_gptq_packed_bytes": (_L, [_L, _L, _I]),
    "b200_gptq_pack": (_I, [_P, _P, _P, _P, _L, _L, _I, _P]),
    "b200_gptq_pack_ex": (_I, [_P, _P, _P, _P, _P, _L, _L, _I, _I, _P]),
    "b200_permute_columns": (_I, [_P, _P, _P, _L, _L, _P]),
    "b200_gemm_w4a16": (_I, [_P, _P, _P, _P, _L, _L, _L, _I, _P, _P]),
    "b200_gemm_w4a16_ex": (_I, [_P, _P, _P, _P, _L, _L, _L, _I, _I, _I, _P, _P]),
}



class B200SplitK(ctypes.Structure):
    _fields_ = [("partial", _P), ("bias", _P), ("tiles_per_unit", ctypes.c_int32), ("tn", ctypes.c_int32), ("nkb", ctypes.c_int32),
                ("units_per_cta", ctypes.c_int32), ("max_contrib", ctypes.c_int32), ("half_tiles", ctypes.c_int32),
                ("N", ctypes.c_int32), ("T", ctypes.c_int32)]


class B200ChooserParams(ctypes.Structure):
    _fields_ = [("logits", _P), ("warped_scratch", _P), ("ld", _L), ("V", _L), ("B", ctypes.c_int32), ("history_len_bias", ctypes.c_int32),
                ("temperature", _P), ("top_k", _P), ("top_p", _P), ("rep_penalty", _P), ("history", _P), ("history_stride", _L),
                ("position_ids", _P), ("rep_exclude_id", _L), ("banned_ids", _P), ("length_penalty_factor", _P), ("eos_id", _L),
                ("seeds", _P), ("counters", _P), ("next_ids", _P), ("logprobs", _P), ("ranks", _P)]


class B200Linear(ctypes.Structure):
    _fields_ = [("weight", _P), ("qweight", _P), ("perm", _P), ("_unused", _P), ("bias", _P), ("N", _L), ("K", _L),
                ("groupsize", ctypes.c_int32), ("layout", ctypes.c_int32)]


class B200LlamaLayer(ctypes.Structure):
    _fields_ = [("input_ln", _P), ("post_ln", _P), ("qkv", B200Linear), ("o", B200Linear), ("gate_up", B200Linear),
                ("down", B200Linear)]


class B200LlamaWeights(ctypes.Structure):
    _fields_ = [("n_layers", ctypes.c_int32), ("hidden_size", ctypes.c_int32), ("n_heads", ctypes.c_int32),
                ("n_kv_heads", ctypes.c_int32), ("head_dim", ctypes.c_int32), ("tp_size", ctypes.c_int32),
                ("tp_rank", ctypes.c_int32), ("_pad", ctypes.c_int32), ("rms_eps", _F), ("softmax_scale", _F),
                ("layers", ctypes.POINTER(B200LlamaLayer)), ("embed", _P), ("vocab_start", _L), ("vocab_rows", _L),
                ("final_norm", _P), ("lm_head", _P), ("vocab_rows_head", _L), ("rope_cos", _P), ("rope_sin", _P)]


class B200LlamaStep(ctypes.Structure):
    _fields_ = [("T", _L), ("B", ctypes.c_int32), ("is_prefill", ctypes.c_int32), ("max_s", ctypes.c_int32),
                ("_pad", ctypes.c_int32), ("input_ids", _P), ("position_ids", _P), ("slot_mapping", _P), ("cu_seqlens", _P),
                ("block_table", _P), ("block_table_stride", _L), ("context_lens", _P), ("kv_pool", _P),
                ("kv_layer_stride_bytes", _L), ("kv_v_offset_bytes", _L), ("hidden", _P), ("residual", _P), ("normed", _P),
                ("qkv", _P), ("attn_out", _P), ("gate_up", _P), ("act", _P), ("perm_x", _P), ("attn_ws", _P), ("attn_ws_bytes", _L),
                ("gemm_ws", _P), ("head_rows", _P), ("n_head_rows", _L), ("head_in", _P), ("logits", _P), ("next_ids", _P), ("banned_ids", _P),
                ("defer_splitk", ctypes.c_int32), ("_pad2", ctypes.c_int32), ("p2p_norm", _P), ("p2p_argmax", _P),
                ("kv_num_blocks", _L), ("max_q", ctypes.c_int32), ("_pad3", ctypes.c_int32)]


_WP, _SP, _KP = ctypes.POINTER(B200LlamaWeights), ctypes.POINTER(B200LlamaStep), ctypes.POINTER(B200SplitK)
SIGNATURES.update({
    "b200_choose_tokens": (_I, [ctypes.POINTER(B200ChooserParams), _P]),
    "b200_gemm_w4a16_deferred": (_I, [_P, _P, _P, _L, _L, _L, _I, _I, _P, _KP, _P]),
    "b200_gemm_f16_deferred": (_I, [_P, _P, _P, _L, _L, _L, _P, _KP, _P]),
    "b200_splitk_reduce": (_I, [_KP, _P, _P]),
    "b200_rmsnorm_residual_splitk": (_I, [_KP, _P, _P, _P, _P, _F, _P]),
    "b200_rope_kv_write_paged_splitk": (_I, [_KP, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "b200_splitk_silu_mul": (_I, [_KP, _P, _P]),
    "b200_p2p_allreduce_rmsnorm": (_I, [_P, _P, _KP, _P, _P, _P, _P, _L, _L, _F, _P]),
    "b200_p2p_argmax": (_I, [_P, _P, _P, _L, _L, _L, _P, _P]),
    "b200_kv_alloc_create": (_P, [ctypes.c_int32]),
    "b200_kv_alloc_destroy": (None, [_P]),
    "b200_kv_alloc_num_free": (ctypes.c_int32, [_P]),
    "b200_kv_alloc_take": (_I, [_P, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]),
    "b200_kv_alloc_release": (_I, [_P, ctypes.POINTER(ctypes.c_int32), ctypes.c_int32]),
    "b200_decode_advance": (_I, [_P, _L, _P, _P, _P, _P, _P, _I, _P]),
    "b200_llama_embed": (_I, [_WP, _SP, _P]),
    "b200_llama_attn_block": (_I, [_WP, _SP, _I, _P]),
    "b200_llama_mlp_block": (_I, [_WP, _SP, _I, _P]),
    "b200_llama_head": (_I, [_WP, _SP, _P]),
    "b200_llama_step": (_I, [_WP, _SP, _P]),
})

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
