"""Model factory of the shard server (reference: models/__init__.py:33-141 `get_model`).

Only the hot-path branch exists: a flash decoder family served by `FlashCausalLM` over the paged KV pool.  FLASH_ATTENTION /
PAGED_ATTENTION (models/__init__.py:15-16) both select it - the KV cache here is always paged -, `deployment_framework` is
forced to `tgis_native` as the reference does for those modes (:48-54, :95-101), and anything else (seq2seq, the non-flash
`CausalLM` path, other engines) raises: that code is outside the scope of this library.
"""
import os
from typing import Optional

from ..utils.dist import get_torch_dtype, print_rank_n

__all__ = ["get_model"]


def get_model(model_name: str, revision: Optional[str], deployment_framework: str, dtype_str: str, quantize: Optional[str],
              max_sequence_length: Optional[int], memory_scaling_model=None):
    """`model_name` is a local directory with config.json, tokenizer files and safetensors shards (resolving hub names is
    the launcher's job, utils/hub.py)."""
    from ..inference_engine import FLASH_TYPES
    dtype = get_torch_dtype(dtype_str)
    if not os.path.isdir(model_name):
        raise ValueError(f"{model_name!r} is not a local model directory")
    from transformers import PretrainedConfig
    config_dict, _ = PretrainedConfig.get_config_dict(model_name)
    model_type = config_dict["model_type"]
    if model_type == "gpt2" and "--bigcode--" in model_name:  # tgis_native.py:37-38: starcoder checkpoints
        model_type = "gpt_bigcode"
    if model_type not in FLASH_TYPES:
        raise NotImplementedError(f"model type {model_type!r}: this library serves the flash decoder families {FLASH_TYPES} only")
    if deployment_framework != "tgis_native":
        print_rank_n(f"WARNING: Using deployment engine tgis_native rather than {deployment_framework} because the paged flash path is the only one")
    if quantize not in (None, "gptq"):
        raise ValueError(f"Unsupported quantization method: {quantize}")
    from .flash_causal_lm import FlashCausalLM
    return FlashCausalLM(model_name, revision, "tgis_native", dtype, quantize, None, max_sequence_length=max_sequence_length)
