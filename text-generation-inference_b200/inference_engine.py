"""The `tgis_native` inference engine for the hot-path models: picks the model class by `model_type`, brings up the
process group, opens the safetensors shards with the TP slicing rules and builds the model.

Mirrors /root/reference/server/text_generation_server/inference_engine/engine.py:11-37 (`BaseInferenceEngine`:
config/tokenizer loading, RANK / WORLD_SIZE, device = rank % device_count) and inference_engine/tgis_native.py:24-139.
Only the flash decoder families of the hot path exist here (llama, gpt_neox, gpt_bigcode and falcon / RefinedWeb); other model
types raise NotImplementedError.
"""
from __future__ import annotations

import glob
import os
from typing import Any, Optional

import torch
import torch.distributed

from .utils.dist import initialize_torch_distributed
from .utils.weights import Weights

FLASH_TYPES = ["llama", "gpt_neox", "gpt_bigcode", "falcon", "RefinedWeb", "RefinedWebModel"]


def local_weight_files(model_path: str, extension: str = ".safetensors"):
    return sorted(glob.glob(os.path.join(model_path, f"*{extension}")))


class InferenceEngine:
    def __init__(self, model_path: str, model_class, dtype: torch.dtype, quantize: Optional[str], model_config: Optional[Any],
                 max_sequence_length: Optional[int], weights: Optional[Weights] = None, tokenizer=None):
        if model_config is None:
            from transformers import AutoConfig
            model_config = AutoConfig.from_pretrained(model_path)
        self._config = model_config
        if tokenizer is None:
            from transformers import AutoTokenizer
            tokenizer = AutoTokenizer.from_pretrained(model_path, padding_side="left", truncation_side="left")
        self.tokenizer = tokenizer
        self.rank = int(os.getenv("RANK", "0"))
        self.world_size = int(os.getenv("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise NotImplementedError("the B200 tgis_native engine needs a CUDA device (no CPU fallback)")
        gpu_count = torch.cuda.device_count()
        assert self.world_size <= gpu_count, f"{self.world_size} shards configured but only {gpu_count} GPUs detected"
        device_index = self.rank % gpu_count
        torch.cuda.set_device(device_index)
        self.device = torch.device("cuda", device_index)

        model_type = self._config.model_type
        if model_type == "gpt2" and "--bigcode--" in model_path:  # tgis_native.py:37-38: starcoder checkpoints
            model_type = "gpt_bigcode"
        if model_type not in FLASH_TYPES:
            raise NotImplementedError(f"Flash attention currently only supported by the following model types: {FLASH_TYPES}")
        aliases = None
        if model_type == "llama":
            if getattr(self._config, "tie_word_embeddings", False):
                aliases = {"lm_head.weight": ["model.embed_tokens.weight"]}
            # one class serves both calling conventions of the reference (flash_llama_modeling.py / paged_llama_modeling.py;
            # models/__init__.py:48-79 picks between them with PAGED_ATTENTION): the KV cache here is always paged
            from .models.custom_modeling.paged_llama_modeling import PagedLlamaForCausalLM
            model_class = PagedLlamaForCausalLM
        elif model_type == "gpt_neox":  # tgis_native.py:75-79
            from .models.custom_modeling.flash_neox_modeling import FlashGPTNeoXForCausalLM
            model_class = FlashGPTNeoXForCausalLM
        elif model_type == "gpt_bigcode":  # tgis_native.py:83-92
            archs = getattr(self._config, "architectures", None) or [""]
            self._config.transpose = archs[0].startswith("GPT2")
            if not getattr(self._config, "multi_query", True):
                raise NotImplementedError("gpt_bigcode without multi_query is not a flash santacoder model")
            from .models.custom_modeling.flash_santacoder_modeling import FlashSantacoderForCausalLM
            model_class = FlashSantacoderForCausalLM
        elif model_type in ("falcon", "RefinedWeb", "RefinedWebModel"):  # tgis_native.py:59-73
            if getattr(self._config, "alibi", False):
                raise NotImplementedError("alibi is not supported by this version of the model")
            from .models.custom_modeling.flash_rw_modeling import FlashRWForCausalLM
            model_class = FlashRWForCausalLM
        self._config.quantize = quantize
        self.process_group = initialize_torch_distributed(self.world_size, self.rank)
        self.master = self.rank == 0
        if self.world_size > 1:
            torch.distributed.barrier(group=self.process_group)
        if weights is None:
            filenames = local_weight_files(model_path)
            if not filenames:
                raise ValueError("No safetensors weights found - required for tgis_native engine")
            weights = Weights(filenames, device=self.device, dtype=dtype, process_group=self.process_group, aliases=aliases)
        if quantize == "gptq":
            weights._set_gptq_params(self._config, model_path)
        model = model_class(self._config, weights)
        if self.world_size > 1:
            torch.distributed.barrier(group=self.process_group)
        if not hasattr(model, "config"):
            model.config = self._config
        self.model = model.eval()

    def get_components(self):
        return self.model.config, self.tokenizer, self.model

    def get_device(self) -> torch.device:
        return self.device
