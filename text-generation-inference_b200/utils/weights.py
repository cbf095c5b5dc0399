"""safetensors lazy loader with the reference's tensor-parallel slicing rules.

Mirrors /root/reference/server/text_generation_server/utils/weights.py:14-229 (same method names, argument meaning
and error behaviour): column shards on the output dim with every fused prefix sharded separately then concatenated
(:115-142), row shards on the input dim with GPTQ scales/zeros sharded by groups (:144-201), int32 tensors never
cast (:72-75, 97-100).
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import torch
from safetensors import safe_open

QUANTIZE_CONFIG_FILENAME = "quantize_config.json"


class Weights:
    def __init__(self, filenames: List[Path], device, dtype, process_group, aliases: Optional[Dict[str, List[str]]] = None):
        routing = {}
        for filename in filenames:
            with safe_open(filename, framework="pytorch") as f:
                for k in f.keys():
                    if k in routing:
                        raise RuntimeError(f"Key {k} was found in multiple files: {filename} and {routing[k]}")
                    routing[k] = filename
        self.aliases = aliases or {}
        self.routing = routing
        self.device = device
        self.dtype = dtype
        self.process_group = process_group
        self._handles = {}

    def _get_handle(self, filename):
        if filename not in self._handles:
            self._handles[filename] = safe_open(filename, framework="pytorch")
        return self._handles[filename]

    def get_filename(self, tensor_name: str) -> Tuple[str, str]:
        filename = self.routing.get(tensor_name, None)
        if filename is None:
            for alias in self.aliases.get(tensor_name, []):
                filename = self.routing.get(alias, None)
                if filename is not None:
                    return str(filename), alias
            raise RuntimeError(f"weight {tensor_name} does not exist")
        return str(filename), tensor_name

    def _get_slice(self, tensor_name: str):
        filename, tensor_name = self.get_filename(tensor_name)
        return self._get_handle(filename).get_slice(tensor_name)

    def get_shape(self, tensor_name: str):
        return self._get_slice(tensor_name).get_shape()

    def get_tensor(self, tensor_name: str):
        filename, tensor_name = self.get_filename(tensor_name)
        tensor = self._get_handle(filename).get_tensor(tensor_name)
        if tensor.dtype not in [torch.int32, torch.int64]:
            tensor = tensor.to(dtype=self.dtype)
        return tensor.to(device=self.device)

    def get_partial_sharded(self, tensor_name: str, dim: int):
        world_size = self.process_group.size()
        rank = self.process_group.rank()
        slice_ = self._get_slice(tensor_name)
        size = slice_.get_shape()[dim]
        block_size = size // world_size
        start, stop = rank * block_size, (rank + 1) * block_size
        if dim == 0:
            tensor = slice_[start:stop]
        elif dim == 1:
            tensor = slice_[:, start:stop]
        else:
            raise NotImplementedError("Let's make that generic when needed")
        if tensor.dtype != torch.int32:
            tensor = tensor.to(dtype=self.dtype)
        return tensor.to(device=self.device)

    def get_sharded(self, tensor_name: str, dim: int):
        size = self._get_slice(tensor_name).get_shape()[dim]
        world_size = self.process_group.size()
        assert size % world_size == 0, f"The choosen size {size} is not compatible with sharding on {world_size} shards"
        return self.get_partial_sharded(tensor_name, dim)

    def get_multi_weights_col(self, prefixes: List[str], quantize: Optional[str], dim: int):
        if quantize == "gptq":
            try:
                qweight = torch.cat([self.get_sharded(f"{p}.qweight", dim=1) for p in prefixes], dim=1)
            except RuntimeError:
                raise RuntimeError("Cannot load `gptq` weight, make sure the model is already quantized")
            qzeros = torch.cat([self.get_sharded(f"{p}.qzeros", dim=1) for p in prefixes], dim=1)
            scales = torch.cat([self.get_sharded(f"{p}.scales", dim=1) for p in prefixes], dim=1)
            w = [self.get_tensor(f"{p}.g_idx") for p in prefixes]
            for w2 in w[1:]:
                torch.testing.assert_close(w2, w[0])
            g_idx = w[0]
            bits, groupsize = self._get_gptq_params()
            return (qweight, qzeros, scales, g_idx, bits, groupsize, bits == 4)
        w = [self.get_sharded(f"{p}.weight", dim=0) for p in prefixes]
        return torch.cat(w, dim=dim)

    def get_multi_weights_row(self, prefix: str, quantize: Optional[str]):
        if quantize == "gptq":
            bits, groupsize = self._get_gptq_params()
            if bits != 4:
                raise NotImplementedError("the B200 GPTQ kernel is 4-bit only (exllamav2.py:105)")
            g_idx_full = self.get_tensor(f"{prefix}.g_idx")
            trivial = torch.equal(g_idx_full.cpu(), (torch.arange(g_idx_full.shape[0], dtype=torch.int32) // groupsize)) \
                if groupsize > 0 else bool((g_idx_full == 0).all())
            if self.process_group.size() > 1 and not trivial and not bool((g_idx_full == 0).all()):
                # weights.py:150-156: act-order cannot be row-sharded by the fused kernel
                raise NotImplementedError("row tensor parallelism with act-order GPTQ is not supported")
            try:
                qweight = self.get_sharded(f"{prefix}.qweight", dim=0)
            except RuntimeError:
                raise RuntimeError("Cannot load `gptq` weight, make sure the model is already quantized")
            if groupsize >= 0:
                qzeros = self.get_sharded(f"{prefix}.qzeros", dim=0)
                scales = self.get_sharded(f"{prefix}.scales", dim=0)
            else:
                qzeros = self.get_tensor(f"{prefix}.qzeros")
                scales = self.get_tensor(f"{prefix}.scales")
            g_idx = g_idx_full if self.process_group.size() == 1 else None
            return (qweight, qzeros, scales, g_idx, bits, groupsize, True)
        return self.get_sharded(f"{prefix}.weight", dim=1)

    def _get_gptq_params(self) -> Tuple[int, int]:
        try:
            bits = self.get_tensor("gptq_bits").item()
            groupsize = self.get_tensor("gptq_groupsize").item()
        except RuntimeError as e:
            try:
                bits = self.gptq_bits
                groupsize = self.gptq_groupsize
            except Exception:
                raise e
        return bits, groupsize

    def _set_gptq_params(self, model_config: Any, model_path: str):
        config = model_config.to_dict() if hasattr(model_config, "to_dict") else dict(model_config)
        quantize_config = config.get("quantization_config")
        if quantize_config is None:
            filename = os.path.join(model_path, QUANTIZE_CONFIG_FILENAME)
            if not os.path.exists(filename):
                return
            with open(filename, "r") as f:
                quantize_config = json.load(f)
        self.gptq_bits = quantize_config["bits"]
        self.gptq_groupsize = quantize_config["group_size"]
