"""Lazy safetensors reader that hands every rank its tensor-parallel share of a checkpoint.

Keeps the public surface of /root/reference/server/text_generation_server/utils/weights.py (`Weights(filenames, device, dtype,
process_group, aliases)`, `get_shape / get_tensor / get_partial_sharded / get_sharded / get_multi_weights_col /
get_multi_weights_row`, `_get_gptq_params / _set_gptq_params`) and its rules, which the golden fixture
tests/golden/weights_shards.npz pins against the reference itself:
  * a rank's share of a dimension of size n is the block [rank * (n // world), (rank + 1) * (n // world)) (:79-90);
    `get_sharded` additionally insists on n % world == 0 (:92-101);
  * floating tensors are cast to the model dtype, integer tensors (GPTQ qweight / qzeros / g_idx) never are (:72-75, :97-100);
  * column-parallel: every fused prefix (q, k, v / gate, up) is sharded on its own output dimension first and the shares
    are concatenated afterwards (:115-142); GPTQ tensors are [in, out], so their output dimension is dim 1;
  * row-parallel: the input dimension is sharded — dim 1 of an fp16 weight, dim 0 of qweight and, when the checkpoint has
    groups, of qzeros / scales (:144-201); g_idx survives only un-sharded.
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import torch
from safetensors import safe_open

QUANTIZE_CONFIG_FILENAME = "quantize_config.json"
_INTEGER_DTYPES = (torch.int32, torch.int64)
_NOT_QUANTIZED = "Cannot load `gptq` weight, make sure the model is already quantized"


class Weights:
    def __init__(self, filenames: List[Path], device, dtype, process_group, aliases: Optional[Dict[str, List[str]]] = None):
        self.device, self.dtype, self.process_group = device, dtype, process_group
        self.aliases = aliases or {}
        self.routing: Dict[str, Any] = {}  # tensor name -> file holding it
        self._handles: Dict[str, Any] = {}
        for filename in filenames:
            with safe_open(filename, framework="pytorch") as f:
                for name in f.keys():
                    other = self.routing.setdefault(name, filename)
                    if other is not filename:
                        raise RuntimeError(f"Key {name} was found in multiple files: {filename} and {other}")

    # ------------------------------------------------------------------------------------------ file access
    def get_filename(self, tensor_name: str) -> Tuple[str, str]:
        """(file, stored name): the name itself, else the first alias present in the checkpoint."""
        for candidate in [tensor_name, *self.aliases.get(tensor_name, [])]:
            if candidate in self.routing:
                return str(self.routing[candidate]), candidate
        raise RuntimeError(f"weight {tensor_name} does not exist")

    def _open(self, filename: str):
        handle = self._handles.get(filename)
        if handle is None:
            handle = self._handles[filename] = safe_open(filename, framework="pytorch")
        return handle

    def _get_slice(self, tensor_name: str):
        filename, stored = self.get_filename(tensor_name)
        return self._open(filename).get_slice(stored)

    def _place(self, tensor: torch.Tensor, keep: Tuple[torch.dtype, ...]) -> torch.Tensor:
        if tensor.dtype not in keep:
            tensor = tensor.to(dtype=self.dtype)
        return tensor.to(device=self.device)

    # ------------------------------------------------------------------------------------------ whole tensors
    def get_shape(self, tensor_name: str):
        return self._get_slice(tensor_name).get_shape()

    def get_tensor(self, tensor_name: str):
        filename, stored = self.get_filename(tensor_name)
        return self._place(self._open(filename).get_tensor(stored), _INTEGER_DTYPES)

    # ------------------------------------------------------------------------------------------ this rank's share
    def _my_block(self, size: int) -> Tuple[int, int]:
        block = size // self.process_group.size()
        first = self.process_group.rank() * block
        return first, first + block

    def get_partial_sharded(self, tensor_name: str, dim: int):
        view = self._get_slice(tensor_name)
        lo, hi = self._my_block(view.get_shape()[dim])
        if dim == 0:
            share = view[lo:hi]
        elif dim == 1:
            share = view[:, lo:hi]
        else:
            raise NotImplementedError("Let's make that generic when needed")
        return self._place(share, (torch.int32,))

    def get_sharded(self, tensor_name: str, dim: int):
        size, world = self.get_shape(tensor_name)[dim], self.process_group.size()
        assert size % world == 0, f"The choosen size {size} is not compatible with sharding on {world} shards"
        return self.get_partial_sharded(tensor_name, dim)

    def _gptq_shares(self, prefix: str, dim: int) -> torch.Tensor:
        try:
            return self.get_sharded(f"{prefix}.qweight", dim=dim)
        except RuntimeError:
            raise RuntimeError(_NOT_QUANTIZED)

    # ------------------------------------------------------------------------------------------ linears
    def get_multi_weights_col(self, prefixes: List[str], quantize: Optional[str], dim: int):
        if quantize != "gptq":
            return torch.cat([self.get_sharded(f"{p}.weight", dim=0) for p in prefixes], dim=dim)
        qweight = torch.cat([self._gptq_shares(p, 1) for p in prefixes], dim=1)
        qzeros, scales = (torch.cat([self.get_sharded(f"{p}.{kind}", dim=1) for p in prefixes], dim=1) for kind in ("qzeros", "scales"))
        g_idx, *others = [self.get_tensor(f"{p}.g_idx") for p in prefixes]
        for other in others:  # fused projections share their input dimension, hence their act-order
            torch.testing.assert_close(other, g_idx)
        bits, groupsize = self._get_gptq_params()
        return (qweight, qzeros, scales, g_idx, bits, groupsize, bits == 4)

    def get_multi_weights_row(self, prefix: str, quantize: Optional[str]):
        if quantize != "gptq":
            return self.get_sharded(f"{prefix}.weight", dim=1)
        bits, groupsize = self._get_gptq_params()
        if bits != 4:
            raise NotImplementedError("the B200 GPTQ kernel is 4-bit only (exllamav2.py:105)")
        sharded = self.process_group.size() > 1
        g_idx = self.get_tensor(f"{prefix}.g_idx")
        all_zero = bool((g_idx == 0).all())
        in_order = all_zero if groupsize <= 0 else torch.equal(
            g_idx.cpu(), torch.arange(g_idx.shape[0], dtype=torch.int32) // groupsize)
        if sharded and not (in_order or all_zero):
            # weights.py:150-156: the reference leaves the exllama path here; this library has no other GPTQ backend
            raise NotImplementedError("row tensor parallelism with act-order GPTQ is not supported")
        qweight = self._gptq_shares(prefix, 0)
        per_group = self.get_sharded if groupsize >= 0 else (lambda name, dim: self.get_tensor(name))
        qzeros, scales = per_group(f"{prefix}.qzeros", dim=0), per_group(f"{prefix}.scales", dim=0)
        return (qweight, qzeros, scales, None if sharded else g_idx, bits, groupsize, True)

    # ------------------------------------------------------------------------------------------ GPTQ meta
    def _get_gptq_params(self) -> Tuple[int, int]:
        """(bits, groupsize): scalar tensors in the checkpoint (quantize.py:814-815), else what _set_gptq_params found."""
        try:
            return self.get_tensor("gptq_bits").item(), self.get_tensor("gptq_groupsize").item()
        except RuntimeError as missing:
            if hasattr(self, "gptq_bits") and hasattr(self, "gptq_groupsize"):
                return self.gptq_bits, self.gptq_groupsize
            raise missing

    def _set_gptq_params(self, model_config: Any, model_path: str):
        """`quantization_config` of config.json, else quantize_config.json next to the weights (:203-229)."""
        as_dict = model_config.to_dict() if hasattr(model_config, "to_dict") else dict(model_config)
        found = as_dict.get("quantization_config")
        if found is None:
            side_file = os.path.join(model_path, QUANTIZE_CONFIG_FILENAME)
            if not os.path.exists(side_file):
                return
            with open(side_file, "r") as f:
                found = json.load(f)
        self.gptq_bits, self.gptq_groupsize = found["bits"], found["group_size"]
