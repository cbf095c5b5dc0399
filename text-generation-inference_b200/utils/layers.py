"""Linear / embedding / rotary building blocks over the B200 kernels.

Mirrors /root/reference/server/text_generation_server/utils/layers.py: `FastLinear` (:88-111), `get_linear` (:172-203),
`TensorParallelHead` (:215-277), `TensorParallelColumnLinear` / `TensorParallelRowLinear` (:280-322),
`TensorParallelEmbedding` (:325-357), `FastLayerNorm` (:360-396), `PositionRotaryEmbedding` (:406-481).  Same names, constructor arguments and
sharding behaviour; the CUDA work goes through the C ABI (ops.py), collectives through torch.distributed (NCCL).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed
from torch import nn

from . import _ops
from .gptq.exllamav2 import Ex4bitLinearV2


class FastLinear(nn.Module):
    def __init__(self, weight, bias) -> None:
        super().__init__()
        self.weight = nn.Parameter(weight.contiguous(), requires_grad=False)
        self.bias = nn.Parameter(bias, requires_grad=False) if bias is not None else None

    @classmethod
    def load(cls, config, prefix: str, weights, bias: bool):
        weight = weights.get_tensor(f"{prefix}.weight")
        b = weights.get_tensor(f"{prefix}.bias") if bias else None
        return cls(weight, b)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        shape = input.shape[:-1] + (self.weight.shape[0],)
        x = input.reshape(-1, input.shape[-1]).contiguous()
        return _ops().gemm_f16(x, self.weight, self.bias).view(shape)


def get_linear(weight, bias, quantize):
    """utils/layers.py:172-203; one GPTQ backend only (no multi-backend dispatch)."""
    if quantize is None:
        return FastLinear(weight, bias)
    if quantize == "gptq":
        try:
            qweight, qzeros, scales, g_idx, bits, groupsize, use_gptq_cuda = weight
        except Exception:
            raise NotImplementedError("The passed weight is not `gptq` compatible, loader needs to be updated.")
        return Ex4bitLinearV2(qweight, qzeros, scales, g_idx, bias, bits, groupsize)
    raise NotImplementedError(f"Quantization `{quantize}` is not implemented yet.")


class SuperLayer(nn.Module):
    def __init__(self, linear):
        super().__init__()
        self.linear = linear

    def forward(self, x):
        return self.linear.forward(x)


class TensorParallelHead(SuperLayer):
    def __init__(self, linear, process_group, should_gather: bool):
        super().__init__(linear)
        self.process_group = process_group
        self.should_gather = should_gather

    @staticmethod
    def load(config, prefix: str, weights):
        if weights.process_group.size() > 1:
            try:
                weight = weights.get_sharded(f"{prefix}.weight", dim=0)
                should_gather = True
            except AssertionError:
                weight = weights.get_tensor(f"{prefix}.weight")
                should_gather = False
        else:
            weight = weights.get_tensor(f"{prefix}.weight")
            should_gather = False
        quantize = None if config.quantize == "gptq" else config.quantize  # GPTQ never quantizes heads (:236-237)
        return TensorParallelHead(get_linear(weight, bias=None, quantize=quantize), process_group=weights.process_group,
                                  should_gather=should_gather)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        output = super().forward(input)
        if not self.should_gather:
            return output
        # utils/layers.py:249-277: gather the vocab shards; rank r owns columns [r*V/tp, (r+1)*V/tp)
        world = self.process_group.size()
        rows = output.shape[0]
        gathered = output.new_empty(world * rows, output.shape[1])  # rank-major rows: the shape both NCCL and Gloo accept
        torch.distributed.all_gather_into_tensor(gathered, output.contiguous(), group=self.process_group)
        return gathered.view(world, rows, -1).permute(1, 0, 2).reshape(rows, -1)


class TensorParallelColumnLinear(SuperLayer):
    @classmethod
    def load(cls, config, prefix: str, weights, bias: bool):
        return cls.load_multi(config, [prefix], weights, bias, dim=0)

    @classmethod
    def load_multi(cls, config, prefixes: List[str], weights, bias: bool, dim: int):
        weight = weights.get_multi_weights_col(prefixes, quantize=config.quantize, dim=dim)
        if bias:
            b = torch.cat([weights.get_sharded(f"{p}.bias", dim=0) for p in prefixes], dim=dim)
        else:
            b = None
        return cls(get_linear(weight, b, config.quantize))


class TensorParallelRowLinear(SuperLayer):
    def __init__(self, linear, process_group):
        super().__init__(linear)
        self.process_group = process_group

    @classmethod
    def load(cls, config, prefix: str, weights, bias: bool):
        weight = weights.get_multi_weights_row(prefix, quantize=config.quantize)
        b = weights.get_tensor(f"{prefix}.bias") if bias and weights.process_group.rank() == 0 else None
        return cls(get_linear(weight, b, config.quantize), process_group=weights.process_group)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        out = super().forward(input)
        if self.process_group.size() > 1:
            torch.distributed.all_reduce(out, group=self.process_group)
        return out


class TensorParallelEmbedding(nn.Module):
    """Vocab-parallel lookup: out-of-shard ids give a zero row, then all-reduce (utils/layers.py:325-357)."""

    def __init__(self, prefix: str, weights, reduce=True):
        super().__init__()
        weight = weights.get_partial_sharded(f"{prefix}.weight", dim=0)
        num_embeddings = weights.get_shape(f"{prefix}.weight")[0]
        process_group = weights.process_group
        world_size, rank = process_group.size(), process_group.rank()
        block_size = num_embeddings // world_size
        self.min_id = rank * block_size
        self.max_id = min(num_embeddings, (rank + 1) * block_size)
        self.null_idx = block_size
        self.process_group = process_group
        self.reduce = reduce
        self.weight = nn.Parameter(weight.contiguous(), requires_grad=False)  # the kernel masks; no padded null row needed

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        out = _ops().embedding(self.weight, input.reshape(-1).to(torch.int64), vocab_start=self.min_id)
        if self.reduce and self.process_group.size() > 1:
            torch.distributed.all_reduce(out, group=self.process_group)
        return out.view(*input.shape, -1)


class FastLayerNorm(nn.Module):
    """Residual-add + LayerNorm in one kernel (utils/layers.py:360-396: FastLayerNorm over dropout_layer_norm).
    forward(hidden_states, residual=None) -> (normed, residual_out); residual None -> residual_out is hidden_states."""

    def __init__(self, weight, bias, eps: float):
        super().__init__()
        self.weight = nn.Parameter(weight.contiguous(), requires_grad=False)
        self.bias = nn.Parameter(bias.contiguous(), requires_grad=False) if bias is not None else None
        self.eps = eps

    @classmethod
    def load(cls, prefix: str, weights, eps: float):
        return cls(weights.get_tensor(f"{prefix}.weight"), weights.get_tensor(f"{prefix}.bias"), eps)

    def forward(self, hidden_states, residual=None):
        return _ops().layernorm_residual(hidden_states, residual, self.weight, self.bias, self.eps)


class PositionRotaryEmbedding(nn.Module):
    """fp32 inv_freq, cos/sin tables cached in the model dtype and gathered by position (utils/layers.py:406-481).
    The rotation itself is fused into the KV-write kernel (ops.rope_kv_write_paged), which reads these tables."""

    def __init__(self, inv_freq, scaling_factor=1.0):
        super().__init__()
        self.inv_freq = inv_freq
        self.scaling_factor = scaling_factor
        self._seq_len_cached = 0
        self._cos_cached = None
        self._sin_cached = None

    @classmethod
    def static(cls, dim, base, device, scaling_factor=1.0):
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, device=device, dtype=torch.float32) / dim))
        return cls(inv_freq, scaling_factor)

    @classmethod
    def load(cls, prefix, weights):
        dtype = weights.dtype
        weights.dtype = torch.float32
        inv_freq = weights.get_tensor(f"{prefix}.inv_freq")
        weights.dtype = dtype
        return cls(inv_freq)

    def _update_cos_sin_cache(self, dtype, device, seqlen):
        if seqlen > self._seq_len_cached or self._cos_cached.device != device or self._cos_cached.dtype != dtype:
            self._seq_len_cached = seqlen
            t = torch.arange(seqlen, device=device, dtype=self.inv_freq.dtype)
            if self.scaling_factor != 1.0:
                t = t / self.scaling_factor
            freqs = torch.outer(t, self.inv_freq.to(device=t.device))
            self._cos_cached = torch.cos(freqs).to(dtype)
            self._sin_cached = torch.sin(freqs).to(dtype)

    def tables(self, max_s: int, dtype, device):
        self._update_cos_sin_cache(dtype, device, max_s)
        return self._cos_cached, self._sin_cached

    def get_cos_sin(self, position_ids: torch.Tensor, max_s: int, dtype: torch.dtype):
        self._update_cos_sin_cache(dtype, position_ids.device, max_s)
        cos = torch.index_select(self._cos_cached, 0, position_ids)
        sin = torch.index_select(self._sin_cached, 0, position_ids)
        return cos.unsqueeze(1), sin.unsqueeze(1)


class LinearScalingPositionRotaryEmbedding(PositionRotaryEmbedding):
    @classmethod
    def static(cls, dim, base, scaling_factor, device):
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, device=device, dtype=torch.float32) / dim))
        return cls(inv_freq, scaling_factor)
