"""GPTQ int4 linear on the tcgen05 in-kernel-dequant GEMM.

Mirrors /root/reference/server/text_generation_server/utils/gptq/exllamav2.py:100-144 (`Ex4bitLinearV2`: same
constructor arguments, `post_init()` one-time repack, `forward`), with `exllamav2_kernels.make_q_matrix` /
`gemm_half_q_half` replaced by `b200_gptq_pack` / `b200_gemm_w4a16`: `post_init()` converts the checkpoint tensors once
into the kernel's streaming layout (`q_handle` = the packed buffer) and drops them.  No scratch `temp_dq`: the kernel
never materialises the fp16 matrix (contrast :65-97, :87).
"""
from __future__ import annotations

import torch
from torch import nn

from .. import _ops


class Ex4bitLinearV2(nn.Module):
    QUANT_TYPE = "exllamav2"

    def __init__(self, qweight, qzeros, scales, g_idx, bias, bits, groupsize):
        super().__init__()
        assert bits == 4, "b200 GPTQ kernel supports 4-bit only"  # exllamav2.py:105
        self.q_handle = None
        self.pack_layout = 0  # ops.W4_LAYOUT_*: how q_handle's unit records are ordered
        self.qweight = qweight
        self.qzeros = qzeros
        self.scales = scales.to(torch.float16)
        self.g_idx = g_idx
        self.bias = bias if bias is not None else None
        self.group_size = groupsize
        self.infeatures = self.qweight.shape[0] // bits * 32
        self.outfeatures = self.qweight.shape[1]
        self.height = self.infeatures
        self.width = self.outfeatures
        assert self.infeatures % 32 == 0 and self.outfeatures % 32 == 0  # exllamav2.py:118-119
        # act-order (exllamav2.py:31-48): a g_idx that is neither k // groupsize nor all zeros gets a row permutation
        self.q_perm = None
        if g_idx is not None and groupsize > 0:
            gi = g_idx.to(torch.int64)
            trivial = torch.equal(gi.cpu(), torch.arange(self.infeatures) // groupsize)
            if not trivial and not bool((gi == 0).all()):
                counts = torch.bincount(gi.cpu(), minlength=(self.infeatures + groupsize - 1) // groupsize)
                if not bool((counts[:-1] == groupsize).all()):
                    raise NotImplementedError("act-order g_idx with groups of unequal size")
                self.q_perm = torch.argsort(gi, stable=True).to(torch.int32).contiguous()

    def post_init(self, temp_dq=None, layout: int = 0):
        """exllamav2.py:124-137 (make_q_matrix): one-time conversion to the kernel layout; the checkpoint tensors are
        released afterwards (the packed buffer holds everything the GEMM reads).  `layout` = ops.W4_LAYOUT_GATE_UP for
        a fused [gate; up] projection (lets the step runtime fuse SiLU * up into the GEMM); results of forward() are
        the same either way."""
        if self.q_handle is None:
            assert self.qweight.device.type == "cuda"
            if layout == 1 and self.outfeatures % 256 != 0:
                layout = 0
            if self.q_perm is not None:
                self.q_perm = self.q_perm.to(self.qweight.device)
            self.q_handle = _ops().gptq_pack(self.qweight.contiguous(), self.qzeros.contiguous(), self.scales.contiguous(),
                                             self.group_size, layout, self.q_perm)
            self.pack_layout = layout
            self.qweight = self.qzeros = self.scales = None

    def forward(self, x, force_cuda=False):
        if self.q_handle is None:
            self.post_init()
        out_shape = x.shape[:-1] + (self.outfeatures,)
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        if self.q_perm is not None:
            x2 = _ops().permute_columns(x2, self.q_perm)
        out = _ops().gemm_w4a16(x2, self.q_handle, self.outfeatures, self.group_size, self.bias, layout=self.pack_layout)
        return out.view(out_shape)

    def temp_dq_size(self):
        return 0

    def temp_fwd_size(self, max_input_len, max_batch_size):
        return 0

    def scratch_space_fixed(self, max_input_len=4096, max_batch_size=16):
        return 0
