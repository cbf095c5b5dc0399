"""CPU checks of the int4 weight stream and of the kernel's dequant arithmetic, restated in numpy
(oracle/gptq.py): the magic-number sequence of csrc/gemm_w4a16.cu on the packed records must give, bit for bit, the formula
of record fp16(scale * (q - (zero + 1))) (utils/gptq/quant_linear.py:184-192) for every stored zero 0..15, group sizes
32 / 64 / 128 / one group, ragged K and N, the gate|up record order and act-order rows."""
import numpy as np
import pytest
import torch

from oracle import gptq as ogptq


@pytest.mark.parametrize("K,N,gs,layout,act_order", [(256, 256, 128, 0, False), (320, 288, 64, 0, False), (128, 160, 32, 0, False),
                                                      (256, 512, 128, 1, False), (384, 256, -1, 0, False), (256, 256, 128, 0, True)])
def test_kernel_dequant_arithmetic_is_bit_exact(K, N, gs, layout, act_order):
    g = torch.Generator().manual_seed(K + N)
    w = torch.randn(N, K, generator=g) * 0.05
    qweight, qzeros, scales, g_idx = ogptq.quantize_rtn(w, gs)
    qzeros = torch.randint(-2 ** 31, 2 ** 31 - 1, qzeros.shape, generator=g, dtype=torch.int64).to(torch.int32)  # all zero values
    perm = None
    if act_order:
        g_idx = (torch.randperm(K, generator=g) // gs).to(torch.int32)
        perm = torch.argsort(g_idx.to(torch.int64), stable=True).to(torch.int32)
    rec = ogptq.unit_records(qweight, qzeros, scales, gs, layout, perm)
    assert rec.shape[1] == ogptq.packed_nkb(K) and rec.shape[0] == ((N + 127) // 128 + 1) // 2
    got = ogptq.dequant_records_like_the_kernel(rec, K, N, gs, layout)
    exp = ogptq.dequantize(qweight, qzeros, scales, g_idx if act_order else None, gs)
    if act_order:
        exp = exp[perm.to(torch.int64)]  # packed rows are in group order
    assert torch.equal(got, exp)


def test_packed_nkb_rule():
    assert [ogptq.packed_nkb(k) for k in (128, 320, 4096, 11008, 14336, 5504)] == [1, 3, 32, 88, 112, 43]
