"""sm_100a replacement of the reference's only in-tree CUDA, `server/custom_kernels`.

`fused_attention_cuda.forward` and `fused_bloom_attention_cuda.forward` keep the argument lists and return tuples of
/root/reference/server/custom_kernels/custom_kernels/fused_attention_cuda.cu:113-242 and
fused_bloom_attention_cuda.cu (the non-flash GPT-NeoX / BLOOM attention, called at neox_modeling.py:214 and
bloom_modeling.py:394).  As in the reference the two batched products are library GEMMs (`torch.bmm` / `baddbmm`, i.e.
cuBLAS - these are plain library GEMMs off the decode hot path); the fused cast + mask + softmax + cast kernel between them is
`b200_masked_softmax` (csrc/elementwise.cu), without the reference kernel's kv_length <= 4096 limit.
"""
from . import fused_attention_cuda, fused_bloom_attention_cuda  # noqa: F401
