"""fp16 GEMM on the per-rank shapes of Llama-2-7B at TP = 8 (K = 1376 is not a multiple of the 64-wide k-block, N = 4000 not of
the 128-feature tile): a quick on-GPU check before the 8-GPU scaling run.  python tools/check_tp8_shapes.py"""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tgis_b200
from tgis_b200 import ops
g = torch.Generator().manual_seed(0)
for T, N, K in [(64, 4096, 1376), (64, 2752, 4096), (64, 1536, 4096), (64, 4096, 512), (64, 4000, 4096)]:
    x = torch.randn(T, K, generator=g).half(); w = (torch.randn(N, K, generator=g) * 0.05).half()
    ref = (x.float() @ w.float().t())
    got = ops.gemm_f16(x.cuda(), w.cuda()).float().cpu()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    print(T, N, K, "rel err", f"{err:.2e}", "OK" if err < 2e-3 else "BAD")
