"""One int4 GEMM shape, a few launches (for ncu captures): python tools/one_w4.py T N K [launches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tgis_b200  # noqa: E402,F401
from tgis_b200 import ops  # noqa: E402

T, N, K = (int(v) for v in sys.argv[1:4])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = "cuda:0"
x = torch.randn(T, K, device=dev).half()
qz = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 128, N // 8), device=dev, dtype=torch.int32)
sc = (torch.rand(K // 128, N, device=dev) * 0.01).half()
packs = [ops.gptq_pack(torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 8, N), device=dev, dtype=torch.int32), qz, sc, 128) for _ in range(n)]
out = torch.empty(T, N, device=dev, dtype=torch.float16)
for i in range(n):
    ops.gemm_w4a16(x, packs[i], N, 128, out=out)
torch.cuda.synchronize()
