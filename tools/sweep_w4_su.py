"""Sweep the stream-K cut (super-units per CTA, B200_W4_SU) of the int4 GEMM for one shape: python tools/sweep_w4_su.py T N K su..."""
import os
import subprocess
import sys

T, N, K = sys.argv[1:4]
for su in sys.argv[4:]:
    env = dict(os.environ, B200_W4_SU=su)
    code = f"""
import sys, torch
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import tgis_b200
from tgis_b200 import ops
sys.path.insert(0, {os.path.dirname(os.path.abspath(__file__))!r})
from bench_gemm import timeit
T, N, K = {T}, {N}, {K}
dev = "cuda:0"
x = torch.randn(T, K, device=dev).half()
qz = torch.randint(-2**31, 2**31 - 1, (K // 128, N // 8), device=dev, dtype=torch.int32)
sc = (torch.rand(K // 128, N, device=dev) * 0.01).half()
nbuf = max(1, min(8, int(300e6 // (N * K // 2 + 1))))
qw = [ops.gptq_pack(torch.randint(-2**31, 2**31 - 1, (K // 8, N), device=dev, dtype=torch.int32), qz, sc, 128) for _ in range(nbuf)]
out = torch.empty(T, N, device=dev, dtype=torch.float16)
us = timeit(lambda i: ops.gemm_w4a16(x, qw[i % nbuf], N, 128, out=out))
print(f"T={{T}} N={{N}} K={{K}} su_per_cta={su}: {{us:.1f}} us")
"""
    subprocess.run([sys.executable, "-c", code], env=env)
