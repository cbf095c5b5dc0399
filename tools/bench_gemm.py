"""Micro-benchmark of the two linears (CUDA events, back-to-back launches, inputs rotated through > L2 of weights)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tgis_b200  # noqa: E402,F401
from tgis_b200 import ops  # noqa: E402

dev = "cuda:0"


def timeit(fn, n=16, reps=5):
    """GPU time per launch: n launches captured in a CUDA graph (no host overhead in the timed region)."""
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3  # us


def main():
    only_w4 = "--w4" in sys.argv
    for a in sys.argv:
        if a.startswith("--flags="):
            from tgis_b200 import _lib
            _lib.load().b200_debug_w4_flags(int(a.split("=")[1]))
            print("debug flags", a)
    shapes = [(64, 128, 128), (64, 128, 4096), (64, 4096, 128), (64, 4096, 4096), (64, 12288, 4096), (64, 22016, 4096), (64, 4096, 11008),
              (64, 6144, 4096), (64, 28672, 4096), (64, 4096, 14336), (64, 32000, 4096), (64, 128256, 4096),
              (1, 4096, 4096), (16, 4096, 4096), (256, 4096, 4096), (4096, 4096, 4096)]
    for T, N, K in shapes:
        nbuf = max(1, min(8, int(300e6 // (N * K // 2 + 1))))
        x = torch.randn(T, K, device=dev).half()
        qz = torch.randint(-2 ** 31, 2 ** 31 - 1, (max(K // 128, 1), N // 8), device=dev, dtype=torch.int32)
        sc = (torch.rand(max(K // 128, 1), N, device=dev) * 0.01).half()
        qw = [ops.gptq_pack(torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 8, N), device=dev, dtype=torch.int32), qz, sc, 128)
              for _ in range(nbuf)]
        out = torch.empty(T, N, device=dev, dtype=torch.float16)
        us = timeit(lambda i: ops.gemm_w4a16(x, qw[i % nbuf], N, 128, out=out))
        bytes_ = N * K / 2 + N * K / 128 * 4 + T * K * 2 + T * N * 2
        print(f"w4a16 T={T:4d} N={N:6d} K={K:6d}: {us:8.1f} us  {bytes_ / us / 1e3:8.1f} GB/s")
        if only_w4:
            continue
        nbuf = max(1, min(8, int(600e6 // (N * K * 2))))
        w = [torch.randn(N, K, device=dev).half() for _ in range(nbuf)]
        us = timeit(lambda i: ops.gemm_f16(x, w[i % nbuf], out=out))
        bytes_ = N * K * 2 + T * K * 2 + T * N * 2
        # cuBLAS on the same shapes, same graph-replayed rotation over > L2 of weights (the library kernel to beat)
        us_cublas = timeit(lambda i: torch.matmul(x, w[i % nbuf].t(), out=out))
        print(f"f16   T={T:4d} N={N:6d} K={K:6d}: {us:8.1f} us  {bytes_ / us / 1e3:8.1f} GB/s   cuBLAS {us_cublas:8.1f} us  "
              f"{bytes_ / us_cublas / 1e3:8.1f} GB/s   ratio {us_cublas / us:5.2f}x")
        del w


if __name__ == "__main__":
    main()
