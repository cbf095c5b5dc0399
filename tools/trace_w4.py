"""Per-CTA phase timeline of one int4 GEMM launch (debug instrumentation in gemm_w4a16.cu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tgis_b200  # noqa: E402,F401
from tgis_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
lib = _lib.load()
for flags, T, N, K in [(0, 64, 22016, 4096), (1, 64, 22016, 4096), (2, 64, 22016, 4096), (3, 64, 22016, 4096)]:
    lib.b200_debug_w4_flags(flags)
    x = torch.randn(T, K, device=dev).half()
    qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 8, N), device=dev, dtype=torch.int32)
    qz = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 128, N // 8), device=dev, dtype=torch.int32)
    sc = (torch.rand(K // 128, N, device=dev) * 0.01).half()
    out = torch.empty(T, N, device=dev, dtype=torch.float16)
    for _ in range(3):
        ops.gemm_w4a16(x, qw, qz, sc, 128, out=out)
    trace = torch.zeros(160, 64, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    lib.b200_debug_w4_trace(trace.data_ptr())
    ops.gemm_w4a16(x, qw, qz, sc, 128, out=out)
    torch.cuda.synchronize()
    lib.b200_debug_w4_trace(None)
    tr = trace.cpu()
    used = tr[:, 0] > 0
    t0 = tr[used, 0].min().item()
    rel = (tr - t0).float() / 1e3  # us
    print(f"flags={flags} T={T} N={N} K={K}: ctas={int(used.sum())}")
    names = {0: "start", 1: "setup done", 2: "first data", 3: "first A tile", 63: "exit"}
    for seg in range(4):
        names.update({4 + 4 * seg: f"seg{seg} units done", 5 + 4 * seg: f"seg{seg} mma done", 6 + 4 * seg: f"seg{seg} epilogue done",
                      7 + 4 * seg: f"seg{seg} fixup done"})
    for j, nm in enumerate(["u8 dequant done", "u8 A free", "u8 next loaded", "u8 retired", "-", "-", "-", "-", "u10 dequant done", "u10 A free",
                            "u10 next loaded", "u10 retired"]):
        names[40 + j] = nm
    names[52] = "u8 next W landed"
    names[53] = "u10 next W landed"
    for k in sorted(names):
        col = rel[used, k]
        ok = tr[used, k] > 0
        if ok.any():
            c = col[ok]
            print(f"   {names[k]:22s} n={int(ok.sum()):4d}  min {c.min():7.2f}  median {c.median():7.2f}  max {c.max():7.2f} us")
