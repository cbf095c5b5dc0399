"""Per-CTA phase timeline of one int4 GEMM launch (debug instrumentation in gemm_w4a16.cu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tgis_b200  # noqa: E402,F401
from tgis_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
lib = _lib.load()
CASES = [(0, 64, 22016, 4096), (1, 64, 22016, 4096), (2, 64, 22016, 4096), (3, 64, 22016, 4096)]
if len(sys.argv) > 1:  # flags,T,N,K ...
    CASES = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
for flags, T, N, K in CASES:
    lib.b200_debug_w4_flags(flags)
    x = torch.randn(T, K, device=dev).half()
    qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 8, N), device=dev, dtype=torch.int32)
    qz = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 128, N // 8), device=dev, dtype=torch.int32)
    sc = (torch.rand(K // 128, N, device=dev) * 0.01).half()
    out = torch.empty(T, N, device=dev, dtype=torch.float16)
    packed = ops.gptq_pack(qw, qz, sc, 128)
    for _ in range(3):
        ops.gemm_w4a16(x, packed, N, 128, out=out)
    trace = torch.zeros(160, 64, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    lib.b200_debug_w4_trace(trace.data_ptr())
    ops.gemm_w4a16(x, packed, N, 128, out=out)
    torch.cuda.synchronize()
    lib.b200_debug_w4_trace(None)
    tr = trace.cpu()
    used = tr[:, 0] > 0
    t0 = tr[used, 0].min().item()
    rel = (tr - t0).float() / 1e3  # us
    print(f"flags={flags} T={T} N={N} K={K}: ctas={int(used.sum())}")
    names = {0: "start", 1: "setup done", 2: "first W landed (team 0)", 3: "first A tile handed over", 4: "team 0 done",
             5: "epilogue warps done", 6: "MMA warp issued all", 63: "exit (fix-ups done)"}
    for sg in range(4):
        names[8 + 2 * sg] = f"seg{sg} accumulator ready"
        names[9 + 2 * sg] = f"seg{sg} epilogue stores issued"
    for q, u in ((0, 8), (1, 12)):
        b = 16 + q * 8
        names.update({b: f"team0 u{u}: iteration start", b + 1: f"team0 u{u}: chunk0 dequantised", b + 2: f"team0 u{u}: A stage free",
                      b + 3: f"team0 u{u}: 4 STTM issued", b + 4: f"team0 u{u}: next unit in registers", b + 5: f"team0 u{u}: handed to MMA"})
    for q in range(4):
        names.update({32 + 4 * q: f"mma su{4 + q}: loop top", 33 + 4 * q: f"mma su{4 + q}: inputs ready", 34 + 4 * q: f"mma su{4 + q}: issued+committed"})
    names.update({40: "fix-up phase starts", 41: "fix0: all partials visible", 42: "fix0: my slice done (thread 0)", 43: "fix0: whole CTA done",
                  44: "fix1: all partials visible", 45: "fix1: my slice done (thread 0)", 46: "fix1: whole CTA done"})
    for k in sorted(names):
        col = rel[used, k]
        ok = tr[used, k] > 0
        if ok.any():
            c = col[ok]
            print(f"   {names[k]:22s} n={int(ok.sum()):4d}  min {c.min():7.2f}  median {c.median():7.2f}  max {c.max():7.2f} us")
