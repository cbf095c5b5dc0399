import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tgis_b200
from tgis_b200 import ops
from oracle import gptq as ogptq
dev = "cuda:0"
for (T, N, K) in [(64, 4096, 4096), (33, 2560, 2048), (64, 4096, 11008)]:
    g = torch.Generator().manual_seed(T + N + K)
    w = torch.randn(N, K, generator=g) * 0.05
    qweight, qzeros, scales, g_idx = ogptq.quantize_rtn(w, 128)
    x = torch.randn(T, K, generator=g).half()
    ref = ogptq.gemm_half_q_half(x, qweight, qzeros, scales, None, 128).float()
    packed = ops.gptq_pack(qweight.to(dev), qzeros.to(dev), scales.to(dev), 128)
    for rep in range(3):
        got = ops.gemm_w4a16(x.to(dev), packed, N, 128).float().cpu()
        err = (got - ref).abs()
        bad = err > (1e-3 * ref.abs().max() + 2e-3 * ref.abs())
        nt = N // 128
        nkb = K // 128
        total = nt * nkb
        upc = (total + 147) // 148
        tiles = []
        for t in range(nt):
            b = bad[:, t * 128:(t + 1) * 128]
            if b.any():
                rows = sorted(set(torch.nonzero(b)[:, 0].tolist()))
                cols = sorted(set(torch.nonzero(b)[:, 1].tolist()))
                c_first = (t * nkb) // upc
                c_last = ((t + 1) * nkb - 1) // upc
                segs = [(c, max(t * nkb, c * upc) - c * upc, min((t + 1) * nkb, (c + 1) * upc) - c * upc) for c in range(c_first, c_last + 1)]
                tiles.append((t, len(rows), len(cols), int(b.sum()), segs))
        print(f"T={T} N={N} K={K} rep={rep} upc={upc}: bad tiles:")
        for tt in tiles:
            print("   ", tt)
