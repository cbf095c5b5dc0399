"""Micro-benchmark of the attention kernels (CUDA events over graph-replayed launches):
  prefill: b200_attn_prefill_varlen (HMMA, fresh q/k/v) vs b200_attn_prefill_paged (tcgen05 + TMA through the block pool)
  decode : b200_attn_decode_paged at the served shapes
  python tools/bench_attn.py [--out gpurun_out/bench_attn.txt]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tgis_b200  # noqa: E402,F401
from tgis_b200 import ops  # noqa: E402

dev = "cuda:0"


def timeit(fn, n=4, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--decode-only", action="store_true")
    a = ap.parse_args()
    lines = []
    g = torch.Generator(device=dev).manual_seed(0)
    for name, B, L, h, kv, d in [("llama3-8b prefill", 64, 1024, 32, 8, 128), ("llama2-7b prefill", 64, 1024, 32, 32, 128),
                                 ("tinyllama prefill", 32, 512, 32, 4, 64), ("8b add-on prefill 16 x 640", 16, 640, 32, 8, 128)]:
        if a.decode_only:
            break
        T = B * L
        qkv = torch.randn(T, (h + 2 * kv) * d, device=dev, generator=g).half()
        q = qkv[:, :h * d].view(T, h, d)
        k = qkv[:, h * d:(h + kv) * d].view(T, kv, d)
        v = qkv[:, (h + kv) * d:].view(T, kv, d)
        cu = torch.arange(0, (B + 1) * L, L, dtype=torch.int32, device=dev)
        pages = (L + 15) // 16
        k_pool, v_pool = ops.kv_pool_alloc(B * pages + 1, kv, d, dev)
        k_pool.normal_(generator=g)
        v_pool.normal_(generator=g)
        bt = torch.arange(B * pages, dtype=torch.int32, device=dev).view(B, pages)
        ctx = torch.full((B,), L, dtype=torch.int32, device=dev)
        out = torch.empty(T, h, d, dtype=torch.float16, device=dev)
        flops = 4.0 * B * h * d * L * L / 2
        t_old = timeit(lambda: ops.attn_prefill_varlen(q, k, v, cu, L, d ** -0.5, True, out))
        t_new = timeit(lambda: ops.attn_prefill_paged(q, k_pool, v_pool, bt, ctx, cu, L, d ** -0.5, out))
        lines.append(f"{name:28s} B={B} L={L} h={h} kv={kv} d={d}: varlen HMMA {t_old:9.1f} us ({flops / t_old / 1e6:7.1f} TFLOP/s)   "
                     f"paged tcgen05 {t_new:9.1f} us ({flops / t_new / 1e6:7.1f} TFLOP/s)")
        del qkv, k_pool, v_pool
    for name, B, L, h, kv, d in [("llama3-8b decode", 64, 1549, 32, 8, 128), ("llama2-7b decode", 64, 1549, 32, 32, 128),
                                 ("8b tp2 decode", 64, 1549, 16, 4, 128), ("8b tp4 decode", 64, 1549, 8, 2, 128),
                                 ("8b tp8 decode", 64, 1549, 4, 1, 128), ("70b tp8 decode bs128", 128, 1536, 8, 1, 128),
                                 ("70b tp8 decode bs96 ctx768", 96, 768, 8, 1, 128), ("7b fp16 tp8 decode", 64, 1549, 4, 4, 128)]:
        pages = (2048 + 15) // 16
        k_pool, v_pool = ops.kv_pool_alloc(B * pages + 1, kv, d, dev)
        k_pool.normal_(generator=g)
        v_pool.normal_(generator=g)
        bt = torch.arange(B * pages, dtype=torch.int32, device=dev).view(B, pages)
        ctx = torch.full((B,), L, dtype=torch.int32, device=dev)
        q = torch.randn(B, h, d, device=dev, generator=g).half()
        out = torch.empty(B, h, d, dtype=torch.float16, device=dev)
        t = timeit(lambda: ops.attn_decode_paged(q, k_pool, v_pool, bt, ctx, 2048, d ** -0.5, kv, out), n=8)
        nbytes = B * L * 2 * kv * d * 2
        lines.append(f"{name:28s} B={B} L={L} h={h} kv={kv} d={d}: {t:8.1f} us  {nbytes / t / 1e3:7.1f} GB/s  ({nbytes / 1e6:.1f} MB)")
        del k_pool, v_pool
    text = "\n".join(lines)
    print(text)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
