"""Per-kernel device timeline of one graph-replayed decode step (b200_debug_step_trace: a %globaltimer stamp after every
launch).  The stamps serialise the stream, so PDL overlap between kernels is lost: read the table as "device time of each
kernel including its launch gap, warm L2 state as in the served step", and compare its sum with the untraced step time.

  python tools/step_trace.py [--workload llama3-8b-gptq] [--context 1536] [--layers N] [--out gpurun_out/step_trace.txt]
  torchrun --nproc-per-node 8 ... tools/step_trace.py ...   (tensor parallel: the launch list of rank 0's step)
"""
import argparse
import collections
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="llama3-8b-gptq")
    ap.add_argument("--context", type=int, default=None, help="default: the middle of the workload's trajectory")
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    arch, quantize, B, L0, L1 = bench.WORKLOADS[a.workload]
    world, rank = int(os.getenv("WORLD_SIZE", "1")), int(os.getenv("RANK", "0"))  # under torchrun: every rank traces, rank 0 reports
    model, cfg = bench.build_model(a.workload, world, rank, a.layers)
    from tgis_b200 import _lib
    lib = _lib.load()
    dev = model.device
    if a.context is None:
        a.context = (L0 + L1) // 2
    L0 = min(L0, a.context - 8)
    batch, errs = model.batch_type.from_pb(bench.make_batch_pb(B, L0, L1 - L0), model.tokenizer, model.dtype, dev, None, None, True)
    lines = []
    with torch.inference_mode():
        model.generate_token(batch, first=True)
        cur = L0
        while cur < a.context:
            model.generate_token(batch)
            cur += 1
        st = batch._fused

        def step(use_graph):
            model._run_fused_step(batch, st, use_graph)

        # untraced graph-replayed step time
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            step(True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            step(True)
        e1.record()
        torch.cuda.synchronize()
        untraced_us = e0.elapsed_time(e1) / 10 * 1e3
        # traced step, captured in its own graph so that host launch latency is not in the picture
        cap = 4096
        buf = torch.zeros(cap, dtype=torch.int64, device=dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            lib.b200_debug_step_trace(buf.data_ptr(), cap)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                lib.b200_debug_step_trace_begin(torch.cuda.current_stream().cuda_stream)
                model._run_fused_step(batch, st, False)
            names_buf = ctypes.create_string_buffer(1 << 20)
            n = lib.b200_debug_step_trace_names(names_buf, len(names_buf))
            lib.b200_debug_step_trace(None, 0)
            for _ in range(3):
                g.replay()
            side.synchronize()
        names = names_buf.value.decode().strip().split("\n")
        t = buf[:n].cpu().tolist()
        dt = [(names[i].split("<")[0].replace("b200::", ""), (t[i] - t[i - 1]) / 1e3) for i in range(1, n)]
        total = (t[n - 1] - t[0]) / 1e3
        lines.append(f"workload {a.workload} tp{world} context {cur} layers {cfg.num_hidden_layers}: untraced graph step {untraced_us:.1f} us, "
                     f"traced (serialised) step {total:.1f} us, {n - 1} launches")
        agg = collections.OrderedDict()
        for k, v in dt:
            agg.setdefault(k, []).append(v)
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            v2 = sorted(v)
            lines.append(f"  {k:40s} n={len(v):4d} median {v2[len(v2) // 2]:8.2f} us  sum {sum(v):9.1f} us  {100 * sum(v) / total:5.1f} %")
        # one layer in launch order (the second layer: steady state)
        per_layer = (n - 1) // max(cfg.num_hidden_layers, 1)
        lines.append("  -- launches 1.." + str(min(n - 1, 3 * per_layer)) + " in order:")
        for i, (k, v) in enumerate(dt[:3 * per_layer]):
            lines.append(f"     {i:3d} {k:40s} {v:8.2f} us")
    text = "\n".join(lines)
    if rank != 0:
        return
    print(text)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
