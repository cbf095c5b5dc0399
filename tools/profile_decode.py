"""Profiling driver (run under ncu on the GPU box): builds a synthetic model of the named workload with a few layers,
prefills B prompts of `--ctx` tokens through the public API and runs `--steps` eager (un-graphed) decode steps so that
every kernel of a decode step appears as an individual launch.

  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv \
      python tools/profile_decode.py --workload llama2-7b-gptq --layers 2 --ctx 1536 --steps 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200_CUDA_GRAPHS"] = "false"

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="llama2-7b-gptq")
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--ctx", type=int, default=1536)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    arch, quantize, B, L0, L1 = bench.WORKLOADS[args.workload]
    B = args.batch or B
    bench.WORKLOADS[args.workload] = (arch, quantize, B, args.ctx, args.ctx + 64)
    model, cfg = bench.build_model(args.workload, 1, 0, args.layers)
    batch, errs = model.batch_type.from_pb(bench.make_batch_pb(B, args.ctx, 64), model.tokenizer, model.dtype, model.device, None, None, True)
    with torch.inference_mode():
        model.generate_token(batch, first=True)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("decode_steps")
        for _ in range(args.steps):
            model.generate_token(batch)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
    print("done")


if __name__ == "__main__":
    main()
