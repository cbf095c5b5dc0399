"""Host-side profile of the e2e decode loop (generate_token driven from host buffers): where the time between two graph replays goes.
  python tools/profile_e2e.py [--workload llama3-8b-gptq] [--steps 200]"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="llama3-8b-gptq")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--layers", type=int, default=None)
    a = ap.parse_args()
    arch, quantize, B, L0, L1 = bench.WORKLOADS[a.workload]
    model, cfg = bench.build_model(a.workload, 1, 0, a.layers)
    batch, _ = model.batch_type.from_pb(bench.make_batch_pb(B, L0, L1 - L0), model.tokenizer, model.dtype, model.device, None, None, True)
    host_ids = torch.empty(B, dtype=torch.int64).pin_memory()
    with torch.inference_mode():
        toks = model.generate_token(batch, first=True)[0]
        for _ in range(8):
            toks = model.generate_token(batch)[0]

        def loop(n):
            nonlocal toks
            for _ in range(n):
                host_ids.copy_(torch.tensor([t.token_id for t in toks], dtype=torch.int64))
                batch.input_ids.copy_(host_ids, non_blocking=True)
                toks = model.generate_token(batch)[0]

        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loop(a.steps)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / a.steps * 1e3
        # device-only time of the same step (graph replay back to back)
        st = batch._fused
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            st["graph"].replay()
        e1.record()
        torch.cuda.synchronize()
        dev_ms = e0.elapsed_time(e1) / 20
        print(f"e2e {wall:.3f} ms/step, graph replay alone {dev_ms:.3f} ms/step -> host + sync overhead {wall - dev_ms:.3f} ms/step")
        pr = cProfile.Profile()
        pr.enable()
        loop(a.steps)
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
        print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
