"""Latency of the tensor-parallel layer boundary, back to back in a CUDA graph (no GEMM in between, so no rank skew):
  fused   b200_p2p_allreduce_rmsnorm            (row-owner reduce + norm + broadcast over NVLink windows, one kernel)
  oneshot b200_p2p_allreduce_f16 + b200_rmsnorm_residual   (pull all-reduce with flags, then the norm kernel)
  nccl    torch.distributed.all_reduce + b200_rmsnorm_residual   (what the reference does, utils/layers.py:318-322)

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_boundary.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    import tgis_b200  # noqa: F401
    from tgis_b200 import _lib, ops
    from tgis_b200.utils.dist import initialize_torch_distributed
    from tgis_b200.utils.p2p import FusedBoundary, LayerBoundaryAllReduce
    group = initialize_torch_distributed(world, rank)
    lib = _lib.load()
    reps = 40
    for T, H in [(64, 4096), (128, 8192), (1, 4096), (256, 8192), (512, 8192), (1100, 8192), (2048, 8192), (2048, 4096)]:
        fb = FusedBoundary(group, H)
        one = LayerBoundaryAllReduce(group, max_bytes=2048 * H * 2)
        g = torch.Generator().manual_seed(rank)
        h = torch.randn(T, H, generator=g).half().cuda()
        res = torch.randn(T, H, generator=g).half().cuda()
        gamma = torch.ones(H, dtype=torch.float16, device="cuda")
        normed = torch.empty_like(h)
        res_out = torch.empty_like(h)
        tmp = torch.empty_like(h)

        def fused():
            _lib.check(lib.b200_p2p_allreduce_rmsnorm(fb.norm, h.data_ptr(), None, res.data_ptr(), gamma.data_ptr(), normed.data_ptr(),
                                                      res_out.data_ptr(), T, H, 1e-5, torch.cuda.current_stream().cuda_stream), "fused")

        def oneshot():
            tmp.copy_(h)
            one(tmp)
            ops.rmsnorm_residual(tmp, res, gamma, 1e-5)

        def nccl():
            tmp.copy_(h)
            dist.all_reduce(tmp, group=group)
            ops.rmsnorm_residual(tmp, res, gamma, 1e-5)

        out = {}
        for name, fn in (("fused", fused), ("oneshot", oneshot), ("nccl", nccl)):
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                for _ in range(3):
                    fn()
                side.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    for _ in range(reps):
                        fn()
                dist.barrier()
                for _ in range(3):
                    graph.replay()
                side.synchronize()
                dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(side)
                for _ in range(5):
                    graph.replay()
                e1.record(side)
                side.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / (5 * reps) * 1e3], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[name] = float(t.item())
        if rank == 0:
            print(f"boundary tp{world} T={T:4d} H={H:5d}: fused {out['fused']:6.2f} us   one-shot + norm {out['oneshot']:6.2f} us   "
                  f"nccl + norm {out['nccl']:6.2f} us  (the last two include a {T * H * 2 // 1024} KB device copy)", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)  # the IPC windows stay mapped in the peers: no orderly teardown to wait for


if __name__ == "__main__":
    main()
