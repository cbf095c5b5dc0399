"""GPU tests (>= 2 GPUs) of the NVLink peer-memory kernels (csrc/p2p_allreduce.cu):
  * the one-shot all-reduce against the sum in rank order with fp32 accumulation, bit for bit and identical on every rank,
    over message sizes from one vector to several chunks per block, back to back (epoch / slot re-use) and replayed from a
    CUDA graph;
  * the fused layer boundary (all-reduce + residual + RMSNorm in one kernel, fed by an fp16 tensor or by a deferred row-parallel
    GEMM's split-K partials) against "all-reduce, then b200_rmsnorm_residual", bit for bit: every normed row on every rank, the
    new residual on the rows the rank owns (t % world == rank: the residual stream of a row lives on its owner);
  * the sharded greedy head ((value, index) exchange) against torch.argmax over the concatenated logits."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

SIZES = [8, 4096, 4104, 64 * 4096, 256 * 4096 - 8, 3 * 8, 512 * 1024]  # fp16 elements; 2 MiB window


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(world, n, it):
    g = torch.Generator().manual_seed(1000 * it + n % 997)
    return [torch.randn(n, generator=g).half() for _ in range(world)]


def _expected(parts):
    acc = torch.zeros(parts[0].numel(), dtype=torch.float32)
    for p in parts:  # rank order, like the kernel
        acc += p.float()
    return acc.half()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      B200_P2P_ALLREDUCE="1")
    import tgis_b200  # noqa: F401
    from tgis_b200.utils.dist import initialize_torch_distributed
    from tgis_b200.utils.p2p import LayerBoundaryAllReduce
    group = initialize_torch_distributed(world, rank)
    reduce = LayerBoundaryAllReduce(group)
    assert reduce.uses_peer_memory
    bad = []
    for it in range(3):
        for n in SIZES:
            parts = _inputs(world, n, it)
            x = parts[rank].cuda()
            reduce(x)
            if not torch.equal(x.cpu(), _expected(parts)):
                bad.append(("eager", it, n))
    # CUDA graph replay: the epoch advances on the device
    n = 64 * 4096
    buf = torch.zeros(n, dtype=torch.float16, device="cuda")
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        reduce(buf)
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            reduce(buf)
            reduce(buf)  # second reduce of the already reduced tensor: world * sum
    for it in range(4):
        parts = _inputs(world, n, 50 + it)
        buf.copy_(parts[rank].cuda())
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        once = _expected(parts)
        twice = _expected([once] * world)
        if not torch.equal(buf.cpu(), twice):
            bad.append(("graph", it, n))
    q.put((rank, bad))
    q.close()
    q.join_thread()
    os._exit(0)


def _run_workers(target, world=2, timeout=240):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=timeout) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    return res


def _boundary_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import ctypes

    import tgis_b200  # noqa: F401
    from tgis_b200 import _lib, ops
    from tgis_b200.utils.dist import initialize_torch_distributed
    from tgis_b200.utils.p2p import FusedBoundary
    group = initialize_torch_distributed(world, rank)
    lib = _lib.load()
    bad = []
    H = 4096
    fb = FusedBoundary(group, H)
    assert fb.norm is not None and fb.argmax is not None
    st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731

    def fused(h, parts, res, gamma):
        T = h.shape[0] if h is not None else parts.shape[0]
        normed = torch.empty(T, H, dtype=torch.float16, device="cuda")
        res_out = torch.full_like(normed, float("nan"))
        _lib.check(lib.b200_p2p_allreduce_rmsnorm(fb.norm, h.data_ptr() if h is not None else None, parts.ref if parts is not None else None,
                                                  res.data_ptr() if res is not None else None, gamma.data_ptr(), normed.data_ptr(),
                                                  res_out.data_ptr(), T, H, 1e-5, st()), "p2p_allreduce_rmsnorm")
        return normed, res_out

    def same(n_got, r_got, n_exp, r_exp):
        """normed rows everywhere; the residual only on this rank's rows, and nothing written on the others"""
        own = torch.arange(n_got.shape[0], device="cuda") % world == rank
        return (torch.equal(n_got, n_exp) and torch.equal(r_got[own], r_exp[own])
                and (r_got is res_out_graph or bool(torch.isnan(r_got[~own]).all())))

    res_out_graph = None

    def expected(h_all, res, gamma):
        acc = torch.zeros_like(h_all[0], dtype=torch.float32)
        for h in h_all:  # rank order, fp32, one rounding: the all-reduced fp16 hidden state
            acc += h.float()
        red = acc.half()
        if res is None:
            normed, _ = ops.rmsnorm_residual(red, None, gamma, 1e-5)
            return normed, red
        return ops.rmsnorm_residual(red, res, gamma, 1e-5)

    for it, T in enumerate([64, 1, 7, 128, 256, 64, 64, 3, 64, 200, 64, 300, 1100, 2048, 257, 64]):
        g = torch.Generator().manual_seed(100 + it)
        h_all = [torch.randn(T, H, generator=g).half().cuda() for _ in range(world)]
        res = torch.randn(T, H, generator=g).half().cuda() if it != 1 else None
        gamma = (1 + 0.1 * torch.randn(H, generator=g)).half().cuda()
        n_got, r_got = fused(h_all[rank], None, res, gamma)
        n_exp, r_exp = expected(h_all, res, gamma)
        torch.cuda.synchronize()
        if not same(n_got, r_got, n_exp, r_exp):
            bad.append(("fp16 input", it, T, int((n_got != n_exp).sum()), int((r_got != r_exp).sum())))
    # fed by a deferred row-parallel GEMM (each rank its own K-slice of the weight): fp16 and int4
    for it, (T, K) in enumerate([(64, 2048), (64, 512), (17, 1376)]):
        g = torch.Generator().manual_seed(200 + it)
        xs = [torch.randn(T, K, generator=g).half().cuda() for _ in range(world)]
        ws = [(torch.randn(H, K, generator=g) * 0.05).half().cuda() for _ in range(world)]
        res = torch.randn(T, H, generator=g).half().cuda()
        gamma = (1 + 0.1 * torch.randn(H, generator=g)).half().cuda()
        h_all = [ops.gemm_f16(x, w) for x, w in zip(xs, ws)]
        # the deferred sum of this rank must equal its own materialised GEMM for the comparison to be bit-exact: use the deferred
        # GEMM's own reduction as the rank-partial reference
        mine = ops.splitk_reduce(ops.gemm_f16_deferred(xs[rank], ws[rank]))
        others = [ops.splitk_reduce(ops.gemm_f16_deferred(x, w)) for x, w in zip(xs, ws)]
        parts = ops.gemm_f16_deferred(xs[rank], ws[rank])
        n_got, r_got = fused(None, parts, res, gamma)
        n_exp, r_exp = expected(others, res, gamma)
        torch.cuda.synchronize()
        if not (torch.equal(others[rank], mine) and same(n_got, r_got, n_exp, r_exp)):
            bad.append(("deferred f16", it, T, K, int((n_got != n_exp).sum()), int((r_got != r_exp).sum())))
    # CUDA graph replay of the fused kernel (epochs advance on the device)
    T = 64
    g = torch.Generator().manual_seed(300)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).half().cuda()
    h_buf = torch.zeros(T, H, dtype=torch.float16, device="cuda")
    res_buf = torch.zeros(T, H, dtype=torch.float16, device="cuda")
    normed = torch.empty_like(h_buf)
    res_out = res_out_graph = torch.empty_like(h_buf)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        def enqueue():
            _lib.check(lib.b200_p2p_allreduce_rmsnorm(fb.norm, h_buf.data_ptr(), None, res_buf.data_ptr(), gamma.data_ptr(), normed.data_ptr(),
                                                      res_out.data_ptr(), T, H, 1e-5, torch.cuda.current_stream().cuda_stream), "fused")
        enqueue()
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            enqueue()
            enqueue()  # back to back in one graph: the second call of a row must not see the first call's cells
    for it in range(7):
        g = torch.Generator().manual_seed(400 + it)
        h_all = [torch.randn(T, H, generator=g).half().cuda() for _ in range(world)]
        res = torch.randn(T, H, generator=g).half().cuda()
        h_buf.copy_(h_all[rank])
        res_buf.copy_(res)
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        n_exp, r_exp = expected(h_all, res, gamma)
        if not same(normed, res_out, n_exp, r_exp):
            bad.append(("graph", it))
    # sharded greedy head: rank r owns global ids [r * V_local, (r + 1) * V_local)
    for it, (B, V_local) in enumerate([(64, 16000), (3, 1001), (256, 4096), (64, 16000)]):
        g = torch.Generator().manual_seed(500 + it)
        full = torch.randn(B, world * V_local, generator=g).half()
        full[0, 5] = full[0, V_local + 7] = 100.0          # a tie across shards: the lowest global id wins
        if B > 2:
            full[2, world * V_local - 1] = 100.0               # last id of the last shard
        banned = torch.full((B,), -1, dtype=torch.int64)
        banned[1] = int(full[1].float().argmax())             # the winner of row 1 is banned (min_new_tokens EOS mask)
        exp_rows = full.float().clone()
        exp_rows[1, banned[1]] = float("-inf")
        exp = exp_rows.argmax(-1)
        local = full[:, rank * V_local:(rank + 1) * V_local].contiguous().cuda()
        out = torch.empty(B, dtype=torch.int64, device="cuda")
        _lib.check(lib.b200_p2p_argmax(fb.argmax, local.data_ptr(), out.data_ptr(), B, V_local, V_local, banned.cuda().data_ptr(), st()),
                   "p2p_argmax")
        torch.cuda.synchronize()
        if not torch.equal(out.cpu(), exp):
            bad.append(("argmax", it, out.cpu()[:4].tolist(), exp[:4].tolist()))
    q.put((rank, bad))
    q.close()
    q.join_thread()
    os._exit(0)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_boundary_and_sharded_argmax(world):
    res = _run_workers(_boundary_worker, world=world)
    assert res == {r: [] for r in range(world)}, str(res)[:3000]


def test_p2p_allreduce_matches_rank_order_sum():
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert res == {0: [], 1: []}, res
