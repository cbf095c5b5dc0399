"""GPU test (>= 2 GPUs, opt-in with B200_EXPERIMENTAL=1 until validated): the one-shot peer-memory all-reduce
(csrc/p2p_allreduce.cu) against the sum in rank order with fp32 accumulation, bit for bit and identical on every rank,
over message sizes from one vector to several chunks per block, back to back (epoch / slot re-use) and replayed from a
CUDA graph."""
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


@pytest.mark.skipif(os.environ.get("B200_EXPERIMENTAL") != "1", reason="experimental: set B200_EXPERIMENTAL=1")
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
