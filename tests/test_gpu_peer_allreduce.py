"""Gradient sum over NVLink peer memory (dp.PeerRegion / GradAllReducer(transport="peer"), csrc/peer.cu): copy-engine
pushes + flag barriers + one owner-side reduction.  Two processes (two GPUs when the box has them, otherwise both on
cuda:0 -- CUDA IPC and the flag protocol are the same, the barriers then wait for the driver's time slices): the sums must
equal the sequentially added fp32 values bit for bit on every rank, eagerly and replayed from a CUDA graph."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _data(rank, n, step):
    g = torch.Generator().manual_seed(100 * step + rank)
    return torch.randn(n, generator=g)


def _worker(rank, world, port, out):
    import cti_b200  # noqa: F401
    from cti_b200.dp import GradAllReducer, PeerRegion
    ndev = torch.cuda.device_count()
    dev = torch.device("cuda", rank % ndev)
    torch.cuda.set_device(dev)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = {}
    # ---- raw region: ranges of odd sizes (empty chunks, ragged last chunk), several rounds
    n = 4 * 1000 + 4
    reg = PeerRegion(n, dev, timeout_s=30.0)
    for step, (o, m) in enumerate([(0, n), (4, 8), (400, 2000), (0, 4)]):
        reg.slab.copy_(_data(rank, n, step))
        reg.stream.wait_stream(torch.cuda.current_stream())
        reg.all_reduce(o, m)
        torch.cuda.current_stream().wait_stream(reg.stream)
        torch.cuda.synchronize()
        reg.check()
        res[("raw", step)] = reg.slab.cpu().clone()
    # ---- the one-kernel exchange (stores over NVLink, flag rounds inside the kernel), eager and mixed with the copy-engine one
    for step, (o, m) in enumerate([(0, n), (8, 4), (1200, 2000), (0, n)]):
        reg.slab.copy_(_data(rank, n, 40 + step))
        if step == 3:                                     # a copy-engine exchange of another range in flight next to it
            reg.stream.wait_stream(torch.cuda.current_stream())
            reg.all_reduce(0, 1000)
            o, m = 1000, n - 1000
        reg.all_reduce_fused(o, m)
        torch.cuda.current_stream().wait_stream(reg.stream)
        torch.cuda.synchronize()
        reg.check()
        res[("fused", step)] = reg.slab.cpu().clone()
    # ---- replayed from a graph: the barrier epochs live on the device
    src = torch.zeros(n, device=dev)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            reg.slab.copy_(src)
            reg.stream.wait_stream(torch.cuda.current_stream())
            reg.all_reduce(0, 2000)
            reg.all_reduce_fused(2000, n - 2000)
            torch.cuda.current_stream().wait_stream(reg.stream)
    torch.cuda.synchronize()
    dist.barrier()
    for step in range(10, 13):
        src.copy_(_data(rank, n, step))
        graph.replay()
        torch.cuda.synchronize()
        reg.check()
        res[("graph", step)] = reg.slab.cpu().clone()
    # ---- through the reducer: explicit in-place buckets launched early + the rest, hook-free path
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in ((33, 7), (128,), (5, 5, 5), (64, 64), (3,))]
    red = GradAllReducer(ps, param_groups=[[ps[3]], [ps[1], ps[2]]], transport="peer")
    red.set_hooks_enabled(False)
    red.mark_in_place([ps[3], ps[1], ps[2]])
    views = {p: v for b in red.buckets for p, v in zip(b.params, b.views)}
    for step in range(20, 22):
        gs = [_data(rank, p.numel(), step * 10 + i).view_as(p).to(dev) for i, p in enumerate(ps)]
        for gi, grp in enumerate([[3], [1, 2]]):
            for i in grp:
                views[ps[i]].copy_(gs[i])
            red.launch_bucket(gi)
        ps[0].grad, ps[4].grad = gs[0], gs[4]
        red.reduce_now()
        torch.cuda.synchronize()
        red.peer.check()
        res[("red", step)] = [p.grad.cpu().clone() for p in ps]
    torch.save(res, f"{out}.{rank}")
    dist.barrier()
    reg.close()                                           # every rank is past its last exchange: unmap and free
    dist.barrier()
    dist.destroy_process_group()


def _total(n, step, world):
    want = _data(0, n, step)
    for r in range(1, world):                 # rank order, as the owner of an element adds its copies
        want = want + _data(r, n, step)
    return want


@pytest.mark.parametrize("world", [2, 3])      # 2: whole-range exchange with credits; 3: chunked two-phase exchange
def test_peer_memory_sum_is_exact_on_every_rank(tmp_path, world):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    out = str(tmp_path / "r.pt")
    port = 32100 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = [torch.load(f"{out}.{r}", weights_only=False) for r in range(world)]
    n = 4 * 1000 + 4
    for step, (o, m) in enumerate([(0, n), (4, 8), (400, 2000), (0, 4)]):
        d = [_data(r, n, step) for r in range(world)]
        for r in range(world):
            want = d[r].clone()
            want[o:o + m] = _total(n, step, world)[o:o + m]
            assert torch.equal(got[r][("raw", step)], want), (step, r)
    for step, (o, m) in enumerate([(0, n), (8, 4), (1200, 2000), (0, n)]):
        d = [_data(r, n, 40 + step) for r in range(world)]
        for r in range(world):
            want = d[r].clone()
            want[o:o + m] = _total(n, 40 + step, world)[o:o + m]
            assert torch.equal(got[r][("fused", step)], want), ("fused", step, r)
    for step in range(10, 13):
        want = _total(n, step, world)
        for r in range(world):
            assert torch.equal(got[r][("graph", step)], want), (step, r)
    sizes = (33 * 7, 128, 125, 64 * 64, 3)
    for step in range(20, 22):
        for i, m in enumerate(sizes):
            want = _total(m, step * 10 + i, world)
            for r in range(world):
                assert torch.equal(got[r][("red", step)][i].flatten(), want), (step, i, r)
