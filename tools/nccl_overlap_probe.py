"""Can an ASYNC all-reduce issued from inside an autograd backward (the autograd engine's worker thread) be captured in a
CUDA graph next to the compute that follows it, and does it overlap on replay?

    timeout 180 torchrun --nproc-per-node 2 tools/nccl_overlap_probe.py

Prints per-replay times of: compute alone, compute + all-reduce issued at the END, compute + all-reduce issued from the
MIDDLE of backward (async, waited at the end).  Always run under `timeout`."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 1024
slab = torch.zeros(16 << 20, device=dev)                 # 64 MB of "gradients"
half = slab[: 8 << 20]
x0 = torch.randn(8192, N, device=dev, dtype=torch.bfloat16)
ws = [torch.randn(N, N, device=dev, dtype=torch.bfloat16, requires_grad=True) for _ in range(24)]
pending = []


class Mark(torch.autograd.Function):
    """identity; its backward launches the async all-reduce of `half` (what a bucket-ready callback would do)"""
    @staticmethod
    def forward(ctx, x, mode):
        ctx.mode = mode
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if ctx.mode == "mid":
            pending.append(dist.all_reduce(half, async_op=True))
        return g, None


def step(mode):
    for w in ws:
        w.grad = None
    h = x0
    for i, w in enumerate(ws):
        h = torch.relu(h @ w)
        if i == 11:
            h = Mark.apply(h, mode)
    h.float().sum().backward()
    if mode == "end":
        dist.all_reduce(half)
    for wk in pending:
        wk.wait()
    pending.clear()


for mode in ("none", "end", "mid"):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step(mode)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    t0 = time.time()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        step(mode)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(rank, mode, "captured in %.2fs, replay %.3f ms" % (time.time() - t0, e0.elapsed_time(e1) / 20), flush=True)
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
