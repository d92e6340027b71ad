"""Timing of the peer-memory gradient sum's pieces (dp.PeerRegion): copy-engine push bandwidth, barrier latency, the whole
all_reduce at the hot path's slab size, and NCCL's all-reduce of the same bytes beside it.

    timeout 120 torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_probe.py"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cti_b200  # noqa: E402,F401
from cti_b200 import _lib  # noqa: E402
from cti_b200.dp import PeerRegion  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W = dist.get_world_size()
N = 15_750_000 // 4 * 4
reg = PeerRegion(N, dev)
lib = _lib.load()


def timed(fn, iters=20, stream=None):
    stream = stream or torch.cuda.current_stream()
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


peer = (rank + 1) % W
for mb in (1, 4, 32):
    nb = mb << 20
    t = timed(lambda: _lib.check(lib.cti_peer_copy(reg._staging(peer, 0), reg._slab(rank, 0), nb, reg.stream.cuda_stream), "copy"),
              stream=reg.stream)
    print(rank, f"push {mb} MB: {t * 1e3:.1f} us  {nb / t / 1e6:.0f} GB/s", flush=True)
    t = timed(lambda: _lib.check(lib.cti_peer_copy(reg._staging(rank, 0), reg._slab(peer, 0), nb, reg.stream.cuda_stream), "copy"),
              stream=reg.stream)
    print(rank, f"pull {mb} MB: {t * 1e3:.1f} us  {nb / t / 1e6:.0f} GB/s", flush=True)
t = timed(lambda: reg.barrier(), stream=reg.stream)
print(rank, f"barrier: {t * 1e3:.1f} us", flush=True)
for n in (N, N // 4, 1 << 16):
    t = timed(lambda: reg.all_reduce(0, n), stream=reg.stream)
    print(rank, f"peer all_reduce {4 * n / 1e6:.1f} MB: {t * 1e3:.1f} us", flush=True)
    x = torch.zeros(n, device=dev)
    t = timed(lambda: dist.all_reduce(x))
    print(rank, f"nccl all_reduce {4 * n / 1e6:.1f} MB: {t * 1e3:.1f} us", flush=True)
torch.cuda.synchronize()
reg.check()
dist.barrier()
os._exit(0)
