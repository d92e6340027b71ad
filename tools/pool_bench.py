"""Micro-benchmark of the pooling kernels (cti_tri_pool_fwd / _bwd) at the bench shape.

    python tools/pool_bench.py [--rows 1024] [--A 6] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cti_b200  # noqa: E402,F401
from cti_b200 import kernels as K_  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1024)
ap.add_argument("--A", type=int, default=6)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
B, K, Q, A, C = args.rows, 50, 12, args.A, 1024
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
mk = lambda n: (torch.relu(torch.randn(B * n, C, generator=g, device=dev)) * 0.5).to(torch.bfloat16)
v, q, a = mk(K), mk(Q), (mk(A) if A else None)
w = torch.softmax(torch.randn(B, K * Q * max(A, 1), generator=g, device=dev), 1)
dout = torch.randn(B, C, generator=g, device=dev)


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


res = {"rows": B, "A": A}
res["fwd_us"] = timed(lambda: K_.tri_pool_fwd(v, q, a, w, w.stride(0), B, K, Q, A, C), args.iters)
res["bwd_us"] = timed(lambda: K_.tri_pool_bwd(v, q, a, w, w.stride(0), dout, B, K, Q, A, C), args.iters)
byt = (B * K * C * 2) * 2 + B * (Q + A) * C * 2 * 2 + B * K * Q * max(A, 1) * 4 * 2 + B * C * 4
res["bwd_gbs"] = byt / res["bwd_us"] / 1e3
print(json.dumps(res))
