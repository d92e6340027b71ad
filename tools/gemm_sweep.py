"""Time the tcgen05 GEMM on the shapes of one CTI step (back-to-back launches between two CUDA events, so host
latency does not pollute the small ones) and a few diagnostic variants.  Usage: python tools/gemm_sweep.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cti_b200  # noqa: E402,F401
from cti_b200 import kernels as K_  # noqa: E402
from cti_b200.functions import _pick_splits  # noqa: E402

DEV = "cuda"
BF16 = torch.bfloat16
REPS = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20
PEAK = 1375.6


def timeit(fn, reps=None):
    reps = reps or REPS
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3      # us


def fwd(M, N, K, bias=True, relu=True, f32=False, tile_n=0):
    a = torch.randn(M, K, device=DEV).to(BF16)
    w = torch.randn(N, K, device=DEV).to(BF16)
    b = torch.randn(N, device=DEV) if bias else None
    ob = torch.empty(M, N, dtype=BF16, device=DEV)
    of = torch.empty(M, N, dtype=torch.float32, device=DEV) if f32 else None

    def run():
        K_._call("cti_gemm_bf16", K_._lib.load().cti_gemm_bf16, (
            a.data_ptr(), K, 0, w.data_ptr(), K, 0, M, N, K, 1.0, K_._ptr(b), int(relu), None, 0,
            None if f32 else ob.data_ptr(), K_._ptr(of), N, 0, 1, tile_n, K_._stream()))
    return timeit(run)


def dgrad(M, N, K, aux=True, f32=False, tile_n=0):
    """dx[M, N] = dz[M, K] W[K, N]: B operand MN-major."""
    dz = torch.randn(M, K, device=DEV).to(BF16)
    w = torch.randn(K, N, device=DEV).to(BF16)
    y = torch.randn(M, N, device=DEV).to(BF16) if aux else None
    ob = torch.empty(M, N, dtype=BF16, device=DEV)
    of = torch.empty(M, N, dtype=torch.float32, device=DEV) if f32 else None

    def run():
        K_._call("cti_gemm_bf16", K_._lib.load().cti_gemm_bf16, (
            dz.data_ptr(), K, 0, w.data_ptr(), N, 1, M, N, K, 1.0, None, 0, K_._ptr(y), N if aux else 0,
            None if f32 else ob.data_ptr(), K_._ptr(of), N, 0, 1, tile_n, K_._stream()))
    return timeit(run)


def wgrad(M, N, K, splits=None, tile_n=None):
    """dW[M, N] = dz[K, M]^T x[K, N]: both MN-major, split-K with fp32 atomics."""
    dz = torch.randn(K, M, device=DEV).to(BF16)
    x = torch.randn(K, N, device=DEV).to(BF16)
    dw = torch.zeros(M, N, dtype=torch.float32, device=DEV)
    tn = tile_n or (256 if N >= 256 else 128)
    tiles = -(-M // 128) * -(-N // tn)
    sp = splits or _pick_splits(tiles, -(-K // 64))

    def run():
        K_._call("cti_gemm_bf16", K_._lib.load().cti_gemm_bf16, (
            dz.data_ptr(), M, 1, x.data_ptr(), N, 1, M, N, K, 1.0, None, 0, None, 0,
            None, dw.data_ptr(), N, 1, sp, tn, K_._stream()))
    return timeit(run), sp


def report(kind, M, N, K, us, extra=""):
    tf = 2.0 * M * N * K / us / 1e6
    print(f"{kind:6s} M={M:6d} N={N:5d} K={K:6d}  {us:8.1f} us  {tf:7.1f} TF/s  {tf / PEAK:5.2f} of peak  {extra}", flush=True)


def main():
    rows = [(51200, "v"), (12800, "v/4"), (12288, "q"), (6144, "a"), (1024, "prj")]
    print("== forward (bias + ReLU -> bf16)")
    for M, who in rows:
        for N, K in ((512, 2048), (1024, 2048), (512, 1024), (1024, 1024), (512, 512)):
            if (K == 2048) != (who in ("v", "v/4")) and not (who in ("v", "v/4") and K == 512):
                continue
            report("fwd", M, N, K, fwd(M, N, K), who)
    print("== forward diagnostics on M=51200 N=512")
    for K in (64, 128, 256, 512, 1024, 2048):
        report("fwd", 51200, 512, K, fwd(51200, 512, K), "bias+relu")
    report("fwd", 51200, 512, 512, fwd(51200, 512, 512, bias=False, relu=False), "plain bf16 store")
    report("fwd", 51200, 512, 512, fwd(51200, 512, 512, f32=True), "fp32 store")
    report("fwd", 51200, 512, 512, fwd(51200, 512, 512, tile_n=128), "tile_n=128")
    report("fwd", 51200, 512, 64, fwd(51200, 512, 64, tile_n=128), "tile_n=128")
    print("== dgrad (ReLU-mask aux -> bf16)")
    for M, who in rows[2:4] + [(51200, "v")]:
        for N, K in ((1024, 1024), (1024, 512), (512, 512)):
            report("dgrad", M, N, K, dgrad(M, N, K), who)
    report("dgrad", 12288, 1024, 1024, dgrad(12288, 1024, 1024, f32=True, aux=False), "fp32 out (dx of the tucker layer)")
    report("dgrad", 1024, 1024, 1024, dgrad(1024, 1024, 1024, f32=True, aux=False), "prj fp32 out")
    print("== wgrad (split-K atomics)")
    for K, who in rows:
        for M, N in ((1024, 2048), (512, 2048), (1024, 1024), (512, 1024), (512, 512)):
            if (N == 2048) != (who in ("v", "v/4")) and not (who in ("v", "v/4") and N == 512):
                continue
            us, sp = wgrad(M, N, K)
            report("wgrad", M, N, K, us, f"{who} splits={sp}")
    for sp in (1, 2, 4, 8, 16):
        us, _ = wgrad(1024, 1024, 1024, splits=sp)
        report("wgrad", 1024, 1024, 1024, us, f"splits={sp}")
    for sp in (1, 2, 4, 8, 16, 32):
        us, _ = wgrad(1024, 1024, 12288, splits=sp)
        report("wgrad", 1024, 1024, 12288, us, f"splits={sp}")


if __name__ == "__main__":
    main()
