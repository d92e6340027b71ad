"""tile_n = 128 vs 256 on the mid-size shapes of a step: python tools/tile_sweep.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gemm_sweep as S  # noqa: E402

for M in (6144, 12288, 1024):
    for N, K in ((1024, 1024), (512, 1024), (512, 512), (3072, 1024)):
        a = S.fwd(M, N, K, tile_n=128)
        b = S.fwd(M, N, K, tile_n=256)
        c = S.fwd(M, N, K, tile_n=0)
        print(f"fwd   M={M:6d} N={N:5d} K={K:5d}  t128={a:6.1f}  t256={b:6.1f}  auto={c:6.1f}", flush=True)
for M in (6144, 12288):
    for N, K in ((1024, 1024), (1024, 512), (512, 512)):
        a = S.dgrad(M, N, K, tile_n=128)
        b = S.dgrad(M, N, K, tile_n=256)
        c = S.dgrad(M, N, K, tile_n=0)
        print(f"dgrad M={M:6d} N={N:5d} K={K:5d}  t128={a:6.1f}  t256={b:6.1f}  auto={c:6.1f}", flush=True)
