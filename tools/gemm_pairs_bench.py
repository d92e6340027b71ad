"""CTA-pair (cta_group::2, 256 x 256 tiles) vs single-CTA (128 x 256) GEMM on the large shapes of a step:
python tools/gemm_pairs_bench.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gemm_sweep as S  # noqa: E402

S.PEAK = 1628.6
for M, N, K in ((51200, 1024, 2048), (51200, 512, 2048), (51200, 512, 512), (18432, 1024, 1024), (12288, 1024, 1024), (12288, 512, 512), (6144, 1024, 1024)):
    for tn in (256, 512):
        S.report("fwd", M, N, K, S.fwd(M, N, K, tile_n=tn), f"tile_n={tn}")
for M, N, K in ((51200, 512, 512), (12288, 1024, 1024), (12288, 1024, 512)):
    for tn in (256, 512):
        S.report("dgrad", M, N, K, S.dgrad(M, N, K, tile_n=tn), f"tile_n={tn}")
for M, N, K in ((1024, 2048, 51200), (512, 2048, 51200), (512, 512, 51200), (1024, 1024, 12288)):
    for tn in (256, 512):
        us, sp = S.wgrad(M, N, K, tile_n=tn)
        S.report("wgrad", M, N, K, us, f"tile_n={tn} splits={sp}")
