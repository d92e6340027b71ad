"""One GEMM shape, a few launches (for ncu): python tools/gemm_one.py fwd|dgrad|wgrad M N K"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gemm_sweep as S  # noqa: E402

kind, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
S.REPS = 3
us = {"fwd": S.fwd, "dgrad": S.dgrad}[kind](M, N, K) if kind != "wgrad" else S.wgrad(M, N, K)[0]
S.report(kind, M, N, K, us)
