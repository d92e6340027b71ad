"""ctypes binding of libcti_sm100.so (C ABI declared in include/cti_sm100.h).

This is the only place the shared library is opened.  There is no fallback: if the
library is missing or a call fails, a RuntimeError is raised (the reference trainer's
OOM-skip logic, src/MC/trainer.py:184-190, relies on RuntimeError).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcti_sm100.so")

# name -> (restype, argtypes); mirrors include/cti_sm100.h one to one.
_P = c_void_p


class GemmDesc(ctypes.Structure):
    """cti_gemm_desc of include/cti_sm100.h."""
    _fields_ = [("a", _P), ("lda", c_int), ("a_mn_major", c_int), ("b", _P), ("ldb", c_int), ("b_mn_major", c_int),
                ("M", c_int), ("N", c_int), ("K", c_int), ("alpha", c_float), ("bias", _P), ("relu", c_int),
                ("relu_aux", _P), ("ld_aux", c_int), ("out_bf16", _P), ("out_f32", _P), ("ldc", c_int),
                ("atomic_f32", c_int), ("k_splits", c_int), ("tile_n", c_int)]


class RankProjProblem(ctypes.Structure):
    """cti_rank_proj_problem of include/cti_sm100.h."""
    _fields_ = [("y", _P), ("w_eff", _P), ("bias", _P), ("out", _P), ("dz", _P), ("dzt", _P), ("dw_accum", _P),
                ("M", c_int64), ("p", c_float), ("seed", ctypes.c_uint64), ("site", ctypes.c_uint64)]


SIGNATURES = {
    "cti_version": (c_int, []),
    "cti_last_error": (c_char_p, []),
    "cti_cast_rows_mask": (c_int, [_P, _P, _P, c_int64, c_int, _P]),
    "cti_rowmask_bf16": (c_int, [_P, _P, c_int64, c_int, _P]),
    "cti_cast_rows_dropout": (c_int, [_P, _P, _P, c_int64, c_int, c_float, ctypes.c_uint64, ctypes.c_uint64, _P]),
    "cti_dropout_f32": (c_int, [_P, c_int64, c_float, ctypes.c_uint64, ctypes.c_uint64, _P]),
    "cti_dropout_bf16": (c_int, [_P, _P, c_int64, c_float, ctypes.c_uint64, ctypes.c_uint64, _P]),
    "cti_dropout_expand": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, c_float, ctypes.c_uint64, ctypes.c_uint64, _P]),
    "cti_dropout_reduce": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, c_float, ctypes.c_uint64, ctypes.c_uint64, _P]),
    "cti_sum_row_groups": (c_int, [_P, _P, c_int64, c_int, c_int64, _P]),
    "cti_grad_sumsq_multi": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P, _P]),
    "cti_adamax_multi": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, _P, c_float, c_float, c_float, c_float, c_float,
                                 c_float, _P, _P]),
    "cti_gru_gate_fwd": (c_int, [_P, c_int64, _P, _P, c_int64, _P, c_int64, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "cti_gru_gate_bwd": (c_int, [_P, _P, c_int64, _P, c_int64, _P, _P, _P, _P, _P, c_int64, _P, c_int64, c_int, _P]),
    "cti_gemm_bf16_pair": (c_int, [ctypes.POINTER(GemmDesc), ctypes.POINTER(GemmDesc), _P]),
    "cti_wn_pack_multi": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, c_int, _P, _P]),
    "cti_wn_grad_multi": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, c_int, _P, _P]),
    "cti_wn_scratch_floats": (c_size_t, [c_int, c_int, c_int]),
    "cti_wn_pack": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "cti_wn_grad": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "cti_gemm_bf16": (c_int, [_P, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_float, _P, c_int, _P, c_int,
                              _P, _P, c_int, c_int, c_int, c_int, _P]),
    "cti_act_bwd_bias": (c_int, [_P, c_int, _P, _P, _P, c_int64, c_int, _P]),
    "cti_kd_loss": (c_int, [_P, _P, c_int, _P, _P, _P, _P, c_int, c_int, c_float, c_float, _P]),
    "cti_masked_softmax_fwd": (c_int, [_P, _P, c_int64, c_int, _P]),
    "cti_masked_softmax_bwd": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, c_int64, c_int, c_int, _P]),
    "cti_trilinear_logits_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "cti_trilinear_n1_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "cti_trilinear_logits_bwd_workspace": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "cti_trilinear_logits_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, c_int, c_int,
                                         c_int, c_int, c_int, c_int, c_int, _P]),
    "cti_tri_pool_fwd": (c_int, [_P, _P, _P, _P, c_int64, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "cti_tri_pool_bwd": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int,
                                 c_int, c_int, _P]),
    "cti_tri_pool_bwd_strided": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int,
                                         c_int, c_int, c_int, _P]),
    "cti_rank_proj_dropout_scale": (c_float, [c_float]),
    "cti_rank_proj_dropout_fwd": (c_int, [ctypes.POINTER(RankProjProblem), c_int, c_int, c_int, _P]),
    "cti_rank_proj_dropout_dgrad": (c_int, [ctypes.POINTER(RankProjProblem), c_int, c_int, c_int, _P]),
    "cti_rank_proj_dropout_wgrad": (c_int, [ctypes.POINTER(RankProjProblem), c_int, c_int, c_int, _P]),
    "cti_rank_proj_dropout_mask": (c_int, [_P, c_int64, c_int, c_int, c_float, ctypes.c_uint64, ctypes.c_uint64, _P]),
    "cti_glimpse_residual_cast": (c_int, [_P, c_int, _P, c_int, _P, _P, c_int, _P, c_int, _P, c_int, c_int64, c_int, _P]),
    "cti_glimpse_token_sum": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, c_int, c_int, _P, _P, c_int64, c_int, _P]),
    "cti_glimpse_bcast_rows": (c_int, [_P, _P, c_int, _P, c_int, c_int64, c_int, _P]),
    "cti_bilinear_logits_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "cti_bilinear_logits_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "cti_peer_alloc": (c_int, [c_size_t, ctypes.POINTER(_P)]),
    "cti_peer_free": (c_int, [_P]),
    "cti_peer_export": (c_int, [_P, _P]),
    "cti_peer_import": (c_int, [_P, ctypes.POINTER(_P)]),
    "cti_peer_close": (c_int, [_P]),
    "cti_peer_barrier": (c_int, [ctypes.POINTER(_P), c_int, c_int, c_int, ctypes.c_double, _P]),
    "cti_peer_error": (c_int, [_P, ctypes.POINTER(c_int)]),
    "cti_peer_barrier_memops": (c_int, [ctypes.POINTER(_P), c_int, c_int, c_int, _P]),
    "cti_peer_flag_ops": (c_int, [_P, ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(c_int), c_int, _P]),
    "cti_peer_flag_op": (c_int, [_P, c_int, ctypes.c_uint32, c_int, _P]),
    "cti_peer_stamp": (c_int, [_P, _P]),
    "cti_peer_allreduce_fused": (c_int, [ctypes.POINTER(_P), ctypes.POINTER(_P), ctypes.POINTER(_P), c_int, c_int, c_int64,
                                         ctypes.c_double, _P]),
    "cti_peer_copy": (c_int, [_P, _P, c_size_t, _P]),
    "cti_sum_staged": (c_int, [_P, _P, c_int, c_int, c_int64, c_int64, _P]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Open libcti_sm100.so (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CTI sm_100a kernels are not built. Run "
            "`python __graft_entry__.py build` (or iccv19_vqa-cti_b200/build.py); there is no CPU/PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().cti_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        kind = "argument error" if rc < 0 else f"CUDA error {rc}"
        raise RuntimeError(f"{what}: {kind}: {last_error()}")
