"""Thin tensor-level wrappers over the C ABI of libcti_sm100.so.

Every function takes CUDA torch tensors, checks dtype / contiguity, allocates the outputs from
torch's caching allocator (the library allocates nothing) and launches on the current stream.
PyTorch is plumbing here: device memory and streams only.  No function has a fallback path.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32
NUM_SMS = 148


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream (the raw getter avoids ~15 us of Python per launch)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


class _Stats:
    """Launch accounting (bench.py's ``gpu_launches``) and optional per-call CUDA-event timing."""
    launches = 0          # CUDA kernels launched through the C ABI since the last reset
    prof = None           # None, or a list receiving (name, tag, flops, bytes, start_event, end_event)


STATS = _Stats()


# GEMMs issued inside ``with gemm_batch():`` are queued (their outputs are allocated and returned at once) and launched
# when the block ends -- or earlier, as soon as anything that could depend on them is launched: any other kernel of the
# library, or a GEMM that reads a queued GEMM's output.  Queued GEMMs are therefore mutually independent, and those
# with the same operand layouts go out two per launch (cti_gemm_bf16_pair).
_PENDING = [None]          # None outside a batch, else the list of queued GEMMs


class gemm_batch:
    def __enter__(self):
        self.outer = _PENDING[0] is not None
        if not self.outer:
            _PENDING[0] = []
        return self

    def __exit__(self, *exc):
        if not self.outer:
            try:
                if exc[0] is None:
                    _flush_gemms()
            finally:
                _PENDING[0] = None
        return False


class gemm_unbatched:
    """Inside a gemm_batch: launch what is queued and run the enclosed GEMMs immediately (code that hands GEMM outputs to
    torch ops, which the queue cannot see)."""
    def __enter__(self):
        self.was = _PENDING[0] is not None
        if self.was:
            _flush_gemms()
            _PENDING[0] = None
        return self

    def __exit__(self, *exc):
        if self.was:
            _PENDING[0] = []
        return False


def _flush_gemms() -> None:
    q = _PENDING[0]
    if not q:
        return
    _PENDING[0] = None                       # the launches below go straight out
    try:
        lib = _lib.load()
        groups = {}
        for e in q:
            groups.setdefault((e["desc"].a_mn_major, e["desc"].b_mn_major), []).append(e)
        for es in groups.values():
            es.sort(key=lambda e: e["flops"])                         # smallest first: (a, q) pair up, v goes alone
            i = 0
            while i < len(es):
                if i + 1 < len(es) and (es[i]["desc"].tile_n == es[i + 1]["desc"].tile_n
                                        or 0 in (es[i]["desc"].tile_n, es[i + 1]["desc"].tile_n)):
                    e0, e1 = es[i + 1], es[i]                         # bigger problem first in the tile order
                    _call("cti_gemm_bf16", lib.cti_gemm_bf16_pair, (e0["desc"], e1["desc"], _stream()),
                          flops=e0["flops"] + e1["flops"], tag="pair " + e0["tag"] + " + " + e1["tag"])
                    i += 2
                else:
                    e = es[i]
                    _call("cti_gemm_bf16", lib.cti_gemm_bf16, e["args"], flops=e["flops"], tag=e["tag"])
                    i += 1
    finally:
        _PENDING[0] = []
        q.clear()


def _overlaps(ranges, lo: int, hi: int) -> bool:
    return any(lo < h and l < hi for l, h in ranges)


def _call(name: str, fn, args, kernels: int = 1, flops: float = 0.0, nbytes: float = 0.0, tag: str = "") -> None:
    """One C-ABI call = `kernels` kernel launches on the current stream; algorithmic work for the roofline."""
    if _PENDING[0]:
        _flush_gemms()                       # queued GEMMs go first: this launch may read what they write
    STATS.launches += kernels
    prof = STATS.prof
    if prof is None:
        _lib.check(fn(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    prof.append((name, tag, flops, nbytes, e0, e1))
    _lib.check(rc, name)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the CTI kernels have no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")


def trilinear_min_flops(K: int, Q: int, A: int, G: int, R: int, d: int = 16) -> float:
    """Algorithmic forward FLOPs per row of the trilinear contraction, cheapest order a -> q -> v
    (T_min)."""
    return R * (2.0 * d ** 3 * G * A + 2.0 * d * d * Q * A * G) + 2.0 * K * (R * d) * Q * A * G


def pool_flops(K: int, Q: int, A: int, C: int) -> float:
    """Algorithmic forward FLOPs per row of the attention-weighted pooling."""
    An = max(A, 1)
    return 2.0 * K * Q * An * C + Q * An * C + 2.0 * K * C


# --------------------------------------------------------------------------- #
def cast_rows(x: torch.Tensor, want_mask: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """fp32 (rows, cols) -> bf16, optionally with the zero-row mask of src/attention.py:55."""
    _req(x, F32, "cast_rows.x")
    rows, cols = x.shape
    out = torch.empty((rows, cols), dtype=BF16, device=x.device)
    mask = torch.empty((rows,), dtype=torch.uint8, device=x.device) if want_mask else None
    _call("cti_cast_rows_mask", _lib.load().cti_cast_rows_mask,
          (x.data_ptr(), out.data_ptr(), _ptr(mask), rows, cols, _stream()), nbytes=6.0 * rows * cols)
    return out, mask


def rowmask_bf16(x: torch.Tensor) -> torch.Tensor:
    """Zero-row mask (reference src/attention.py:55) of a bf16 (rows, cols) matrix."""
    _req(x, BF16, "rowmask_bf16.x")
    rows, cols = x.shape
    mask = torch.empty((rows,), dtype=torch.uint8, device=x.device)
    _call("cti_rowmask_bf16", _lib.load().cti_rowmask_bf16, (x.data_ptr(), mask.data_ptr(), rows, cols, _stream()),
          nbytes=2.0 * rows * cols)
    return mask


def cast_rows_dropout(x: torch.Tensor, drop, want_mask: bool = False):
    """bf16(dropout(x)) for fp32 (rows, cols); drop = (p, seed, offset).  The zero-row mask is of the undropped rows."""
    _req(x, F32, "cast_rows_dropout.x")
    rows, cols = x.shape
    out = torch.empty((rows, cols), dtype=BF16, device=x.device)
    mask = torch.empty((rows,), dtype=torch.uint8, device=x.device) if want_mask else None
    _call("cti_cast_rows_dropout", _lib.load().cti_cast_rows_dropout,
          (x.data_ptr(), out.data_ptr(), _ptr(mask), rows, cols, drop[0], drop[1], drop[2], _stream()),
          nbytes=6.0 * rows * cols)
    return out, mask


def dropout_f32_(x: torch.Tensor, drop) -> torch.Tensor:
    """In place x *= keep / (1 - p): backward of cast_rows_dropout with the same (p, seed, offset)."""
    _req(x, F32, "dropout_f32_.x")
    _call("cti_dropout_f32", _lib.load().cti_dropout_f32, (x.data_ptr(), x.numel(), drop[0], drop[1], drop[2], _stream()),
          nbytes=8.0 * x.numel())
    return x


def dropout_bf16(x: torch.Tensor, drop) -> torch.Tensor:
    _req(x, BF16, "dropout_bf16.x")
    out = torch.empty_like(x)
    _call("cti_dropout_bf16", _lib.load().cti_dropout_bf16,
          (x.data_ptr(), out.data_ptr(), x.numel(), drop[0], drop[1], drop[2], _stream()), nbytes=4.0 * x.numel())
    return out


def dropout_expand(x: torch.Tensor, rank_group: int, r0: int, drop) -> torch.Tensor:
    """(M, H) bf16 -> (M, rank_group * H): column block j holds x masked with rank (r0 + j)'s own dropout mask."""
    _req(x, BF16, "dropout_expand.x")
    M, H = x.shape
    xt = torch.empty((M, rank_group * H), dtype=BF16, device=x.device)
    _call("cti_dropout_expand", _lib.load().cti_dropout_expand,
          (x.data_ptr(), xt.data_ptr(), M, H, rank_group, r0, drop[0], drop[1], drop[2], _stream()),
          nbytes=2.0 * M * H * (1 + rank_group))
    return xt


def dropout_reduce_(dxt: torch.Tensor, acc: torch.Tensor, rank_group: int, r0: int, drop) -> None:
    """acc (M, H) fp32 += sum_j dxt[:, j*H:(j+1)*H] * mask_{r0+j}: backward of dropout_expand."""
    _req(dxt, BF16, "dropout_reduce_.dxt")
    _req(acc, F32, "dropout_reduce_.acc")
    M, H = acc.shape
    _call("cti_dropout_reduce", _lib.load().cti_dropout_reduce,
          (dxt.data_ptr(), acc.data_ptr(), M, H, rank_group, r0, drop[0], drop[1], drop[2], _stream()),
          nbytes=2.0 * M * H * rank_group + 8.0 * M * H)


def wn_pack(v: torch.Tensor, g: torch.Tensor, n_groups: int, pad_rows_to: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    """v (n_groups*rows_per_group, cols) fp32, g (n_groups,) fp32 -> (W_eff bf16, sumsq fp32[n_groups]).
    pad_rows_to: the pack gets zero rows up to a multiple of it (output widths that are not a multiple of 8 elements
    cannot be a TMA operand pitch in the backward GEMMs; the padded outputs are exact zeros)."""
    _req(v, F32, "wn_pack.v")
    _req(g, F32, "wn_pack.g")
    rows, cols = v.shape
    rows_p = -(-rows // pad_rows_to) * pad_rows_to
    if rows_p != rows:
        w = torch.zeros((rows_p, cols), dtype=BF16, device=v.device)
    else:
        w = torch.empty((rows, cols), dtype=BF16, device=v.device)
    lib = _lib.load()
    sumsq = torch.empty((lib.cti_wn_scratch_floats(n_groups, rows // n_groups, cols),), dtype=F32, device=v.device)
    _call("cti_wn_pack", lib.cti_wn_pack, (v.data_ptr(), g.data_ptr(), w.data_ptr(), sumsq.data_ptr(), n_groups,
                                           rows // n_groups, cols, _stream()), kernels=2, nbytes=10.0 * rows * cols)
    return w, sumsq[:n_groups]


def wn_grad(dw: torch.Tensor, v: torch.Tensor, g: torch.Tensor, sumsq: torch.Tensor,
            n_groups: int) -> Tuple[torch.Tensor, torch.Tensor]:
    _req(dw, F32, "wn_grad.dw")
    _req(v, F32, "wn_grad.v")
    rows, cols = v.shape
    dv = torch.empty_like(v)
    dg = torch.empty((n_groups,), dtype=F32, device=v.device)
    lib = _lib.load()
    ws = torch.empty((lib.cti_wn_scratch_floats(n_groups, rows // n_groups, cols),), dtype=F32, device=v.device)
    _call("cti_wn_grad", lib.cti_wn_grad,
          (dw.data_ptr(), v.data_ptr(), g.data_ptr(), sumsq.data_ptr(), dv.data_ptr(), dg.data_ptr(), ws.data_ptr(),
           n_groups, rows // n_groups, cols, _stream()), kernels=2, nbytes=20.0 * rows * cols)
    return dv, dg


def _req_rows(t: torch.Tensor, name: str) -> None:
    """2-D bf16 CUDA tensor whose rows are contiguous (a column slice of a wider matrix is fine: pitch = stride(0))."""
    if not t.is_cuda or t.dtype != BF16 or t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"{name}: expected a 2-D bf16 CUDA tensor with unit column stride")


def gemm(a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int, *, a_mn: bool = False, b_mn: bool = False,
         bias: Optional[torch.Tensor] = None, relu: bool = False, relu_aux: Optional[torch.Tensor] = None,
         out_bf16: bool = True, out_f32: bool = False, accum_f32: Optional[torch.Tensor] = None, k_splits: int = 1,
         alpha: float = 1.0, tile_n: int = 0,
         out: Optional[torch.Tensor] = None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """C[M,N] = epilogue(alpha * A . B^T) on the tcgen05 GEMM.  a / b are 2-D bf16 row-major buffers (row pitch =
    stride(0), so column slices work): K-major operands are stored [M or N][K], MN-major ones [K][M or N].
    accum_f32: a zeroed (M,N) fp32 buffer to atomically accumulate into (split-K).
    out: optional preallocated bf16 (M,N) view with unit column stride (e.g. a column block of a wider matrix)."""
    _req_rows(a, "gemm.a")
    _req_rows(b, "gemm.b")
    ldc = N
    if out is not None:
        _req_rows(out, "gemm.out")
        ob, ldc = out, out.stride(0)
    else:
        ob = torch.empty((M, N), dtype=BF16, device=a.device) if (out_bf16 and accum_f32 is None) else None
    of = accum_f32 if accum_f32 is not None else (torch.empty((M, N), dtype=F32, device=a.device) if out_f32 else None)
    if relu_aux is not None:
        _req(relu_aux, BF16, "gemm.relu_aux")
    tag = f"{'wgrad' if a_mn else ('dgrad' if b_mn else 'fwd')} M={M} N={N} K={K}"
    args = (a.data_ptr(), a.stride(0), int(a_mn), b.data_ptr(), b.stride(0), int(b_mn), M, N, K, float(alpha), _ptr(bias),
            int(relu), _ptr(relu_aux), 0 if relu_aux is None else relu_aux.shape[1], _ptr(ob), _ptr(of), ldc,
            int(accum_f32 is not None), int(k_splits), int(tile_n))
    pending = _PENDING[0]
    if pending is None:
        _call("cti_gemm_bf16", _lib.load().cti_gemm_bf16, args + (_stream(),), flops=2.0 * M * N * K, tag=tag)
        return ob, of
    # queued: flush first if this GEMM touches the output of a queued one (reads it, or accumulates into it)
    span = lambda t: (t.data_ptr(), t.data_ptr() + (t.shape[0] - 1) * t.stride(0) * t.element_size() + t.shape[1] * t.element_size()) \
        if t.dim() == 2 else (t.data_ptr(), t.data_ptr() + t.numel() * t.element_size())
    touched = [span(t) for t in (a, b, bias, relu_aux, ob, of) if t is not None]
    if any(_overlaps(e["writes"], lo, hi) for e in pending for lo, hi in touched):
        _flush_gemms()
        pending = _PENDING[0]
    desc = _lib.GemmDesc(*args)
    pending.append({"desc": desc, "args": args + (_stream(),), "flops": 2.0 * M * N * K, "tag": tag,
                    "writes": [span(t) for t in (ob, of) if t is not None],
                    "keep": (a, b, bias, relu_aux, ob, of)})
    return ob, of


def act_bwd_bias(dy: torch.Tensor, y: Optional[torch.Tensor], want_dz: bool,
                 dbias: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """dz = dy * (y > 0) as bf16 (if want_dz); dbias[n] += sum_m dz[m, n]."""
    rows, cols = dy.shape
    if dy.dtype not in (F32, BF16) or not dy.is_contiguous():
        raise RuntimeError("act_bwd_bias.dy: expected contiguous fp32 / bf16")
    dz = torch.empty((rows, cols), dtype=BF16, device=dy.device) if want_dz else None
    _call("cti_act_bwd_bias", _lib.load().cti_act_bwd_bias,
          (dy.data_ptr(), int(dy.dtype == BF16), _ptr(y), _ptr(dz), _ptr(dbias), rows, cols, _stream()),
          nbytes=float(rows) * cols * (dy.element_size() + (2 if y is not None else 0) + (2 if want_dz else 0)))
    return dz


def softmax_fwd(logits: torch.Tensor, rows: int, length: int) -> torch.Tensor:
    _req(logits, F32, "softmax_fwd.logits")
    p = torch.empty_like(logits)
    _call("cti_masked_softmax_fwd", _lib.load().cti_masked_softmax_fwd,
          (logits.data_ptr(), p.data_ptr(), rows, length, _stream()), nbytes=8.0 * rows * length)
    return p


def softmax_bwd(p: torch.Tensor, dp: torch.Tensor, sb: int, sg: int, se: int, batch: int, groups: int,
                length: int) -> torch.Tensor:
    _req(p, F32, "softmax_bwd.p")
    if dp.dtype != F32:
        raise RuntimeError("softmax_bwd.dp: expected fp32")
    dl = torch.empty_like(p)
    _call("cti_masked_softmax_bwd", _lib.load().cti_masked_softmax_bwd,
          (p.data_ptr(), dp.data_ptr(), sb, sg, se, dl.data_ptr(), batch, groups, length, _stream()),
          nbytes=12.0 * batch * groups * length)
    return dl


def sum_row_groups(x: torch.Tensor, rep: int, row_elems: int) -> torch.Tensor:
    """x (groups*rep rows of row_elems) bf16 -> (groups rows) bf16: sum of each run of `rep` rows (fp32 accumulate)."""
    _req(x, BF16, "sum_row_groups.x")
    groups = x.numel() // (rep * row_elems)
    out = torch.empty((groups * row_elems // x.shape[-1], x.shape[-1]), dtype=BF16, device=x.device)
    _call("cti_sum_row_groups", _lib.load().cti_sum_row_groups,
          (x.data_ptr(), out.data_ptr(), groups, rep, row_elems, _stream()), nbytes=2.0 * (rep + 1) * groups * row_elems)
    return out


_PERM_IDX = {}


def tpack_perm(tpack: torch.Tensor) -> Optional[torch.Tensor]:
    """The packed core (R, 16, 16*G*16) with its last axis in the accumulator-lane order of the tcgen05 forward kernel
    (include/cti_sm100.h): position (j % 4) * 128 + g * 64 + i * 4 + j / 4 holds element i * 32 + g * 16 + j.
    G = 2 only (None otherwise: the generic kernel reads the plain pack)."""
    if tpack.shape[-1] != 512:
        return None
    idx = _PERM_IDX.get(tpack.device)
    if idx is None:
        x = torch.arange(512)
        t, g, i, jj = x >> 7, (x >> 6) & 1, (x >> 2) & 15, x & 3
        idx = _PERM_IDX[tpack.device] = (i * 32 + g * 16 + 4 * jj + t).to(tpack.device)
    return tpack.index_select(-1, idx)


def trilinear_fwd(vc, qc, ac, tpack, rowmask, B, K, Q, A, G, R, v_rep: int = 1, tpack_p=None, save_n1: bool = False):
    """vc (and rowmask) hold B / v_rep samples: rows b*v_rep .. b*v_rep + v_rep - 1 of qc / ac share image b.
    tpack_p: tpack_perm(tpack) if the caller already has it (built here otherwise).
    save_n1 (training): returns (logits, n1) -- n1 is the kernel's bf16 intermediate N1 = T x_l Ac (None for shapes outside
    the tcgen05 path); trilinear_bwd(..., n1=n1) then runs the fast backward without recomputing it."""
    for t, n in ((vc, "vc"), (qc, "qc"), (ac, "ac"), (tpack, "tpack")):
        _req(t, BF16, "trilinear_fwd." + n)
    if tpack_p is None:
        tpack_p = tpack_perm(tpack)
    logits = torch.empty((B, G, K, Q, A), dtype=F32, device=vc.device)
    lib = _lib.load()
    n1 = None
    if save_n1:
        nbytes = lib.cti_trilinear_n1_bytes(B, K, Q, A, G, R)
        if nbytes:
            n1 = torch.empty((nbytes,), dtype=torch.uint8, device=vc.device)
    _call("cti_trilinear_logits_fwd", lib.cti_trilinear_logits_fwd,
          (vc.data_ptr(), qc.data_ptr(), ac.data_ptr(), tpack.data_ptr(), _ptr(tpack_p), _ptr(rowmask), logits.data_ptr(),
           _ptr(n1), B, K, Q, A, G, R, v_rep, _stream()), flops=float(B) * trilinear_min_flops(K, Q, A, G, R))
    return (logits, n1) if save_n1 else logits


def trilinear_bwd(vc, qc, ac, tpack, dlogits, B, K, Q, A, G, R, v_rep: int = 1, n1=None):
    """Returns dzv, dzq, dza (bf16, pre-activation), dbv, dbq, dba (fp32, R*16), dtpack (fp32).
    v_rep > 1: the per-row dzv is folded onto the B / v_rep shared images before it is returned.
    n1: the tensor trilinear_fwd(..., save_n1=True) returned (selects the tcgen05 backward; None: generic kernel)."""
    _req(dlogits, F32, "trilinear_bwd.dlogits")
    lib = _lib.load()
    dev = vc.device
    dzv, dzq, dza = torch.empty((B * K, R * 16), dtype=BF16, device=dev), torch.empty_like(qc), torch.empty_like(ac)
    zeros = torch.zeros((3 * R * 16 + tpack.numel(),), dtype=F32, device=dev)
    dbv, dbq, dba = zeros[:R * 16], zeros[R * 16:2 * R * 16], zeros[2 * R * 16:3 * R * 16]
    dtpack = zeros[3 * R * 16:]
    nbytes = lib.cti_trilinear_logits_bwd_workspace(B, K, Q, A, G, R)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    _call("cti_trilinear_logits_bwd", lib.cti_trilinear_logits_bwd,
          (vc.data_ptr(), qc.data_ptr(), ac.data_ptr(), tpack.data_ptr(), dlogits.data_ptr(), _ptr(n1), dzv.data_ptr(),
           dzq.data_ptr(), dza.data_ptr(), dbv.data_ptr(), dbq.data_ptr(), dba.data_ptr(), dtpack.data_ptr(),
           ws.data_ptr(), nbytes, B, K, Q, A, G, R, v_rep, _stream()), kernels=2,
          flops=2.0 * B * trilinear_min_flops(K, Q, A, G, R))
    if v_rep > 1:
        dzv = sum_row_groups(dzv, v_rep, K * R * 16)
    return dzv, dzq, dza, dbv, dbq, dba, dtpack


def tri_pool_fwd(v, q, a, w, w_stride_b, B, K, Q, A, C, v_rep: int = 1) -> torch.Tensor:
    out = torch.empty((B, C), dtype=F32, device=v.device)
    _call("cti_tri_pool_fwd", _lib.load().cti_tri_pool_fwd,
          (v.data_ptr(), q.data_ptr(), _ptr(a), w.data_ptr(), w_stride_b, out.data_ptr(), B, K, Q, A, C, v_rep, _stream()),
          flops=float(B) * pool_flops(K, Q, A, C))
    return out


def tri_pool_bwd(v, q, a, w, w_stride_b, dout, B, K, Q, A, C, v_rep: int = 1, dw_out: Optional[torch.Tensor] = None):
    """Returns dzv, dzq, dza (bf16), dbv, dbq, dba (fp32, C), dw (B,K,Q[,A]) fp32.  A == 0: bilinear.
    dw_out: optional fp32 view the attention gradient is written into: row b at dw_out.stride(0), contiguous inside a row
    (the glimpse slice of a (B, G, K*Q*A) buffer)."""
    _req(dout, F32, "tri_pool_bwd.dout")
    dev = v.device
    dzv, dzq = torch.empty((B * K, C), dtype=BF16, device=dev), torch.empty_like(q)
    dza = torch.empty_like(a) if A > 0 else None
    zeros = torch.zeros((3 * C,), dtype=F32, device=dev)
    dbv, dbq, dba = zeros[:C], zeros[C:2 * C], zeros[2 * C:]
    if dw_out is None:
        dw = torch.empty((B, K, Q, A) if A > 0 else (B, K, Q), dtype=F32, device=dev)
        _call("cti_tri_pool_bwd", _lib.load().cti_tri_pool_bwd,
              (v.data_ptr(), q.data_ptr(), _ptr(a), w.data_ptr(), w_stride_b, dout.data_ptr(), dzv.data_ptr(), dzq.data_ptr(),
               _ptr(dza), dbv.data_ptr(), dbq.data_ptr(), dba.data_ptr(), dw.data_ptr(), B, K, Q, A, C, v_rep, _stream()),
              kernels=2, flops=2.0 * B * pool_flops(K, Q, A, C))
    else:
        dw = dw_out
        if dw.dtype != F32 or dw.shape[0] != B or dw[0].numel() != K * Q * max(A, 1) or not dw[0].is_contiguous():
            raise RuntimeError("tri_pool_bwd.dw_out: expected an fp32 (B, K*Q*A) view with contiguous rows")
        _call("cti_tri_pool_bwd", _lib.load().cti_tri_pool_bwd_strided,
              (v.data_ptr(), q.data_ptr(), _ptr(a), w.data_ptr(), w_stride_b, dout.data_ptr(), dzv.data_ptr(), dzq.data_ptr(),
               _ptr(dza), dbv.data_ptr(), dbq.data_ptr(), dba.data_ptr(), dw.data_ptr(), dw.stride(0), B, K, Q, A, C, v_rep,
               _stream()), kernels=2, flops=2.0 * B * pool_flops(K, Q, A, C))
    if v_rep > 1:
        dzv = sum_row_groups(dzv, v_rep, K * C)
    return dzv, dzq, dza, dbv, dbq, (dba if A > 0 else None), dw


def bilinear_fwd(vb, qb, hmat, hbias, rowmask, B, K, Q, G, C) -> torch.Tensor:
    _req(hmat, F32, "bilinear_fwd.hmat")
    _req(hbias, F32, "bilinear_fwd.hbias")
    logits = torch.empty((B, G, K, Q), dtype=F32, device=vb.device)
    _call("cti_bilinear_logits_fwd", _lib.load().cti_bilinear_logits_fwd,
          (vb.data_ptr(), qb.data_ptr(), hmat.data_ptr(), hbias.data_ptr(), _ptr(rowmask), logits.data_ptr(), B, K, Q, G,
           C, _stream()), flops=2.0 * B * K * Q * G * C)
    return logits


def bilinear_bwd(vb, qb, hmat, dlogits, B, K, Q, G, C):
    """Returns dzv, dzq (bf16), dbv, dbq (C), dhmat (G,C), dhbias (G)."""
    _req(dlogits, F32, "bilinear_bwd.dlogits")
    dev = vb.device
    dzv, dzq = torch.empty_like(vb), torch.empty_like(qb)
    zeros = torch.zeros((2 * C + G * C + G,), dtype=F32, device=dev)
    dbv, dbq = zeros[:C], zeros[C:2 * C]
    dh = zeros[2 * C:2 * C + G * C].view(G, C)
    dhb = zeros[2 * C + G * C:]
    _call("cti_bilinear_logits_bwd", _lib.load().cti_bilinear_logits_bwd,
          (vb.data_ptr(), qb.data_ptr(), hmat.data_ptr(), dlogits.data_ptr(), dzv.data_ptr(), dzq.data_ptr(),
           dbv.data_ptr(), dbq.data_ptr(), dh.data_ptr(), dhb.data_ptr(), B, K, Q, G, C, _stream()),
          flops=4.0 * B * K * Q * G * C)
    return dzv, dzq, dbv, dbq, dh, dhb


# --------------------------------------------------------------------------- #
# GRU timestep, pointwise part (the products are gemm() calls)
# --------------------------------------------------------------------------- #
def gru_gate_fwd(gx_t: torch.Tensor, gh: torch.Tensor, h_prev: Optional[torch.Tensor], h_out_t: torch.Tensor,
                 h_bf16: torch.Tensor, r: torch.Tensor, z: torch.Tensor, n: torch.Tensor, ghn: torch.Tensor) -> None:
    """gx_t (B, 3H) fp32 view with any row stride, gh (B, 3H) fp32 contiguous, h_prev / h_out_t (B, H) fp32 views with
    any row stride (h_prev None = zeros); h_bf16, r, z, n, ghn (B, H) bf16 contiguous outputs."""
    B, H = h_bf16.shape
    for t, name in ((gx_t, "gx"), (h_out_t, "h_out")) + (((h_prev, "h_prev"),) if h_prev is not None else ()):
        if t.dtype != F32 or t.stride(1) != 1:
            raise RuntimeError(f"gru_gate_fwd.{name}: expected fp32 with unit column stride")
    _req(gh, F32, "gru_gate_fwd.gh")
    _call("cti_gru_gate_fwd", _lib.load().cti_gru_gate_fwd,
          (gx_t.data_ptr(), gx_t.stride(0), gh.data_ptr(), _ptr(h_prev), 0 if h_prev is None else h_prev.stride(0),
           h_out_t.data_ptr(), h_out_t.stride(0), h_bf16.data_ptr(), r.data_ptr(), z.data_ptr(), n.data_ptr(), ghn.data_ptr(),
           B, H, _stream()), nbytes=float(B) * H * (6 * 4 + 4 + 4 + 5 * 2))


def gru_gate_bwd(dh: torch.Tensor, dout_t: torch.Tensor, h_prev: Optional[torch.Tensor], r, z, n, ghn,
                 dgx_t: torch.Tensor, dgh: torch.Tensor) -> None:
    """In place on dh (B, H) fp32 (see include/cti_sm100.h); dgx_t (B, 3H) bf16 view with any row stride, dgh contiguous."""
    B, H = dh.shape
    _req(dh, F32, "gru_gate_bwd.dh")
    _req(dgh, BF16, "gru_gate_bwd.dgh")
    if dout_t.dtype != F32 or dout_t.stride(1) != 1 or dgx_t.dtype != BF16 or dgx_t.stride(1) != 1:
        raise RuntimeError("gru_gate_bwd: dout must be fp32, dgx bf16, both with unit column stride")
    _call("cti_gru_gate_bwd", _lib.load().cti_gru_gate_bwd,
          (dh.data_ptr(), dout_t.data_ptr(), dout_t.stride(0), _ptr(h_prev), 0 if h_prev is None else h_prev.stride(0),
           r.data_ptr(), z.data_ptr(), n.data_ptr(), ghn.data_ptr(), dgx_t.data_ptr(), dgx_t.stride(0), dgh.data_ptr(), B, H,
           _stream()), nbytes=float(B) * H * (4 * 3 + 4 * 2 + 6 * 2 + 4))


# --------------------------------------------------------------------------- #
# caller glue of the glimpse loop (fused call, glimpse.py)
# --------------------------------------------------------------------------- #
def _res_array(res):
    arr = (ctypes.c_void_p * 4)(*([t.data_ptr() for t in res] + [None] * (4 - len(res))))
    return arr


def _tok(t: torch.Tensor, name: str):
    if not t.is_cuda or t.dtype not in (F32, BF16) or t.dim() != 3 or not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous (rows, tokens, dim) fp32 / bf16 CUDA tensor")
    return t.data_ptr(), int(t.dtype == BF16)


def glimpse_residual_cast(q: torch.Tensor, res_q, a: Optional[torch.Tensor], res_a):
    """bf16((q + res_q[0][:, None] + ...)) as (B*Tq, D) and the same for a: the operands of the next glimpse's q_tucker /
    a_tucker GEMMs (reference src/MC/base_model.py:147-148 followed by the cast)."""
    B, Tq, D = q.shape
    for r in list(res_q) + list(res_a):
        _req(r, F32, "glimpse_residual_cast.res")
    qp, qb = _tok(q, "glimpse_residual_cast.q")
    oq = torch.empty((B * Tq, D), dtype=BF16, device=q.device)
    ap, ab, Ta, oa = None, 0, 0, None
    if a is not None:
        ap, ab = _tok(a, "glimpse_residual_cast.a")
        Ta = a.shape[1]
        oa = torch.empty((B * Ta, D), dtype=BF16, device=q.device)
    rq, ra = _res_array(res_q), _res_array(res_a)
    _call("cti_glimpse_glue", _lib.load().cti_glimpse_residual_cast,
          (qp, qb, rq, Tq, oq.data_ptr(), ap, ab, ra, Ta, _ptr(oa), len(res_q), B, D, _stream()),
          nbytes=float(B) * (Tq + Ta) * D * (q.element_size() + 2))
    return oq, oa


def glimpse_token_sum(q: torch.Tensor, res_q, a: Optional[torch.Tensor], res_a) -> torch.Tensor:
    """sum_t (q[:, t] + res_q...) + sum_t (a[:, t] + res_a...) -> (B, D) fp32 (reference src/MC/base_model.py:150)."""
    B, Tq, D = q.shape
    qp, qb = _tok(q, "glimpse_token_sum.q")
    ap, ab, Ta = None, 0, 0
    if a is not None:
        ap, ab = _tok(a, "glimpse_token_sum.a")
        Ta = a.shape[1]
    out = torch.empty((B, D), dtype=F32, device=q.device)
    rq, ra = _res_array(res_q), _res_array(res_a)
    _call("cti_glimpse_glue", _lib.load().cti_glimpse_token_sum,
          (qp, qb, rq, Tq, ap, ab, ra, Ta, len(res_q), out.data_ptr(), None, B, D, _stream()),
          nbytes=float(B) * (Tq + Ta) * D * q.element_size())
    return out


def glimpse_bcast_rows(x: torch.Tensor, Tq: int, Ta: int):
    """x (B, D) fp32 -> (B, Tq, D), (B, Ta, D) fp32 copies of it along the token axis."""
    _req(x, F32, "glimpse_bcast_rows.x")
    B, D = x.shape
    oq = torch.empty((B, Tq, D), dtype=F32, device=x.device)
    oa = torch.empty((B, Ta, D), dtype=F32, device=x.device) if Ta > 0 else None
    _call("cti_glimpse_glue", _lib.load().cti_glimpse_bcast_rows,
          (x.data_ptr(), oq.data_ptr(), Tq, _ptr(oa), Ta, B, D, _stream()), nbytes=4.0 * B * (Tq + Ta + 1) * D)
    return oq, oa


# --------------------------------------------------------------------------- #
# per-rank projections with per-rank input dropout, masks applied in registers (rank_proj.cu)
# --------------------------------------------------------------------------- #
def rank_proj_fused_ok(H: int, R: int) -> bool:
    """Shapes the fused kernels are built for (the reference's: h_mm = 512, rank = 32)."""
    return H == 512 and R in (16, 32)


def rank_proj_scale(p: float) -> float:
    """1 / (1 - p_eff): the drop rate is quantised to round(256 p) / 256 (include/cti_sm100.h)."""
    return float(_lib.load().cti_rank_proj_dropout_scale(float(p)))


def _rank_proj_call(fn_name: str, tag: str, probs, R: int) -> None:
    """probs: list of dicts with the cti_rank_proj_problem fields as tensors (+ 'drop')."""
    arr = (_lib.RankProjProblem * len(probs))()
    flops = 0.0
    H = probs[0]["y"].shape[1]
    for q, d in zip(arr, probs):
        for k in ("y", "w_eff", "bias", "out", "dz", "dzt", "dw_accum"):
            setattr(q, k, _ptr(d.get(k)))
        q.M = d["y"].shape[0]
        q.p, q.seed, q.site = d["drop"]
        flops += 2.0 * q.M * H * R * 16
    lib = _lib.load()
    kinds = len({d["drop"][0] == 0.5 for d in probs})          # one launch per mask kind
    _call("cti_rank_proj_dropout", getattr(lib, fn_name), (arr, len(probs), H, R, _stream()), kernels=kinds, flops=flops,
          tag=tag + " rows " + "+".join(str(d["y"].shape[0]) for d in probs))


def rank_proj_dropout_fwd(ys, ws, biases, R: int, drops):
    """Batched over modalities: lists of y (M_i, H) bf16, w_eff, bias, drop -> list of out (M_i, R*16) bf16."""
    probs = []
    for y, w, b, drop in zip(ys, ws, biases, drops):
        _req(y, BF16, "rank_proj_dropout_fwd.y")
        _req(w, BF16, "rank_proj_dropout_fwd.w")
        _req(b, F32, "rank_proj_dropout_fwd.bias")
        probs.append({"y": y, "w_eff": w, "bias": b, "out": torch.empty((y.shape[0], R * 16), dtype=BF16, device=y.device),
                      "drop": drop})
    _rank_proj_call("cti_rank_proj_dropout_fwd", "fwd", probs, R)
    return [d["out"] for d in probs]


def rank_proj_dropout_dgrad(dzs, ws, ys, R: int, drops):
    """-> list of pre-activation gradients of the layers that produced the y (bf16, y's ReLU mask applied)."""
    probs = []
    for dz, w, y, drop in zip(dzs, ws, ys, drops):
        _req(dz, BF16, "rank_proj_dropout_dgrad.dz")
        _req(y, BF16, "rank_proj_dropout_dgrad.y")
        probs.append({"y": y, "w_eff": w, "dz": dz, "dzt": torch.empty(y.shape, dtype=BF16, device=y.device), "drop": drop})
    _rank_proj_call("cti_rank_proj_dropout_dgrad", "dgrad", probs, R)
    return [d["dzt"] for d in probs]


def rank_proj_dropout_wgrad(dzs, ys, dws, R: int, drops) -> None:
    """dws[i] (R*16, H) fp32 += masked wgrad of problem i."""
    probs = []
    for dz, y, dw, drop in zip(dzs, ys, dws, drops):
        _req(dz, BF16, "rank_proj_dropout_wgrad.dz")
        _req(y, BF16, "rank_proj_dropout_wgrad.y")
        _req(dw, F32, "rank_proj_dropout_wgrad.dw")
        probs.append({"y": y, "dz": dz, "dw_accum": dw, "drop": drop})
    _rank_proj_call("cti_rank_proj_dropout_wgrad", "wgrad", probs, R)


def rank_proj_dropout_mask(M: int, H: int, R: int, drop, device) -> torch.Tensor:
    """keep[r, m, k] (uint8) exactly as the fused kernels regenerate it (tests)."""
    keep = torch.empty((R, M, H), dtype=torch.uint8, device=device)
    _call("cti_rank_proj_dropout", _lib.load().cti_rank_proj_dropout_mask,
          (keep.data_ptr(), M, H, R, drop[0], drop[1], drop[2], _stream()))
    return keep
