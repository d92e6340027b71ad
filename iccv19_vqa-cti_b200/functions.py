"""torch.autograd.Function wrappers: each one is a whole reference module call (forward and the
hand-written backward) expressed as a sequence of C-ABI kernel launches.  No torch math op runs
on activations here; torch supplies memory, streams and the autograd tape.

Backward formulas: SURVEY.md appendix B (each verified against the reference's autograd there).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch.autograd import Function

from . import kernels as K_

BF16 = torch.bfloat16
F32 = torch.float32


# --------------------------------------------------------------------------- #
# one weight-normed linear layer group (reference src/fc.py:22-29): helpers, not autograd
# --------------------------------------------------------------------------- #
class Packed:
    """bf16 effective weight W_eff = V g/||V|| of one layer group and the squared norms.
    ``dw`` is set by ``prepack`` when the layer's weight-norm backward is DEFERRED: the layer's backward then writes
    dW_eff into that persistent buffer and hands it to the proxy of ``weight_v`` instead of computing dV / dg itself --
    one multi-tensor launch (``cti_wn_grad_multi``) finishes every layer of the model at the end of the backward pass."""
    __slots__ = ("w", "sumsq", "dw")

    def __init__(self, w: torch.Tensor, sumsq: torch.Tensor, dw: Optional[torch.Tensor] = None):
        self.w = w
        self.sumsq = sumsq
        self.dw = dw


def pack_layer(V: torch.Tensor, g: torch.Tensor, n_groups: int) -> Packed:
    """Single layers are padded with zero rows to a multiple of 8 outputs (classifier heads: 2 or 3129 classes) and
    with zero columns to a multiple of 8 inputs (``c_prj = FCNet([11, num_hid])``, reference src/MC/base_model.py:176):
    a TMA operand's row pitch must be a multiple of 16 bytes.  Zero columns leave ||V||_F unchanged."""
    Vd = V.detach()
    if n_groups == 1 and Vd.shape[1] % 8:
        Vd = _pad_cols(Vd, -(-Vd.shape[1] // 8) * 8)
    w, sumsq = K_.wn_pack(Vd.contiguous(), g.detach().reshape(n_groups).contiguous(), n_groups,
                          pad_rows_to=8 if n_groups == 1 else 1)
    return Packed(w, sumsq)


def _pad_cols(x: torch.Tensor, n: int) -> torch.Tensor:
    """x (..., c) -> (..., n) with zero columns appended (no-op when c == n)."""
    if x.shape[-1] == n:
        return x
    out = torch.zeros((*x.shape[:-1], n), dtype=x.dtype, device=x.device)
    out[..., :x.shape[-1]] = x
    return out


def lin_fwd(x: torch.Tensor, pk: Packed, bias: torch.Tensor, relu: bool, out_bf16: bool = True,
            out_f32: bool = False):
    """y = act(x W_eff^T + b); x (M, K_in) bf16 -> (bf16 | None, fp32 | None)."""
    M, Kin = x.shape
    N = pk.w.shape[0]                                    # padded width of the pack
    return K_.gemm(x, pk.w, M, N, Kin, bias=_pad_cols(bias.detach(), N), relu=relu, out_bf16=out_bf16, out_f32=out_f32)


def _pick_splits(tiles: int, k_blocks: int) -> int:
    """Split-K factor for the wgrad GEMM: the smallest split whose CTA count fills >= 95 % of its waves (the first
    version stopped at 85 %: 64 tiles x 2 splits = 128 CTAs left 20 of 148 SMs idle for the whole launch), keeping at
    least 4 K blocks per split; otherwise the best fill seen."""
    best, best_eff = 1, 0.0
    for s in range(1, 33):
        if s > 1 and s * 4 > k_blocks:
            break
        units = tiles * s
        eff = units / (-(-units // K_.NUM_SMS) * K_.NUM_SMS)
        if eff > best_eff + 1e-9:
            best, best_eff = s, eff
        if eff >= 0.95:
            return s
    return best


# --------------------------------------------------------------------------- #
# training-mode dropout: a site is (p, seed, offset); the kernels regenerate the Philox mask from it
# --------------------------------------------------------------------------- #
_DROP_SITES = [0]
# Input dropout of the R per-rank nets of a TCNet: "independent" = one mask per rank, the reference's semantics
# (src/tc.py:29-31); "shared" = one mask per modality shared by its R nets (same marginals, ~1.6x faster training step).
RANK_DROPOUT = "independent"


def new_drop(p: float, training: bool):
    """A fresh dropout site, or None when dropout is the identity (eval mode / p == 0).  Masks are a pure function
    of (torch.initial_seed(), site counter, element index): deterministic under torch.manual_seed and call order."""
    if not training or p <= 0:
        return None
    _DROP_SITES[0] += 1
    return (float(p), torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, _DROP_SITES[0])


def drop_features(v2d: torch.Tensor, drop) -> torch.Tensor:
    """bf16(dropout(v)) of the image features: fp32 input goes through the fused cast + dropout kernel, bf16 input (the
    loader wire format) through the bf16 dropout kernel."""
    if v2d.dtype == BF16:
        return K_.dropout_bf16(v2d, drop)
    return K_.cast_rows_dropout(v2d, drop)[0]


def cast_in(x2d: torch.Tensor, drop):
    x2d = x2d.detach().contiguous()
    if x2d.dtype != F32:
        x2d = x2d.float()
    return K_.cast_rows(x2d)[0] if drop is None else K_.cast_rows_dropout(x2d, drop)[0]


def zero_slab(device, shapes: Sequence[Tuple[int, ...]]) -> List[torch.Tensor]:
    """fp32 zero tensors of the given shapes carved out of ONE allocation / ONE fill kernel (the split-K accumulators
    and bias-gradient sums of a whole backward call; each is 16-byte aligned for the TMA reduce-add)."""
    sizes = [-(-int(torch.Size(sh).numel()) // 4) * 4 for sh in shapes]
    slab = torch.zeros((sum(sizes),), dtype=F32, device=device)
    out, o = [], 0
    for sh, n in zip(shapes, sizes):
        out.append(slab[o:o + int(torch.Size(sh).numel())].view(sh))
        o += n
    return out


_TOK_ATTR = "_cti_b200_tok"


def cast_tokens(t: torch.Tensor, drop) -> torch.Tensor:
    """bf16 (rows * tokens, dim) copy of a (rows, tokens, dim) question / answer tensor.  Without dropout the copy is
    cached on the tensor object (keyed by its version counter): the attention and the first glimpse's pooling receive
    the very same q and a (reference src/MC/base_model.py:143-146), so the second cast is free."""
    x2d = t.reshape(t.shape[0] * t.shape[1], -1)
    if drop is not None:
        return cast_in(x2d, drop)
    if t.dtype == BF16:                  # tokens that already travel as bf16 (loader wire format): no cast pass
        return x2d.detach().contiguous()
    key = (t._version, t.data_ptr(), tuple(t.shape))
    hit = getattr(t, _TOK_ATTR, None)
    if hit is not None and hit[0] == key:
        return hit[1]
    xb = cast_in(x2d, None)
    try:
        setattr(t, _TOK_ATTR, (key, xb))
    except Exception:
        pass
    return xb


def _f32_dx(dtype, drop) -> bool:
    """fp32 dgrad output unless the token tensor is bf16 (then its gradient is bf16 too and leaves the GEMM epilogue as
    such); with input dropout the fp32 path is kept (the mask kernel works on fp32)."""
    return dtype != BF16 or drop is not None


def _finish_dx(dx, drop, dtype, shape):
    if dx is None:
        return None
    if drop is not None:
        K_.dropout_f32_(dx, drop)
    if dx.dtype != dtype:
        dx = dx.to(dtype)
    return dx.view(shape)


def lin_bwd(x: torch.Tensor, dz: torch.Tensor, V: torch.Tensor, g: torch.Tensor, pk: Packed, n_groups: int,
            need_dx: bool, dx_relu_aux: Optional[torch.Tensor] = None, dx_f32: bool = False, alpha: float = 1.0,
            dx_alpha: float = 1.0, dw: Optional[torch.Tensor] = None):
    """Backward of one layer group given the pre-activation gradient dz (M, N) bf16.
    Returns dV (like V), dg (like g), dx (bf16 masked by dx_relu_aux > 0, or fp32) or None.
    dw: optional zeroed (N, K_in) fp32 accumulator (see zero_slab)."""
    M, Kin = x.shape
    N = dz.shape[1]
    # wgrad: dW_eff[N, K_in] = dz^T x  -- both operands MN-major, split over the row (reduction) dimension
    tile_n = 256 if Kin >= 256 else 128
    tiles = -(-N // 128) * -(-Kin // tile_n)
    splits = _pick_splits(tiles, -(-M // 64))
    deferred = pk.dw is not None
    if deferred:
        dw = pk.dw                                   # persistent, zeroed by prepack at the top of the step
    elif dw is None:
        dw = torch.zeros((N, Kin), dtype=F32, device=x.device)
    K_.gemm(dz, x, N, Kin, M, a_mn=True, b_mn=True, accum_f32=dw, k_splits=splits, tile_n=tile_n, alpha=alpha)
    if deferred:                                     # dW_eff goes to the proxy of weight_v; dV / dg come from prepack's backward
        dV, dg = dw[:V.shape[0]], None
    else:
        dV, dg = K_.wn_grad(dw[:V.shape[0]], V.detach().contiguous(), g.detach().reshape(n_groups).contiguous(), pk.sumsq,
                            n_groups)
        dg = dg.reshape(g.shape)
    dx = None
    if need_dx:
        # dgrad: dx[M, K_in] = dz W_eff  (W_eff stored [N][K_in] = MN-major B operand)
        ob, of = K_.gemm(dz, pk.w, M, Kin, N, b_mn=True, relu_aux=dx_relu_aux, out_bf16=not dx_f32, out_f32=dx_f32,
                         alpha=dx_alpha)
        dx = of if dx_f32 else ob
    return dV, dg, dx


# --------------------------------------------------------------------------- #
# per-rank nets with INDEPENDENT input dropout masks (reference src/tc.py:29-31: each of the R FCNets owns a Dropout)
# --------------------------------------------------------------------------- #
RANK_GROUP = 4          # ranks handled per GEMM: the masked copies of the input are materialised 4 ranks at a time


def _block_diag_weights(w_eff: torch.Tensor, R: int, rg: int) -> torch.Tensor:
    """(R*d, H) -> (R/rg, rg*d, rg*H): group g's weight is block diagonal, block j = W_eff of rank g*rg + j."""
    d, H = w_eff.shape[0] // R, w_eff.shape[1]
    src = w_eff.view(R // rg, rg, d, H)
    out = torch.zeros((R // rg, rg * d, rg * H), dtype=w_eff.dtype, device=w_eff.device)
    for j in range(rg):
        out[:, j * d:(j + 1) * d, j * H:(j + 1) * H] = src[:, j]
    return out


def rank_proj_fwd(y: torch.Tensor, pk: Packed, bias: torch.Tensor, drop, R: int) -> torch.Tensor:
    """Vc[:, r*d:(r+1)*d] = relu(W_r (keep_r * y) / (1-p) + b_r) with one mask per rank: the masked copies of y are
    materialised RANK_GROUP ranks at a time and multiplied with a block-diagonal weight by the ordinary GEMM."""
    M, H = y.shape
    N = pk.w.shape[0]
    d = N // R
    if d == 16 and K_.rank_proj_fused_ok(H, R):       # masks applied to the A fragments in registers (rank_proj.cu)
        return K_.rank_proj_dropout_fwd([y], [pk.w], [bias.detach().contiguous()], R, [drop])[0]
    rg = RANK_GROUP if R % RANK_GROUP == 0 else 1
    out = torch.empty((M, N), dtype=BF16, device=y.device)
    b = bias.detach()
    with K_.gemm_unbatched():                        # torch ops below read what earlier GEMMs wrote
        wt = _block_diag_weights(pk.w, R, rg)
        for gi in range(R // rg):
            xt = K_.dropout_expand(y, rg, gi * rg, drop)
            K_.gemm(xt, wt[gi], M, rg * d, rg * H, bias=b[gi * rg * d:(gi + 1) * rg * d], relu=True,
                    out=out[:, gi * rg * d:(gi + 1) * rg * d])
    return out


def rank_proj_fused(ys, pks, drops, R: int) -> bool:
    """All modalities of the call go through the fused kernels (rank_proj.cu) in one batched launch per pass."""
    return (all(dr is not None for dr in drops) and all(pk.w.shape[0] == R * 16 for pk in pks)
            and all(K_.rank_proj_fused_ok(y.shape[1], R) for y in ys))


def rank_proj_bwd_fused(ys, dzs, Vs, gs, pks, drops, R: int):
    """Backward of the fused per-rank projections for a list of modalities: one wgrad and one dgrad call for all of them.
    -> [(dV, dg, dzt)]: dzt is the bf16 pre-activation gradient of the layer that produced y (y's ReLU mask applied)."""
    dws = [pk.dw if pk.dw is not None else torch.zeros((R * 16, y.shape[1]), dtype=F32, device=y.device)
           for pk, y in zip(pks, ys)]
    K_.rank_proj_dropout_wgrad(dzs, ys, dws, R, drops)
    dzts = K_.rank_proj_dropout_dgrad(dzs, [pk.w for pk in pks], ys, R, drops)
    out = []
    for V, g, pk, dw, dzt in zip(Vs, gs, pks, dws, dzts):
        if pk.dw is not None:                        # deferred weight-norm backward (see Packed)
            out.append((dw, None, dzt))
        else:
            dV, dg = K_.wn_grad(dw, V.detach().contiguous(), g.detach().reshape(R).contiguous(), pk.sumsq, R)
            out.append((dV, dg.reshape(g.shape), dzt))
    return out


def rank_proj_bwd(y: torch.Tensor, dz: torch.Tensor, V: torch.Tensor, g: torch.Tensor, pk: Packed, drop, R: int):
    """Backward of rank_proj_fwd.  Returns dV, dg and the fp32 gradient w.r.t. y (before y's own ReLU mask)."""
    M, H = y.shape
    N = pk.w.shape[0]
    d = N // R
    if d == 16 and K_.rank_proj_fused_ok(H, R):
        return rank_proj_bwd_fused([y], [dz], [V], [g], [pk], [drop], R)[0]
    rg = RANK_GROUP if R % RANK_GROUP == 0 else 1
    wt = _block_diag_weights(pk.w, R, rg)
    dw = torch.empty((N, H), dtype=F32, device=y.device)
    acc = torch.zeros((M, H), dtype=F32, device=y.device)
    splits = _pick_splits(-(-rg * d // 128) * -(-rg * H // 256), -(-M // 64))
    with K_.gemm_unbatched():                        # torch ops below read what the GEMMs wrote
        for gi in range(R // rg):
            xt = K_.dropout_expand(y, rg, gi * rg, drop)                     # same masks as the forward
            dzg = dz[:, gi * rg * d:(gi + 1) * rg * d]
            dwt = torch.zeros((rg * d, rg * H), dtype=F32, device=y.device)
            K_.gemm(dzg, xt, rg * d, rg * H, M, a_mn=True, b_mn=True, accum_f32=dwt, k_splits=splits, tile_n=256)
            for j in range(rg):
                dw[(gi * rg + j) * d:(gi * rg + j + 1) * d] = dwt[j * d:(j + 1) * d, j * H:(j + 1) * H]
            dxt, _ = K_.gemm(dzg, wt[gi], M, rg * H, rg * d, b_mn=True)
            K_.dropout_reduce_(dxt, acc, rg, gi * rg, drop)
    if pk.dw is not None:                            # deferred weight-norm backward (see Packed)
        pk.dw.copy_(dw)
        return pk.dw, None, acc
    dV, dg = K_.wn_grad(dw, V.detach().contiguous(), g.detach().reshape(R).contiguous(), pk.sumsq, R)
    return dV, dg.reshape(g.shape), acc


def _colsum(dz: torch.Tensor, n: int, db: Optional[torch.Tensor] = None) -> torch.Tensor:
    if db is None:
        db = torch.zeros((n,), dtype=F32, device=dz.device)
    K_.act_bwd_bias(dz, None, False, db)
    return db


# --------------------------------------------------------------------------- #
# T_g <-> packed core  (reference src/Tensor.py:6-8 glimpse-axis re-view)
# --------------------------------------------------------------------------- #
_TPACK_IDX = {}


def tpack_index(R: int, d: int, G: int, device) -> torch.Tensor:
    """int64 index p such that tpack.flat = T_g.flat[p], tpack[r][l][(i,g,j)] = T_eff[r,i,j,l,g].
    T_eff follows from pushing indices through the reference's own view chain: mode 1 flattens the
    core in (l, j, g) order but re-reads that axis as (g', l', j')."""
    key = (R, d, G, str(device))
    if key not in _TPACK_IDX:
        idx = torch.arange(R * d * d * d * G).view(R, d, d, d, G)            # (r,i,j,l,g) -> flat T_g offset
        reread = idx.permute(0, 1, 3, 2, 4).reshape(R, d, G, d, d)           # (r,i,g',l',j')
        teff = reread.permute(0, 1, 4, 3, 2)                                 # (r,i,j',l',g')
        _TPACK_IDX[key] = teff.permute(0, 3, 1, 4, 2).reshape(-1).to(device)  # (r,l,i,g,j)
    return _TPACK_IDX[key]


def pack_core(T_g: torch.Tensor) -> torch.Tensor:
    _, R, d, _, _, G, ho = T_g.shape
    idx = tpack_index(R, d, G, T_g.device)
    return T_g.detach().reshape(-1)[idx].to(BF16).view(R, d, d * G * d)


def unpack_core_grad(dtpack: torch.Tensor, T_g: torch.Tensor) -> torch.Tensor:
    _, R, d, _, _, G, ho = T_g.shape
    idx = tpack_index(R, d, G, T_g.device)
    out = torch.empty(T_g.numel(), dtype=F32, device=T_g.device)
    out[idx] = dtpack.reshape(-1)
    return out.view_as(T_g)


# --------------------------------------------------------------------------- #
class WNLinearFn(Function):
    """y = act(x W^T + b), W = V g/||V||_F: one FCNet layer (reference src/fc.py:27-29,33-34).
    x fp32 (M, K_in) -> fp32 (M, N)."""

    @staticmethod
    def forward(ctx, x, V, g, bias, relu: bool, pk: Optional[Packed], drop=None):
        if pk is None:
            pk = pack_layer(V, g, 1)
        Kp = pk.w.shape[1]                                  # input width padded to a multiple of 8 (zero columns)
        ctx.kin = x.shape[1]
        xb = cast_in(x if Kp == x.shape[1] else _pad_cols(x.detach(), Kp), drop)
        ctx.drop = drop
        yb, yf = lin_fwd(xb, pk, bias, relu, out_bf16=relu, out_f32=True)
        ctx.save_for_backward(xb, yb if relu else None, V, g)
        ctx.pk = pk
        ctx.relu = relu
        ctx.need_dx = x.requires_grad
        return yf if yf.shape[1] == V.shape[0] else yf[:, :V.shape[0]]

    @staticmethod
    def backward(ctx, dy):
        xb, yb, V, g = ctx.saved_tensors
        N, Np = V.shape[0], ctx.pk.w.shape[0]
        db = torch.zeros((Np,), dtype=F32, device=dy.device)
        dz = K_.act_bwd_bias(_pad_cols(dy, Np).contiguous(), yb if ctx.relu else None, True, db)
        Kin, Kp = ctx.kin, ctx.pk.w.shape[1]
        Vp = V if Kp == Kin else _pad_cols(V.detach(), Kp)
        dV, dg, dx = lin_bwd(xb, dz, Vp, g, ctx.pk, 1, ctx.need_dx, dx_f32=True)
        if dx is not None and ctx.drop is not None:
            K_.dropout_f32_(dx, ctx.drop)
        if Kp != Kin:
            dV = dV[:, :Kin].contiguous()
            dx = None if dx is None else dx[:, :Kin].contiguous()
        return dx, dV, dg, db[:N], None, None, None


# --------------------------------------------------------------------------- #
class TriLogitsFn(Function):
    """TCNet.forward (reference src/tc.py:41-52): tucker projections, R per-rank projections (one
    grouped GEMM per modality), rank-R trilinear contraction; optional zero-row mask to -inf
    (src/attention.py:55-56).  Returns the (B,K,Q,A,G) view of a (B,G,K,Q,A) buffer."""

    @staticmethod
    def forward(ctx, dims, packs, drops, v_bf16, rowmask, q, a, T_g, *w):
        """drops: None (eval) or (v_f32_2d, dv, dq, da, dvn, dqn, dan) -- input dropout of the three tucker nets and of
        the per-rank nets (RANK_DROPOUT: one mask per rank as in the reference, or one per modality)."""
        B, K, Q, A, G, R = dims
        vr = (B * K) // v_bf16.shape[0]               # rows sharing one image (v has B / vr samples; see tc.py)
        ctx.vr = vr
        dv = dq = da = dvn = dqn = dan = None
        if drops is not None:
            v_f32, dv, dq, da, dvn, dqn, dan = drops
            if dv is not None:
                v_bf16 = K_.dropout_bf16(v_bf16, dv)      # the cached bf16 cast is the source: half the bytes of a cast + dropout pass over fp32 v
        # w = (V, g, b) x [v_tucker, q_tucker, a_tucker, v_net, q_net, a_net]
        groups = (1, 1, 1, R, R, R)
        pk: List[Packed] = [packs[i] if packs is not None else pack_layer(w[3 * i], w[3 * i + 1], groups[i])
                            for i in range(6)]
        xq = cast_tokens(q, dq)
        xa = cast_tokens(a, da)
        with K_.gemm_batch():                 # the question- and answer-side projections share one launch
            yv, _ = lin_fwd(v_bf16, pk[0], w[2], True)
            yq, _ = lin_fwd(xq, pk[1], w[5], True)
            ya, _ = lin_fwd(xa, pk[2], w[8], True)
        independent = RANK_DROPOUT == "independent"
        ctx.independent = independent
        ctx.drops = (dq, da, dvn, dqn, dan)

        def rank_nets(y, pki, bias, drop):
            if drop is None:
                return y, lin_fwd(y, pki, bias, True)[0]
            if independent:                                   # one mask per rank, as in the reference
                return y, rank_proj_fwd(y, pki, bias, drop, R)
            # shared mask: the dropped copy doubles as the ReLU-and-dropout mask of the dgrad epilogue
            yd = K_.dropout_bf16(y, drop)
            return yd, lin_fwd(yd, pki, bias, True)[0]
        ctx.rank_fused = independent and rank_proj_fused((yv, yq, ya), pk[3:6], (dvn, dqn, dan), R)
        if ctx.rank_fused:                    # per-rank masks applied in registers; the three modalities share the launches
            vc, qc, ac = K_.rank_proj_dropout_fwd([yv, yq, ya], [pk[3].w, pk[4].w, pk[5].w],
                                                  [w[11].detach().contiguous(), w[14].detach().contiguous(),
                                                   w[17].detach().contiguous()], R, [dvn, dqn, dan])
        else:
            with K_.gemm_batch():
                yv, vc = rank_nets(yv, pk[3], w[11], dvn)
                yq, qc = rank_nets(yq, pk[4], w[14], dqn)
                ya, ac = rank_nets(ya, pk[5], w[17], dan)
        tpack = pack_core(T_g)
        # training: the kernel also leaves its bf16 N1 intermediate in HBM, the backward picks it up instead of recomputing it
        n1 = None
        if any(ctx.needs_input_grad):
            logits, n1 = K_.trilinear_fwd(vc, qc, ac, tpack, rowmask, B, K, Q, A, G, R, vr, save_n1=True)
        else:
            logits = K_.trilinear_fwd(vc, qc, ac, tpack, rowmask, B, K, Q, A, G, R, vr)
        ctx.n1 = n1
        ctx.save_for_backward(v_bf16, xq, xa, yv, yq, ya, vc, qc, ac, tpack, T_g, *w)
        ctx.pk = pk
        ctx.dims = dims
        ctx.need = (q.requires_grad, a.requires_grad)
        ctx.tok_dtypes = (q.dtype, a.dtype)
        return logits.permute(0, 2, 3, 4, 1)

    @staticmethod
    def backward(ctx, dlogits):
        B, K, Q, A, G, R = ctx.dims
        v_bf16, xq, xa, yv, yq, ya, vc, qc, ac, tpack, T_g = ctx.saved_tensors[:11]
        w = ctx.saved_tensors[11:]
        pk = ctx.pk
        dl = dlogits.permute(0, 4, 1, 2, 3).contiguous()
        dzv, dzq, dza, dbvn, dbqn, dban, dtpack = K_.trilinear_bwd(vc, qc, ac, tpack, dl, B, K, Q, A, G, R, ctx.vr, n1=ctx.n1)
        ctx.n1 = None
        # per-rank nets: input = tucker output (post-ReLU), so dx is masked by it -> dz of the tucker layer
        dq_drop, da_drop, dvn, dqn, dan = ctx.drops
        sc = lambda d: 1.0 if d is None else 1.0 / (1.0 - d[0])
        H = yv.shape[1]

        # every split-K accumulator and bias-gradient sum of this call: one allocation, one fill kernel
        RD = R * 16
        # (deferred layers accumulate dW_eff in prepack's persistent buffers: only the bias sums need zeroes then)
        dw_shapes = [(RD, H)] * 3 + [(H, v_bf16.shape[1]), (H, xq.shape[1]), (H, xa.shape[1])]
        dw_pk = [pk[3], pk[4], pk[5], pk[0], pk[1], pk[2]]
        need = [i for i in range(6) if dw_pk[i].dw is None]
        slab = zero_slab(yv.device, [(H,)] * 3 + [dw_shapes[i] for i in need])
        zs = [None] * 9
        zs[3:6] = slab[:3]
        for j, i in enumerate(need):
            zs[i if i < 3 else i + 3] = slab[3 + j]

        def rank_nets_bwd(y, dz, V, g, pki, drop, dw_):
            """-> dV, dg, pre-activation gradient of the tucker layer (bf16); None where its ReLU mask / bias sum is
            still to be applied (independent per-rank dropout: the fp32 accumulator comes back instead)"""
            if drop is not None and ctx.independent:
                dV_, dg_, acc = rank_proj_bwd(y, dz, V, g, pki, drop, R)
                if acc.dtype == BF16:                          # fused kernels: ReLU mask of y already applied
                    return dV_, dg_, acc, None
                return dV_, dg_, None, acc
            dV_, dg_, dzt = lin_bwd(y, dz, V, g, pki, R, True, dx_relu_aux=y, dx_alpha=sc(drop), dw=dw_)
            return dV_, dg_, dzt, None

        def finish(y, dzt, acc, db_):
            if dzt is None:
                return K_.act_bwd_bias(acc, y, True, db_), db_
            return dzt, _colsum(dzt, H, db_)
        # the GEMMs of the three modalities are independent: the question- and answer-side ones pair up (gemm_batch)
        if ctx.rank_fused:
            (dVvn, dgvn, dzvt), (dVqn, dgqn, dzqt), (dVan, dgan, dzat) = rank_proj_bwd_fused(
                [yv, yq, ya], [dzv, dzq, dza], [w[9], w[12], w[15]], [w[10], w[13], w[16]], pk[3:6], [dvn, dqn, dan], R)
            accv = accq = acca = None
        else:
            with K_.gemm_batch():
                dVvn, dgvn, dzvt, accv = rank_nets_bwd(yv, dzv, w[9], w[10], pk[3], dvn, zs[0])
                dVqn, dgqn, dzqt, accq = rank_nets_bwd(yq, dzq, w[12], w[13], pk[4], dqn, zs[1])
                dVan, dgan, dzat, acca = rank_nets_bwd(ya, dza, w[15], w[16], pk[5], dan, zs[2])
        dzvt, dbvt = finish(yv, dzvt, accv, zs[3])
        dzqt, dbqt = finish(yq, dzqt, accq, zs[4])
        dzat, dbat = finish(ya, dzat, acca, zs[5])
        with K_.gemm_batch():
            dVvt, dgvt, _ = lin_bwd(v_bf16, dzvt, w[0], w[1], pk[0], 1, False, dw=zs[6])
            # the gradient of a token tensor has the tensor's dtype: fp32, or bf16 straight from the dgrad epilogue
            dVqt, dgqt, dq = lin_bwd(xq, dzqt, w[3], w[4], pk[1], 1, ctx.need[0], dw=zs[7],
                                     dx_f32=_f32_dx(ctx.tok_dtypes[0], dq_drop))
            dVat, dgat, da = lin_bwd(xa, dzat, w[6], w[7], pk[2], 1, ctx.need[1], dw=zs[8],
                                     dx_f32=_f32_dx(ctx.tok_dtypes[1], da_drop))
        dT = unpack_core_grad(dtpack, T_g)
        dq = _finish_dx(dq, dq_drop, ctx.tok_dtypes[0], (B, Q, -1))
        da = _finish_dx(da, da_drop, ctx.tok_dtypes[1], (B, A, -1))
        return (None, None, None, None, None, dq, da, dT,
                dVvt, dgvt, dbvt, dVqt, dgqt, dbqt, dVat, dgat, dbat,
                dVvn, dgvn, dbvn.view_as(w[11]), dVqn, dgqn, dbqn.view_as(w[14]), dVan, dgan, dban.view_as(w[17]))


# --------------------------------------------------------------------------- #
class MaskedSoftmaxFn(Function):
    """softmax over the flattened attention domain per (b, g) (reference src/attention.py:58 / :39).
    x: logical (B, *dom, G) view of a (B, G, L) buffer when g_last, else (B, G, *dom) contiguous."""

    @staticmethod
    def forward(ctx, x, g_last: bool):
        if g_last:
            nd = x.dim()
            xi = x.permute(0, nd - 1, *range(1, nd - 1)).contiguous()        # (B, G, *dom): no copy in the native layout
        else:
            xi = x.contiguous()
        B, G = xi.shape[0], xi.shape[1]
        L = xi[0, 0].numel()
        p = K_.softmax_fwd(xi, B * G, L)
        ctx.save_for_backward(p)
        ctx.g_last = g_last
        if g_last:
            nd = p.dim()
            return p.permute(0, *range(2, nd), 1)
        return p

    @staticmethod
    def backward(ctx, dp):
        (p,) = ctx.saved_tensors
        B, G = p.shape[0], p.shape[1]
        L = p[0, 0].numel()
        if ctx.g_last:
            nd = dp.dim()
            dpi = dp.permute(0, nd - 1, *range(1, nd - 1))                   # logical (B, G, *dom), any strides
        else:
            dpi = dp
        # the kernel addresses dp[b,g,e] with three strides: the domain must flatten with one stride
        dom_sizes, dom_strides = dpi.shape[2:], dpi.stride()[2:]
        se = dom_strides[-1]
        ok = dpi.dtype == F32
        run = se
        for sz, st in zip(reversed(dom_sizes), reversed(dom_strides)):
            if sz != 1 and st != run:
                ok = False
            run *= sz
        if not ok:
            dpi = dpi.to(F32).contiguous()
            se = 1
        dl = K_.softmax_bwd(p, dpi, dpi.stride(0), dpi.stride(1), se, B, G, L)
        if ctx.g_last:
            nd = dl.dim()
            return dl.permute(0, *range(2, nd), 1), None
        return dl, None


# --------------------------------------------------------------------------- #
def _sample_contiguous(w: torch.Tensor) -> torch.Tensor:
    """Attention weights must be contiguous within a sample (any batch stride)."""
    run = 1
    for sz, st in zip(reversed(w.shape[1:]), reversed(w.stride()[1:])):
        if sz != 1 and st != run:
            return w.contiguous()
        run *= sz
    return w


class PoolFn(Function):
    """TCNet.forward_with_weights (reference src/tc.py:54-61) and, with a = None,
    BCNet.forward_with_weights (src/bc.py:70-74): projections + attention-weighted pooling."""

    @staticmethod
    def forward(ctx, dims, packs, drops, v_bf16, q, a, wts, *w):
        """drops: None (eval) or (v_f32_2d, dv, dq, da): input dropout of the three projections."""
        B, K, Q, A, C = dims
        vr = (B * K) // v_bf16.shape[0]               # rows sharing one image
        ctx.vr = vr
        n = 3 if A > 0 else 2
        pk = [packs[i] if packs is not None else pack_layer(w[3 * i], w[3 * i + 1], 1) for i in range(n)]
        dq_drop = da_drop = None
        if drops is not None:
            v_f32, dv, dq_drop, da_drop = drops
            if dv is not None:
                v_bf16 = K_.dropout_bf16(v_bf16, dv)      # the cached bf16 cast is the source: half the bytes of a cast + dropout pass over fp32 v
        ctx.drops = (dq_drop, da_drop)
        xq = cast_tokens(q, dq_drop)
        xa = ap = None
        if A > 0:
            xa = cast_tokens(a, da_drop)
        with K_.gemm_batch():                 # the question- and answer-side projections share one launch
            vp, _ = lin_fwd(v_bf16, pk[0], w[2], True)
            qp, _ = lin_fwd(xq, pk[1], w[5], True)
            if A > 0:
                ap, _ = lin_fwd(xa, pk[2], w[8], True)
        wd = _sample_contiguous(wts.detach())
        if wd.dtype != F32:
            wd = wd.float()
        out = K_.tri_pool_fwd(vp, qp, ap, wd, wd.stride(0), B, K, Q, A, C, vr)
        ctx.save_for_backward(v_bf16, xq, xa, vp, qp, ap, wd, *w)
        ctx.pk = pk
        ctx.dims = dims
        ctx.need = (q.requires_grad, a is not None and a.requires_grad, wts.requires_grad)
        ctx.tok_dtypes = (q.dtype, a.dtype if a is not None else F32)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, K, Q, A, C = ctx.dims
        v_bf16, xq, xa, vp, qp, ap, wd = ctx.saved_tensors[:7]
        w = ctx.saved_tensors[7:]
        pk = ctx.pk
        dzv, dzq, dza, dbv, dbq, dba, dw = K_.tri_pool_bwd(vp, qp, ap, wd, wd.stride(0), dout.contiguous(), B, K, Q, A,
                                                          C, ctx.vr)
        shapes = [(C, v_bf16.shape[1]), (C, xq.shape[1])] + ([(C, xa.shape[1])] if A > 0 else [])
        need = [i for i in range(len(shapes)) if pk[i].dw is None]
        slab = zero_slab(vp.device, [shapes[i] for i in need]) if need else []
        zs = [None] * 3
        for j, i in enumerate(need):
            zs[i] = slab[j]
        da = None
        with K_.gemm_batch():
            dVv, dgv, _ = lin_bwd(v_bf16, dzv, w[0], w[1], pk[0], 1, False, dw=zs[0])
            dVq, dgq, dq = lin_bwd(xq, dzq, w[3], w[4], pk[1], 1, ctx.need[0], dw=zs[1],
                                   dx_f32=_f32_dx(ctx.tok_dtypes[0], ctx.drops[0]))
            if A > 0:
                dVa, dga, da = lin_bwd(xa, dza, w[6], w[7], pk[2], 1, ctx.need[1], dw=zs[2],
                                       dx_f32=_f32_dx(ctx.tok_dtypes[1], ctx.drops[1]))
        grads = [dVv, dgv, dbv, dVq, dgq, dbq]
        if A > 0:
            grads += [dVa, dga, dba]
            da = _finish_dx(da, ctx.drops[1], ctx.tok_dtypes[1], (B, A, -1))
        dq = _finish_dx(dq, ctx.drops[0], ctx.tok_dtypes[0], (B, Q, -1))
        return (None, None, None, None, dq, da, dw if ctx.need[2] else None, *grads)


# --------------------------------------------------------------------------- #
class BiLogitsFn(Function):
    """BCNet.forward, h_out <= 32 branch (reference src/bc.py:52-58), with the optional zero-row mask
    of BiAttention (src/attention.py:36-37) written as -inf.  hmat (G, C) is the effective h_mat."""

    @staticmethod
    def forward(ctx, dims, packs, drops, v_bf16, rowmask, q, hmat, hbias, *w):
        """drops: None (eval) or (v_f32_2d, dv, dq, datt): input dropout of v_net / q_net and the attention dropout on
        the v projection (reference src/bc.py:53)."""
        B, K, Q, G, C = dims
        pk = [packs[i] if packs is not None else pack_layer(w[3 * i], w[3 * i + 1], 1) for i in range(2)]
        dq_drop = datt = None
        if drops is not None:
            v_f32, dv, dq_drop, datt = drops
            if dv is not None:
                v_bf16 = K_.dropout_bf16(v_bf16, dv)      # the cached bf16 cast is the source: half the bytes of a cast + dropout pass over fp32 v
        ctx.drops = (dq_drop, datt)
        xq = cast_tokens(q, dq_drop)
        vb, _ = lin_fwd(v_bf16, pk[0], w[2], True)
        if datt is not None:
            vb = K_.dropout_bf16(vb, datt)             # the kernels' (vb > 0) mask then also carries the dropout mask
        qb, _ = lin_fwd(xq, pk[1], w[5], True)
        hm = hmat.detach().reshape(G, C).contiguous()
        hb = hbias.detach().reshape(G).contiguous()
        logits = K_.bilinear_fwd(vb, qb, hm, hb, rowmask, B, K, Q, G, C)
        ctx.save_for_backward(v_bf16, xq, vb, qb, hm, hmat, hbias, *w)
        ctx.pk = pk
        ctx.dims = dims
        ctx.need_dq = q.requires_grad
        ctx.q_dtype = q.dtype
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        B, K, Q, G, C = ctx.dims
        v_bf16, xq, vb, qb, hm, hmat, hbias = ctx.saved_tensors[:7]
        w = ctx.saved_tensors[7:]
        pk = ctx.pk
        dzv, dzq, dbv, dbq, dh, dhb = K_.bilinear_bwd(vb, qb, hm, dlogits.contiguous(), B, K, Q, G, C)
        dq_drop, datt = ctx.drops
        sc = 1.0 if datt is None else 1.0 / (1.0 - datt[0])
        dVv, dgv, _ = lin_bwd(v_bf16, dzv, w[0], w[1], pk[0], 1, False, alpha=sc)
        if datt is not None:
            dbv = dbv * sc
        dVq, dgq, dq = lin_bwd(xq, dzq, w[3], w[4], pk[1], 1, ctx.need_dq, dx_f32=_f32_dx(ctx.q_dtype, dq_drop))
        dq = _finish_dx(dq, dq_drop, ctx.q_dtype, (B, Q, -1))
        return (None, None, None, None, None, dq, dh.view_as(hmat), dhb.view_as(hbias), dVv, dgv, dbv, dVq, dgq, dbq)


# --------------------------------------------------------------------------- #
class GRUFn(Function):
    """One-layer unidirectional GRU over a short sequence, zero initial state: ``nn.GRU(in, H, 1, batch_first=True)``
    as used by QuestionEmbedding.forward_all (reference src/language_model.py:56-61,93-98).  x (B, T, in) fp32 ->
    all hidden states (B, T, H) fp32.  x W_ih^T for every timestep is one GEMM; each step is one GEMM (h W_hh^T) and
    one pointwise kernel; backward is BPTT with one accumulating GEMM + one pointwise kernel per step and three GEMMs
    over the stacked steps for dW_hh, dW_ih, dx.  Gate order r, z, n (PyTorch's)."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, packs):
        B, T, Din = x.shape
        H = w_hh.shape[1]
        dev = x.device
        wih_b, whh_b = packs if packs is not None else gru_pack(w_ih, w_hh)
        Dp = wih_b.shape[1]                                   # in_dim padded to a multiple of 8 (TMA pitch)
        x2 = x.detach().reshape(B * T, Din)
        if Dp != Din:
            x2 = _pad_cols(x2, Dp)
        xb = cast_in(x2, None)
        bih, bhh = b_ih.detach().contiguous(), b_hh.detach().contiguous()
        _, gx = K_.gemm(xb, wih_b, B * T, 3 * H, Dp, bias=bih, out_bf16=False, out_f32=True)
        gx = gx.view(B, T, 3 * H)
        out = torch.empty((B, T, H), dtype=F32, device=dev)
        hb = torch.zeros((T + 1, B, H), dtype=BF16, device=dev)          # hb[t] = h_{t-1} in bf16; hb[0] = 0
        gates = torch.empty((4, T, B, H), dtype=BF16, device=dev)        # r, z, n, gh_n per step
        for t in range(T):
            _, gh = K_.gemm(hb[t], whh_b, B, 3 * H, H, bias=bhh, out_bf16=False, out_f32=True)
            K_.gru_gate_fwd(gx[:, t], gh, out[:, t - 1] if t > 0 else None, out[:, t], hb[t + 1], gates[0, t], gates[1, t],
                            gates[2, t], gates[3, t])
        ctx.save_for_backward(xb, hb, gates, out, wih_b, whh_b, w_ih, w_hh)
        ctx.dims = (B, T, Din, Dp, H)
        ctx.need_dx = x.requires_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        xb, hb, gates, out, wih_b, whh_b, w_ih, w_hh = ctx.saved_tensors
        B, T, Din, Dp, H = ctx.dims
        dev = dout.device
        dout = dout.contiguous() if dout.stride(2) != 1 else dout
        if dout.dtype != F32:
            dout = dout.float()
        dgx = torch.empty((B, T, 3 * H), dtype=BF16, device=dev)
        dgh = torch.empty((T, B, 3 * H), dtype=BF16, device=dev)
        dh, dwhh, dwih, dbih, dbhh = zero_slab(dev, [(B, H), (3 * H, H), (3 * H, Dp), (3 * H,), (3 * H,)])
        splits_h = _pick_splits(-(-B // 128) * -(-H // 256), 3 * H // 64)
        for t in range(T - 1, -1, -1):
            K_.gru_gate_bwd(dh, dout[:, t], out[:, t - 1] if t > 0 else None, gates[0, t], gates[1, t], gates[2, t],
                            gates[3, t], dgx[:, t], dgh[t])
            if t > 0:                                          # dh_{t-1} += dgh_t W_hh   (W_hh stored [3H][H] = MN-major B)
                K_.gemm(dgh[t], whh_b, B, H, 3 * H, b_mn=True, accum_f32=dh, k_splits=splits_h, tile_n=256)
        dgx2, dgh2 = dgx.view(B * T, 3 * H), dgh.view(T * B, 3 * H)
        hprev = hb[:T].view(T * B, H)
        kb = -(-B * T // 64)
        K_.gemm(dgh2, hprev, 3 * H, H, T * B, a_mn=True, b_mn=True, accum_f32=dwhh,
                k_splits=_pick_splits(-(-3 * H // 128) * -(-H // 256), kb), tile_n=256)
        K_.gemm(dgx2, xb, 3 * H, Dp, B * T, a_mn=True, b_mn=True, accum_f32=dwih,
                k_splits=_pick_splits(-(-3 * H // 128) * -(-Dp // 256), kb), tile_n=256)
        K_.act_bwd_bias(dgx2, None, False, dbih)
        K_.act_bwd_bias(dgh2, None, False, dbhh)
        dx = None
        if ctx.need_dx:
            _, dx = K_.gemm(dgx2, wih_b, B * T, Dp, 3 * H, b_mn=True, out_bf16=False, out_f32=True)
            dx = (dx if Dp == Din else dx[:, :Din]).reshape(B, T, Din)
        return dx, (dwih if Dp == Din else dwih[:, :Din]), dwhh, dbih, dbhh, None


def gru_pack(w_ih: torch.Tensor, w_hh: torch.Tensor):
    """bf16 copies of the GRU weights; the input width is zero-padded to a multiple of 8 (300-d embeddings)."""
    Din = w_ih.shape[1]
    Dp = -(-Din // 8) * 8
    return (K_.cast_rows(_pad_cols(w_ih.detach(), Dp).contiguous())[0], K_.cast_rows(w_hh.detach().contiguous())[0])
