"""Opt-in fused form of the glimpse loop the reference's models run around the hot-path modules:

    for g in range(glimpse):                                                  # src/MC/base_model.py:145-148
        b_emb[g] = t_net[g].forward_with_weights(v, q_emb, ans_emb, att[:, :, :, :, g])     # src/FFOE/base_model.py:125-128
        q_emb = q_prj[g](b_emb[g].unsqueeze(1)) + q_emb
        ans_emb = a_prj[g](b_emb[g].unsqueeze(1)) + ans_emb
    q_emb = q_emb.sum(1) + ans_emb.sum(1)                                     # :150 / :130

``glimpse_joint(t_net, q_prj, a_prj, v, q_emb, ans_emb, att)`` returns that last ``q_emb`` (B, num_hid) and, on
request, the ``b_emb`` list.  Same modules, same parameters, same values; what changes is who runs the glue.  Written
with torch ops the loop costs ~50 launches and ~0.45 ms per 1024-row training step (residual adds over (B, T, D)
tensors, token sums, the zero-fill + strided-copy + add of every ``att[..., g]`` slice's backward, the adds that
combine the gradients of q_emb).  Here:

* the updated ``q_emb`` / ``ans_emb`` are never materialised: the next glimpse consumes only their bf16 copy
  (``cti_glimpse_residual_cast``), the final token sums re-add the residuals in registers (``cti_glimpse_token_sum``);
* ``q_prj[g]`` and ``a_prj[g]`` read the same ``b_emb[g]``: one cast, paired GEMM launches;
* backward: d q_emb / d ans_emb start as the broadcast joint gradient (``cti_glimpse_bcast_rows``) and every glimpse's
  dgrad GEMM accumulates into them (TMA reduce-add); the attention gradient of glimpse g is written by the pooling
  kernel straight into the g-slice of one (B, G, K*Q*A) buffer -- the layout the softmax backward reads.

The default drop-in path (the modules called one by one from the unmodified reference model) is untouched; this call is
what a maintainer swaps in for those six lines.  Training-mode dropout falls back to the module calls.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
from torch.autograd import Function

from . import functions as F_
from . import kernels as K_
from .fc import FCNet, cast_features
from .tc import TCNet, _rows

F32, BF16 = torch.float32, torch.bfloat16
N_PER_GLIMPSE = 15          # (weight_v, weight_g, bias) x (v_tucker, q_tucker, a_tucker, q_prj, a_prj)


class GlimpseLoopFn(Function):
    @staticmethod
    def forward(ctx, dims, packs, v_bf16, q, a, att, *w):
        B, K, Q, A, C, D, G = dims
        vr = (B * K) // v_bf16.shape[0]               # rows sharing one image (tc._rows)
        att_p = att.detach().permute(0, 4, 1, 2, 3)   # logical (B, G, K, Q, A); native layout of cti_b200.TriAttention
        xq, xa = F_.cast_tokens(q, None), F_.cast_tokens(a, None)
        qd, ad = q.detach().contiguous(), a.detach().contiguous()
        res_q: List[torch.Tensor] = []
        res_a: List[torch.Tensor] = []
        saved = []
        b_embs = []
        for g in range(G):
            pkv, pkq, pka, pkqp, pkap = packs[g]
            wg = w[g * N_PER_GLIMPSE:(g + 1) * N_PER_GLIMPSE]
            with K_.gemm_batch():
                vp, _ = F_.lin_fwd(v_bf16, pkv, wg[2], True)
                qp, _ = F_.lin_fwd(xq, pkq, wg[5], True)
                ap, _ = F_.lin_fwd(xa, pka, wg[8], True)
            wd = F_._sample_contiguous(att_p[:, g])
            if wd.dtype != F32:
                wd = wd.float()
            b_emb = K_.tri_pool_fwd(vp, qp, ap, wd, wd.stride(0), B, K, Q, A, C, vr)
            bb = K_.cast_rows(b_emb)[0]
            with K_.gemm_batch():                     # q_prj[g] and a_prj[g] read the same operand: one launch
                _, pq = F_.lin_fwd(bb, pkqp, wg[11], False, out_bf16=False, out_f32=True)
                _, pa = F_.lin_fwd(bb, pkap, wg[14], False, out_bf16=False, out_f32=True)
            res_q.append(pq)
            res_a.append(pa)
            saved += [xq, xa, vp, qp, ap, wd, bb]
            b_embs.append(b_emb)
            if g < G - 1:
                xq, xa = K_.glimpse_residual_cast(qd, res_q, ad, res_a)
        joint = K_.glimpse_token_sum(qd, res_q, ad, res_a)
        ctx.save_for_backward(v_bf16, *saved, *w)
        ctx.packs = packs
        ctx.dims = dims
        ctx.vr = vr
        ctx.need = (q.requires_grad, a.requires_grad, att.requires_grad)
        ctx.tok_dtypes = (q.dtype, a.dtype)
        ctx.mark_non_differentiable(*b_embs)
        return (joint, *b_embs)

    @staticmethod
    def backward(ctx, dj, *_unused):
        B, K, Q, A, C, D, G = ctx.dims
        t = ctx.saved_tensors
        v_bf16, saved, w = t[0], t[1:1 + 7 * G], t[1 + 7 * G:]
        packs = ctx.packs
        dev = dj.device
        dj = dj.contiguous()
        if dj.dtype != F32:
            dj = dj.float()
        eq, ea = K_.glimpse_bcast_rows(dj, Q, A)      # d q_emb, d ans_emb: every dgrad below accumulates into them
        eq2, ea2 = eq.view(B * Q, D), ea.view(B * A, D)
        datt = torch.empty((B, G, K * Q * A), dtype=F32, device=dev)
        grads = [None] * len(w)
        dbs = F_.zero_slab(dev, [(D,)] * (2 * G))     # bias gradients of q_prj / a_prj: one fill for the whole loop
        for g in range(G - 1, -1, -1):
            xq, xa, vp, qp, ap, wd, bb = saved[7 * g:7 * g + 7]
            wg = w[g * N_PER_GLIMPSE:(g + 1) * N_PER_GLIMPSE]
            pkv, pkq, pka, pkqp, pkap = packs[g]
            # d(q_prj[g] output) = sum over tokens of d q_emb_{g+1} (the broadcast add's backward)
            if g == G - 1:                            # nothing accumulated yet: Q (A) copies of the joint gradient
                dyq, dya = dj * float(Q), dj * float(A)
            else:
                dyq = K_.glimpse_token_sum(eq, [], None, [])
                dya = K_.glimpse_token_sum(ea, [], None, [])
            dzq_p = K_.act_bwd_bias(dyq, None, True, dbs[2 * g])
            dza_p = K_.act_bwd_bias(dya, None, True, dbs[2 * g + 1])
            with K_.gemm_batch():
                dVqp, dgqp, _ = F_.lin_bwd(bb, dzq_p, wg[9], wg[10], pkqp, 1, False)
                dVap, dgap, _ = F_.lin_bwd(bb, dza_p, wg[12], wg[13], pkap, 1, False)
            # d b_emb[g] = dz_q W_qprj + dz_a W_aprj: the second GEMM accumulates onto the first
            _, db_emb = K_.gemm(dzq_p, pkqp.w, B, C, pkqp.w.shape[0], b_mn=True, out_bf16=False, out_f32=True)
            K_.gemm(dza_p, pkap.w, B, C, pkap.w.shape[0], b_mn=True, accum_f32=db_emb, k_splits=1)
            dzv, dzq, dza, dbv, dbq, dba, _ = K_.tri_pool_bwd(vp, qp, ap, wd, wd.stride(0), db_emb, B, K, Q, A, C, ctx.vr,
                                                            dw_out=datt[:, g])
            with K_.gemm_batch():
                dVv, dgv, _ = F_.lin_bwd(v_bf16, dzv, wg[0], wg[1], pkv, 1, False)
                dVq, dgq, _ = F_.lin_bwd(xq, dzq, wg[3], wg[4], pkq, 1, False)
                dVa, dga, _ = F_.lin_bwd(xa, dza, wg[6], wg[7], pka, 1, False)
                K_.gemm(dzq, pkq.w, B * Q, D, C, b_mn=True, accum_f32=eq2, k_splits=1)
                K_.gemm(dza, pka.w, B * A, D, C, b_mn=True, accum_f32=ea2, k_splits=1)
            grads[g * N_PER_GLIMPSE:(g + 1) * N_PER_GLIMPSE] = [dVv, dgv, dbv, dVq, dgq, dbq, dVa, dga, dba,
                                                               dVqp, dgqp, dbs[2 * g], dVap, dgap, dbs[2 * g + 1]]
        dq = eq if ctx.need[0] else None
        da = ea if ctx.need[1] else None
        if dq is not None and dq.dtype != ctx.tok_dtypes[0]:
            dq = dq.to(ctx.tok_dtypes[0])
        if da is not None and da.dtype != ctx.tok_dtypes[1]:
            da = da.to(ctx.tok_dtypes[1])
        d_att = datt.view(B, G, K, Q, A).permute(0, 2, 3, 4, 1) if ctx.need[2] else None
        return (None, None, None, dq, da, d_att, *grads)


def _single_linear(net: FCNet):
    if len(net._plan) != 1:
        raise RuntimeError("glimpse_joint: q_prj / a_prj must be single-layer FCNets (reference src/MC/base_model.py:166-167)")
    p, idx, act = net._plan[0]
    if act != '':
        raise RuntimeError("glimpse_joint: q_prj / a_prj are built without activation in the reference (act='')")
    return net.main[idx], p


def _unfused(t_net, q_prj, a_prj, v, q_emb, ans_emb, att, want_b_emb):
    b_emb = []
    for g in range(len(t_net)):
        b_emb.append(t_net[g].forward_with_weights(v, q_emb, ans_emb, att[:, :, :, :, g]))
        q_emb = q_prj[g](b_emb[g].unsqueeze(1)) + q_emb
        ans_emb = a_prj[g](b_emb[g].unsqueeze(1)) + ans_emb
    joint = q_emb.sum(1) + ans_emb.sum(1)
    return (joint, b_emb) if want_b_emb else joint


def glimpse_joint(t_net: Sequence[TCNet], q_prj: Sequence[FCNet], a_prj: Sequence[FCNet], v: torch.Tensor,
                  q_emb: torch.Tensor, ans_emb: torch.Tensor, att: torch.Tensor, return_b_emb: bool = False):
    """The glimpse loop + final token sums of ``TanModel.forward`` / ``CTIModel.forward`` (see the module docstring).

    t_net / q_prj / a_prj: the model's ModuleLists (cti_b200 modules); v (B or B/n, K, v_dim); q_emb (B, Q, num_hid);
    ans_emb (B, A, num_hid); att (B, K, Q, A, G) as returned by ``TriAttention``.  -> (B, num_hid) fp32
    [, list of G detached b_emb (B, h_dim)].  Gradients flow to q_emb, ans_emb, att and every parameter."""
    G = len(t_net)
    if not (len(q_prj) == len(a_prj) == G) or att.shape[-1] != G:
        raise RuntimeError("glimpse_joint: t_net, q_prj, a_prj and the glimpse axis of att must have the same length")
    if G > 4:
        raise RuntimeError("glimpse_joint: at most 4 glimpses (the reference trains with 2)")
    lins, dropout_on = [], False
    for g in range(G):
        tn = t_net[g]
        trip = [tn.v_tucker.single(), tn.q_tucker.single(), tn.a_tucker.single(), _single_linear(q_prj[g]),
                _single_linear(a_prj[g])]
        mods = (tn, tn, tn, q_prj[g], a_prj[g])
        dropout_on |= any(m.training and p > 0 for m, (_, p) in zip(mods, trip))
        lins.append([l for l, _ in trip])
    if dropout_on:
        # every input dropout draws its own mask per call site: the module calls do that; same values as the reference
        return _unfused(t_net, q_prj, a_prj, v, q_emb, ans_emb, att, return_b_emb)
    B, K = _rows(v, q_emb, ans_emb), v.shape[1]
    Q, A, D = q_emb.shape[1], ans_emb.shape[1], q_emb.shape[2]
    C = t_net[0].h_dim
    for g in range(G):
        shapes = [tuple(l.weight_v.shape) for l in lins[g]]
        if (shapes[1] != (C, D) or shapes[2] != (C, ans_emb.shape[2]) or shapes[3] != (D, C) or shapes[4] != (D, C)
                or t_net[g].h_dim != C or ans_emb.shape[2] != D):
            raise RuntimeError(f"glimpse_joint: layer shapes of glimpse {g} do not chain ({shapes})")
    v_bf16, _ = cast_features(v)
    packs = [[l.packed() for l in lins[g]] for g in range(G)]
    w = [t for g in range(G) for l in lins[g] for t in (l.v_in(), l.weight_g, l.bias)]
    out = GlimpseLoopFn.apply((B, K, Q, A, C, D, G), packs, v_bf16, q_emb, ans_emb, att, *w)
    return (out[0], list(out[1:])) if return_b_emb else out[0]
