"""Drop-in for the reference's ``src/bc.py``: ``BCNet`` -- the bilinear connect network of BAN
(reference src/bc.py:16-78) that the distilled student uses.

``forward`` implements the ``h_out <= 32`` branch (the only one the shipped builders reach, via
``BiAttention``) as one fused kernel that applies ``h_mat[g]`` on the fly; ``forward_with_weights``
is the bilinear pooling.  The ``h_out is None`` and ``h_out > 32`` branches keep their parameters
(``h_net``) for state_dict compatibility but are not part of the accelerated path and raise.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functions as F_
from .fc import FCNet, WNLinear, cast_features, features_f32_2d


class BCNet(nn.Module):
    """Simple class for non-linear bilinear connect network (same signature as reference src/bc.py:19)."""

    def __init__(self, v_dim, q_dim, h_dim, h_out, act='ReLU', dropout=[.2, .5], k=1):
        super().__init__()
        self.c = 32
        self.k = k
        self.v_dim = v_dim
        self.q_dim = q_dim
        self.h_dim = h_dim
        self.h_out = h_out

        self.v_net = FCNet([v_dim, h_dim * self.k], act=act, dropout=dropout[0])
        self.q_net = FCNet([q_dim, h_dim * self.k], act=act, dropout=dropout[0])
        self.dropout = nn.Dropout(dropout[1])                     # attention dropout on v_ (src/bc.py:53)
        if 1 < k:
            self.p_net = nn.AvgPool1d(self.k, stride=self.k)
        if h_out is None:
            pass
        elif h_out <= self.c:
            self.h_mat = nn.Parameter(torch.Tensor(1, h_out, 1, h_dim * self.k).normal_())
            self.h_bias = nn.Parameter(torch.Tensor(1, h_out, 1, 1).normal_())
        else:
            self.h_net = WNLinear(h_dim * self.k, h_out)

    def _effective_h_mat(self):
        """h_mat, or its weight-normed form when BiAttention re-parametrised it (src/attention.py:19-20)."""
        if 'h_mat_v' in self._parameters:
            hv = self.h_mat_v
            return hv * (self.h_mat_g / hv.norm())                 # 2*3072 elements: parameter plumbing
        return self.h_mat

    def _h_params(self):
        """(h_mat (G, C), h_bias (G)) of the bilinear map.  h_out <= 32: the parameters themselves (src/bc.py:36-37);
        h_out > 32: the weight-normed ``h_net`` (src/bc.py:39), whose projection of the outer products
        ``h_net(v_ (x) q_)`` (:66) is the same bilinear form with h_mat = W_eff and h_bias = bias."""
        if self.h_out <= self.c:
            return self._effective_h_mat(), self.h_bias
        hn = self.h_net
        return hn.weight_v * (hn.weight_g / hn.weight_v.norm()), hn.bias      # (h_out, C) parameter plumbing

    def _logits(self, v, q, rowmask_wanted: bool):
        B, K = v.shape[0], v.shape[1]
        if q.shape[0] != B:
            raise RuntimeError(f"batch mismatch: v has {B} samples, q {q.shape[0]}")
        Q = q.shape[1]
        C = self.h_dim * self.k
        v_bf16, rowmask = cast_features(v)
        lv, pv = self.v_net.single()
        lq, pq = self.q_net.single()
        if self.h_out is None:
            # reference src/bc.py:42-47: sum_{k,q} v_[b,k,c] q_[b,q,c] -- the bilinear pooling with unit weights
            drops = None
            if self.training:
                sites = [F_.new_drop(p, True) for p in (pv, pq)]
                if any(d is not None for d in sites):
                    drops = (features_f32_2d(v) if sites[0] is not None else None, *sites, None)
            ones = torch.ones((B, K, Q), dtype=torch.float32, device=v.device)
            out = F_.PoolFn.apply((B, K, Q, 0, C), [lv.packed(), lq.packed()], drops, v_bf16, q, None, ones, lv.v_in(),
                                  lv.weight_g, lv.bias, lq.v_in(), lq.weight_g, lq.bias)
            return out.unsqueeze(1)
        drops = None
        if self.training:
            sites = [F_.new_drop(p, True) for p in (pv, pq, self.dropout.p)]
            if any(d is not None for d in sites):
                drops = (features_f32_2d(v) if sites[0] is not None else None, *sites)
        hmat, hbias = self._h_params()
        hmat, hbias = hmat.reshape(self.h_out, C), hbias.reshape(self.h_out)
        outs = []
        # one call takes up to 14 maps (tcgen05 path: 4); more maps -- h_out 15..32 and the h_net branch (src/bc.py:63-68),
        # which no shipped builder reaches -- run in chunks of 4 that share the dropout sites, i.e. the same masks (each
        # chunk repeats the two projections: correct, not fast)
        step = self.h_out if self.h_out <= 14 else 4
        for g0 in range(0, self.h_out, step):
            G = min(step, self.h_out - g0)
            outs.append(F_.BiLogitsFn.apply((B, K, Q, G, C), [lv.packed(), lq.packed()], drops, v_bf16,
                                            rowmask if rowmask_wanted else None, q, hmat[g0:g0 + G], hbias[g0:g0 + G],
                                            lv.v_in(), lv.weight_g, lv.bias, lq.v_in(), lq.weight_g, lq.bias))
        return outs[0] if len(outs) == 1 else torch.cat(outs, 1)

    def forward(self, v, q):
        """v (B,K,v_dim), q (B,Q,q_dim) -> bilinear logits (B, h_out, K, Q)."""
        return self._logits(v, q, False)

    def forward_with_weights(self, v, q, w):
        """Attention-weighted bilinear pooling: w (B,K,Q) -> (B, h_dim) (sum-pooled over k channel groups)."""
        B, K = v.shape[0], v.shape[1]
        if q.shape[0] != B:
            raise RuntimeError(f"batch mismatch: v has {B} samples, q {q.shape[0]}")
        Q = q.shape[1]
        C = self.h_dim * self.k
        v_bf16, _ = cast_features(v)
        lv, pv = self.v_net.single()
        lq, pq = self.q_net.single()
        drops = None
        if self.training:
            sites = [F_.new_drop(p, True) for p in (pv, pq)]
            if any(d is not None for d in sites):
                drops = (features_f32_2d(v) if sites[0] is not None else None, *sites, None)
        out = F_.PoolFn.apply((B, K, Q, 0, C), [lv.packed(), lq.packed()], drops, v_bf16, q, None, w, lv.v_in(),
                              lv.weight_g, lv.bias, lq.v_in(), lq.weight_g, lq.bias)
        if 1 < self.k:
            out = out.view(B, -1, self.k).sum(2)                   # AvgPool1d(k) * k (src/bc.py:75-77)
        return out
