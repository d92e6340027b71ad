"""Drop-in for the hot-path classes of the reference's ``src/attention.py``:
``TriAttention`` (reference src/attention.py:43-59) and ``BiAttention`` (:14-40).

Returned tensors have the reference's logical shapes and values; strides differ (documented in
DESIGN.md): attention maps are views of a (B, G, domain) buffer, so ``att[..., g]`` /
``att[:, g]`` -- what the model forwards slice -- is contiguous per sample.
"""
from __future__ import annotations

import torch.nn as nn

from . import functions as F_
from .bc import BCNet
from .tc import TCNet


class BiAttention(nn.Module):
    def __init__(self, x_dim, y_dim, z_dim, glimpse, dropout=[.2, .5]):
        super().__init__()
        self.glimpse = glimpse
        net = BCNet(x_dim, y_dim, z_dim, glimpse, dropout=dropout, k=3)
        # weight_norm(net, name='h_mat', dim=None) (reference src/attention.py:19-20): h_mat is replaced by the
        # scalar h_mat_g and the direction h_mat_v, registered after h_bias like the legacy hook does.
        h = net.h_mat.data
        del net._parameters['h_mat']
        net.register_parameter('h_mat_g', nn.Parameter(h.norm().clone()))
        net.register_parameter('h_mat_v', nn.Parameter(h.clone()))
        self.logits = net

    def forward(self, v, q, v_mask=True):
        """v: [batch, k, vdim], q: [batch, q_len, qdim] -> (p, logits), both (B, G, K, Q)."""
        return self.forward_all(v, q, v_mask)

    def forward_all(self, v, q, v_mask=True):
        logits = self.logits._logits(v, q, bool(v_mask))           # -inf already written at all-zero regions
        p = F_.MaskedSoftmaxFn.apply(logits, False)
        return p, logits


class TriAttention(nn.Module):
    def __init__(self, v_dim, q_dim, a_dim, h_dim, h_out, rank, glimpse, k, dropout=[.2, .5]):
        super().__init__()
        self.glimpse = glimpse
        self.TriAtt = TCNet(v_dim, q_dim, a_dim, h_dim, h_out, rank, glimpse, dropout=dropout, k=k)

    def forward(self, v, q, a):
        """-> (p, logits), both logical (B, K, Q, A, G); p sums to 1 over (k, q, a) for each (b, g);
        logits hold -inf at regions whose feature row is all zero (reference src/attention.py:55-56)."""
        if self.glimpse < 2:
            raise RuntimeError("TriAttention needs glimpse >= 2: with one glimpse the reference squeezes the "
                               "glimpse axis (src/tc.py:52) and its 5-d mask no longer fits (src/attention.py:55)")
        logits = self.TriAtt._logits(v, q, a, True)
        p = F_.MaskedSoftmaxFn.apply(logits, True)
        return p, logits
