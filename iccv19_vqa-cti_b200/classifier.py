"""Drop-in for the reference's ``src/classifier.py``: ``SimpleClassifier`` (reference src/classifier.py:11-29), the step
right after the hot path: ``weight_norm(Linear(in, hid))`` -> activation ->
``Dropout(p, inplace)`` -> ``weight_norm(Linear(hid, out))``.

Same constructor (``in_dim, hid_dim, out_dim, args`` with ``args.activation`` / ``args.dropout``), same ``main``
Sequential indices, hence the same state_dict keys (``main.0.*``, ``main.3.*``).  Both layers run on the tcgen05 GEMM
with the weight-norm scale folded into the bf16 pack; bias and ReLU live in the first GEMM's epilogue and the hidden
dropout is the second layer's fused input dropout (Philox mask regenerated in backward).  ``swish`` keeps its
activation as a stock elementwise op between the two GEMMs (it is not used by the shipped configurations).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functions as F_
from .fc import WNLinear


class _Swish(nn.Module):                        # reference src/activation.py: x * sigmoid(x)
    def forward(self, x):
        return x * torch.sigmoid(x)


class SimpleClassifier(nn.Module):
    def __init__(self, in_dim, hid_dim, out_dim, args):
        super().__init__()
        activation = getattr(args, "activation", None)
        if activation not in ("relu", "swish"):
            raise AssertionError(str(activation) + " is not supported yet!")
        self._relu = activation == "relu"
        self.main = nn.Sequential(
            WNLinear(in_dim, hid_dim),
            nn.ReLU() if self._relu else _Swish(),
            nn.Dropout(args.dropout, inplace=True),
            WNLinear(hid_dim, out_dim),
        )

    def forward(self, x):
        shape = x.shape
        y = x.reshape(-1, shape[-1])
        l0, l1 = self.main[0], self.main[3]
        h = F_.WNLinearFn.apply(y, l0.v_in(), l0.weight_g, l0.bias, self._relu, l0.packed(), None)
        if not self._relu:
            h = self.main[1](h)
        logits = F_.WNLinearFn.apply(h, l1.v_in(), l1.weight_g, l1.bias, False, l1.packed(),
                                     F_.new_drop(self.main[2].p, self.training))
        return logits.view(*shape[:-1], logits.shape[-1])
