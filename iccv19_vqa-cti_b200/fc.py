"""Drop-in for the reference's ``src/fc.py``: ``FCNet`` (reference src/fc.py:10-34).

Same constructor, same ``main`` Sequential indices and the same old-style weight-norm parameter
names (``main.N.bias``, ``main.N.weight_g`` of shape ``()``, ``main.N.weight_v``), so reference
checkpoints load unchanged.  The compute is the tcgen05 GEMM of libcti_sm100.so with the
weight-norm scale folded into a bf16 weight pack and bias + ReLU in the epilogue.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import functions as F_
from . import kernels as K_


class WNLinear(nn.Module):
    """Parameters of ``weight_norm(nn.Linear(in, out), dim=None)`` (reference src/fc.py:22,27):
    W = weight_v * weight_g / ||weight_v||_F.  Registration order bias, weight_g, weight_v matches
    what torch's legacy weight_norm leaves behind."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        v = torch.empty(out_features, in_features)
        nn.init.kaiming_uniform_(v, a=math.sqrt(5))                     # nn.Linear.reset_parameters
        bound = 1.0 / math.sqrt(in_features)
        self.bias = nn.Parameter(torch.empty(out_features).uniform_(-bound, bound))
        self.weight_g = nn.Parameter(v.norm().detach().clone())          # scalar, shape ()
        self.weight_v = nn.Parameter(v)
        self._pack: Optional[Tuple[tuple, F_.Packed]] = None
        self._vproxy = None         # (key, proxy of weight_v) while prepack defers this layer's weight-norm backward
        self._vlazy = None          # (plan, group) until the layer's gradient group has its node (prepack._Plan.open_group)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias=True"

    def v_in(self) -> torch.Tensor:
        """What the autograd functions take as this layer's ``weight_v``: the parameter itself, or -- after
        ``prepack`` with gradients enabled -- its proxy, an alias that carries dW_eff back to prepack's backward."""
        d = self._vproxy
        if d is not None and torch.is_grad_enabled() and d[0] == (self.weight_v._version, self.weight_g._version,
                                                                   self.weight_v.data_ptr()):
            lz = self.__dict__.get("_vlazy")
            if lz is not None:                  # first layer of its gradient group to run: create the group's node now
                lz[0].open_group(lz[1])
                d = self._vproxy
            return d[1]
        return self.weight_v

    def packed(self) -> F_.Packed:
        """bf16 W_eff, cached until weight_v / weight_g change (in-place updates bump ``_version``)."""
        key = (self.weight_v._version, self.weight_g._version, self.weight_v.data_ptr())
        if self._pack is None or self._pack[0] != key:
            self._pack = (key, F_.pack_layer(self.weight_v, self.weight_g, 1))
            self._vproxy = None
        pk = self._pack[1]
        if pk.dw is not None and (self._vproxy is None or self._vproxy[0] != key or not torch.is_grad_enabled()):
            return F_.Packed(pk.w, pk.sumsq)         # no live proxy: the layer does its own weight-norm backward
        return pk


def features_f32_2d(v: torch.Tensor) -> torch.Tensor:
    """(B, K, Dv) features as a contiguous (B*K, Dv) matrix for the fused cast + dropout kernel: fp32, or bf16 when the
    features arrive in the loader's bf16 wire format (the dropout kernel then works on bf16 directly)."""
    x = v.detach()
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return x.reshape(-1, x.shape[-1]).contiguous()


class FCNet(nn.Module):
    """Simple class for non-linear fully connect network (same signature as reference src/fc.py:13)."""

    def __init__(self, dims, act='ReLU', dropout=0):
        super().__init__()
        layers: List[nn.Module] = []
        self._plan = []                                   # (dropout p, index of the WNLinear in main, act name)
        for i in range(len(dims) - 1):
            if 0 < dropout:
                layers.append(nn.Dropout(dropout))
            layers.append(WNLinear(dims[i], dims[i + 1]))
            self._plan.append((float(dropout), len(layers) - 1, act))
            if '' != act:
                layers.append(getattr(nn, act)())
        self.main = nn.Sequential(*layers)

    def forward(self, x):
        shape = x.shape
        y = x.reshape(-1, shape[-1])
        for p, idx, act in self._plan:
            lin = self.main[idx]
            fused = act in ('', 'ReLU')
            y = F_.WNLinearFn.apply(y, lin.v_in(), lin.weight_g, lin.bias, act == 'ReLU', lin.packed(),
                                    F_.new_drop(p, self.training))
            if not fused:                                 # activations other than ReLU are not on the CTI path
                y = self.main[idx + 1](y)
        return y.view(*shape[:-1], y.shape[-1])

    # the single layer used by the fused TCNet / BCNet paths
    def single(self) -> Tuple[WNLinear, float]:
        if len(self._plan) != 1 or self._plan[0][2] != 'ReLU':
            raise RuntimeError("fused CTI kernels expect single-layer ReLU FCNets")
        p, idx, _ = self._plan[0]
        return self.main[idx], p


_FEAT_ATTR = "_cti_b200_feat"


def cast_features(v: torch.Tensor):
    """bf16 copy of the image features (B,K,Dv) as (B*K, Dv) plus the zero-row mask of
    reference src/attention.py:55, computed once per tensor and reused by every module of the same
    forward pass (the models hand the same ``v`` to the attention and to each glimpse's pooling)."""
    if v.dim() != 3:
        raise RuntimeError("image features must be (batch, regions, dim)")
    if not v.is_cuda:
        raise RuntimeError("cti_b200 modules run on CUDA tensors only (no CPU fallback)")
    cached = getattr(v, _FEAT_ATTR, None)
    key = (v._version, v.data_ptr())
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    x = v.detach()
    if x.dtype == torch.bfloat16:
        # loader wire format (loader.py): bf16 features need no cast; the mask comes from the loader (prime_features)
        # or from one read-only pass
        xb = x.reshape(-1, x.shape[-1]).contiguous()
        mask = K_.rowmask_bf16(xb)
    else:
        if x.dtype != torch.float32:
            x = x.float()
        xb, mask = K_.cast_rows(x.reshape(-1, x.shape[-1]).contiguous(), want_mask=True)
    try:
        setattr(v, _FEAT_ATTR, (key, xb, mask))
    except Exception:
        pass
    return xb, mask
