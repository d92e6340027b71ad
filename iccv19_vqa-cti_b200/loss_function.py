"""Drop-in for the reference's ``src/loss_function.py``: ``Distillation_Loss`` (reference src/loss_function.py:12-25),
the training loss of the distilled BAN student and of the CTI free-form teacher (BASELINE configs 3 / 4):

    KLDiv(log_softmax(x / T), softmax(teacher / T)).sum(1).mean() * alpha T^2 + BCEWithLogits_sum(x, y) / B * (1 - alpha)

Same constructor and ``forward(input, knowledge, target)``.  Forward and gradient are one fused kernel pass
(``cti_kd_loss``); the teacher logits may be handed in as the fp16 tensor the reference stores them in
(src/FFOE/test.py:129) -- the loader-side wire format of SURVEY.md section 8f row 5 -- or as fp32.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib
from . import kernels as K_


class _KDLossFn(Function):
    @staticmethod
    def forward(ctx, x, teacher, target, T: float, alpha: float):
        if not x.is_cuda:
            raise RuntimeError("cti_b200.Distillation_Loss runs on CUDA tensors only (no CPU fallback)")
        B, N = x.shape
        xd = x.detach().float().contiguous()
        td = teacher.detach()
        if td.dtype not in (torch.float16, torch.float32):
            td = td.float()
        td = td.contiguous()
        yd = target.detach().float().contiguous()
        if td.shape != xd.shape or yd.shape != xd.shape:
            raise RuntimeError(f"Distillation_Loss: shapes differ: input {tuple(x.shape)}, knowledge {tuple(teacher.shape)}, "
                               f"target {tuple(target.shape)}")
        dx = torch.empty_like(xd)
        scratch = torch.empty((B + 1,), dtype=torch.float32, device=x.device)
        K_._call("cti_kd_loss", _lib.load().cti_kd_loss,
                 (xd.data_ptr(), td.data_ptr(), int(td.dtype == torch.float16), yd.data_ptr(), dx.data_ptr(),
                  scratch.data_ptr(), scratch[B:].data_ptr(), B, N, float(T), float(alpha), K_._stream()), kernels=2,
                 nbytes=float(B) * N * (4 + td.element_size() + 4 + 4))
        ctx.save_for_backward(dx)
        return scratch[B]

    @staticmethod
    def backward(ctx, dloss):
        (dx,) = ctx.saved_tensors
        return dx * dloss, None, None, None, None


class Distillation_Loss(nn.Module):
    def __init__(self, T, alpha):
        super().__init__()
        self.T = T
        self.alpha = alpha

    def forward(self, input, knowledge, target):
        return _KDLossFn.apply(input, knowledge, target, float(self.T), float(self.alpha))
