"""The reference trainer's update tail as three kernel launches over all parameters:

    flat_grads.div_(grad_denom); clip_grad_norm_(flat_grads, clip_norm); copy back     (src/MC/trainer.py:208-219)
    torch.optim.Adamax(...).step()                                                      (src/MC/train.py:32, trainer.py:252-256)

``FusedClipAdamax`` keeps the interface the trainer touches -- ``param_groups[0]['lr']`` (the warm-up / decay schedule
of src/MC/train.py:59-66 writes it), ``step()``, ``zero_grad()``, ``state_dict()`` -- and adds ``grad_denom`` /
``clip_norm`` to ``step`` so the flatten / clip / un-flatten passes disappear: gradients are read where autograd (or
the all-reduce buckets of ``dp.GradAllReducer``) left them.  fp32 state, same arithmetic as the reference; the
pre-clip gradient norm comes back as a device scalar (no host synchronisation inside ``step``).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from . import _lib
from . import kernels as K_

_CHUNK = 8192          # elements per thread block


class FusedClipAdamax:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 2e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 clip_norm: float = 0.25, modules=None):
        """modules: optional module (or list of modules) whose weight-norm packs are rebuilt right after every update
        (``prepack``: two launches for all layers instead of two per layer during the next forward pass)."""
        self.modules = modules
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedClipAdamax: parameters must be contiguous fp32 CUDA tensors (no CPU path)")
        dev = self.params[0].device
        self.param_groups = [{"params": self.params, "lr": lr, "betas": tuple(betas), "eps": eps, "clip_norm": clip_norm}]
        self.step_count = 0
        total = sum(p.numel() for p in self.params)
        self._state = torch.zeros(2 * total, dtype=torch.float32, device=dev)      # exp_avg | exp_inf, flat
        self.exp_avg, self.exp_inf, o = [], [], 0
        for p in self.params:
            n = p.numel()
            self.exp_avg.append(self._state[o:o + n].view_as(p))
            self.exp_inf.append(self._state[total + o:total + o + n].view_as(p))
            o += n
        chunk_tensor, chunk_start = [], []
        for t, p in enumerate(self.params):
            for s in range(0, p.numel(), _CHUNK):
                chunk_tensor.append(t)
                chunk_start.append(s)
        self.n_chunks = len(chunk_tensor)
        i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=dev)
        self._numel = i64([p.numel() for p in self.params])
        self._chunk_tensor = torch.tensor(chunk_tensor, dtype=torch.int32, device=dev)
        self._chunk_start = i64(chunk_start)
        self._p_ptrs = i64([p.data_ptr() for p in self.params])
        self._m_ptrs = i64([m.data_ptr() for m in self.exp_avg])
        self._u_ptrs = i64([u.data_ptr() for u in self.exp_inf])
        self._g_ptrs = torch.zeros(len(self.params), dtype=torch.int64, device=dev)
        # the gradient-pointer table reaches the device by an async copy from pinned memory; the host may run several
        # steps ahead of the GPU, so the staging buffer is double-buffered and each half is only rewritten after the
        # copy that last read it has completed (an event per half)
        self._g_host = [torch.zeros(len(self.params), dtype=torch.int64).pin_memory() for _ in range(2)]
        self._g_copied = [None, None]
        self._g_turn = 0
        self._g_key: Optional[tuple] = None
        self._partials = torch.empty(self.n_chunks, dtype=torch.float32, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)            # pre-clip norm of the last step

    # ------------------------------------------------------------------ #
    def _refresh_grad_table(self) -> None:
        grads = []
        for p in self.params:
            g = p.grad
            if g is None:
                raise RuntimeError("Model parameter did not receive gradient (reference src/MC/trainer.py:226-228)")
            if g.dtype != torch.float32 or not g.is_contiguous():
                raise RuntimeError("FusedClipAdamax: gradients must be contiguous fp32")
            grads.append(g.data_ptr())
        key = tuple(grads)
        if key != self._g_key:                   # bucket views / graph-owned gradients keep their addresses: no copy
            i = self._g_turn
            self._g_turn ^= 1
            if self._g_copied[i] is not None:
                self._g_copied[i].synchronize()  # the DMA that last read this half is done
            self._g_host[i].copy_(torch.tensor(grads, dtype=torch.int64))
            self._g_ptrs.copy_(self._g_host[i], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._g_copied[i] = ev
            self._g_key = key

    @torch.no_grad()
    def step(self, grad_denom: float = 1.0) -> torch.Tensor:
        """p.grad / grad_denom -> clip to ``clip_norm`` by global L2 norm -> Adamax.  Returns the pre-clip norm (device)."""
        grp = self.param_groups[0]
        self._refresh_grad_table()
        self.step_count += 1
        beta1, beta2 = grp["betas"]
        clr = grp["lr"] / (1.0 - beta1 ** self.step_count)
        lib = _lib.load()
        s = K_._stream()
        n = sum(p.numel() for p in self.params)
        K_._call("cti_grad_sumsq_multi", lib.cti_grad_sumsq_multi,
                 (self._g_ptrs.data_ptr(), self._numel.data_ptr(), self._chunk_tensor.data_ptr(),
                  self._chunk_start.data_ptr(), self.n_chunks, _CHUNK, self._partials.data_ptr(), self._sumsq.data_ptr(), s),
                 kernels=2, nbytes=4.0 * n)
        K_._call("cti_adamax_multi", lib.cti_adamax_multi,
                 (self._p_ptrs.data_ptr(), self._g_ptrs.data_ptr(), self._m_ptrs.data_ptr(), self._u_ptrs.data_ptr(),
                  self._numel.data_ptr(), self._chunk_tensor.data_ptr(), self._chunk_start.data_ptr(), self.n_chunks, _CHUNK,
                  self._sumsq.data_ptr(), 1.0 / float(grad_denom), float(grp["clip_norm"]), float(clr), float(beta1),
                  float(beta2), float(grp["eps"]), self.grad_norm.data_ptr(), s), nbytes=28.0 * n)
        # the kernel wrote through raw pointers: tell autograd / the weight-pack caches that the parameters changed
        torch.autograd.graph.increment_version(self.params)
        if self.modules is not None:
            from .prepack import prepack
            prepack(self.modules)
        return self.grad_norm

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def state_dict(self) -> dict:
        """The layout of ``torch.optim.Adamax.state_dict()`` -- what the reference saves as ``optimizer_state`` and
        reloads (src/MC/main.py:120, src/FFOE/main.py:127): ``state[i] = {step, exp_avg, exp_inf}`` per parameter in
        registration order, ``param_groups[0]`` with lr / betas / eps / weight_decay and the parameter indices.
        ``clip_norm`` rides along as an extra group key (torch ignores unknown keys on load)."""
        grp = self.param_groups[0]
        state = {}
        if self.step_count > 0:
            for i, (m, u) in enumerate(zip(self.exp_avg, self.exp_inf)):
                state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": m.clone(), "exp_inf": u.clone()}
        group = {"lr": grp["lr"], "betas": tuple(grp["betas"]), "eps": grp["eps"], "weight_decay": 0, "foreach": None,
                 "maximize": False, "differentiable": False, "capturable": False, "clip_norm": grp["clip_norm"],
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        """Accepts ``torch.optim.Adamax.state_dict()`` (torch 1.1 ... 2.x: ``step`` an int or a tensor) written by the
        reference trainer or by this class."""
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        if groups[0].get("weight_decay", 0) != 0:
            raise ValueError("FusedClipAdamax has no weight decay (the reference trains with weight_decay=0)")
        steps = set()
        self._state.zero_()
        for j, idx in enumerate(groups[0]["params"]):
            st = sd["state"].get(idx, sd["state"].get(str(idx)))
            if st is None:
                continue
            steps.add(int(st["step"].item() if torch.is_tensor(st["step"]) else st["step"]))
            self.exp_avg[j].copy_(st["exp_avg"])
            self.exp_inf[j].copy_(st["exp_inf"])
        if len(steps) > 1:
            raise ValueError(f"parameters carry different step counts {sorted(steps)}: the fused update keeps one")
        self.step_count = steps.pop() if steps else 0
        for k in ("lr", "betas", "eps", "clip_norm"):
            if k in groups[0]:
                self.param_groups[0][k] = tuple(groups[0][k]) if k == "betas" else groups[0][k]
