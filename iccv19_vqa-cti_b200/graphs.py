"""CUDA-graph capture of a whole hot-path step (forward, backward, gradient all-reduce).

The path launches ~100 of its own kernels plus PyTorch's allocator / autograd bookkeeping per step;
issued eagerly from Python that is several milliseconds of host time, more than the kernels take.
Capturing the step once and replaying it removes the host from the loop (no tracing compiler: the
graph holds exactly the kernels the eager step launched, on the stream they were launched on).

Every launch goes through the C ABI on ``torch.cuda.current_stream()``, tensor maps are passed by
value, and all buffers come from torch's caching allocator (graph-private pool during capture), so
the step is capturable as is.  The only host-side state that must not leak into a capture are the
version-keyed caches (bf16 weight packs, the bf16 copy of the image features): they are dropped
before capture so that the pack / cast kernels are part of the graph and re-run on every replay.
"""
from __future__ import annotations

from typing import Callable, Iterable

import torch

from . import fc as _fc


def reset_caches(modules: Iterable[torch.nn.Module], tensors: Iterable[torch.Tensor] = ()) -> None:
    from .prepack import _PLANS
    for plan in _PLANS.values():
        plan.lazy = None                                   # proxies of an earlier step (they keep its autograd graph alive)
    for root in modules:
        for m in root.modules():
            if hasattr(m, "_pack"):
                m._pack = None
            if hasattr(m, "_vproxy"):
                m._vproxy = None
                m._vlazy = None
            if hasattr(m, "_rank_pack"):
                m._rank_pack = None
                m.__dict__["_rank_proxy"] = None
                m.__dict__["_rank_lazy"] = None
    for t in tensors:
        for attr in (_fc._FEAT_ATTR, "_cti_b200_tok"):
            if hasattr(t, attr):
                delattr(t, attr)


class GraphedStep:
    """``GraphedStep(step, modules, static_tensors)``: warm ``step`` up on a side stream, capture it, then
    ``replay()``.  ``step`` must read its inputs from static tensors (copy new data into them before a
    replay, or make the host-to-device copies part of ``step``) and must set ``p.grad = None`` itself if it
    runs a backward pass.  List every input tensor object the step hands to the modules in ``static_tensors``:
    their cached bf16 copies (image features, question / answer tokens) are dropped before capture so that the
    casts are part of the graph."""

    def __init__(self, step: Callable[[], object], modules: Iterable[torch.nn.Module],
                 static_tensors: Iterable[torch.Tensor] = (), warmup: int = 3, allow_fixed_dropout: bool = False,
                 capture_error_mode: str = "global"):
        """capture_error_mode: "thread_local" when the step issues NCCL collectives (``GradAllReducer.reduce_now()``
        inside the step): NCCL's watchdog thread queries events while the capture is open, which the default "global"
        mode turns into a hang (tools/nccl_graph_probe.py)."""
        modules, static_tensors = list(modules), list(static_tensors)
        # dropout sites (p, seed, offset) are host scalars baked into the captured launches: a graph captured in
        # train() mode would replay the SAME masks on every step.  Refuse unless the caller asks for exactly that.
        if not allow_fixed_dropout:
            for root in modules:
                for m in root.modules():
                    if isinstance(m, torch.nn.Dropout) and m.training and m.p > 0:
                        raise RuntimeError("GraphedStep: a module is in train() mode with dropout p > 0; a replay would "
                                           "reuse the captured dropout masks on every step.  Capture in eval() mode, run "
                                           "the training step eagerly, or pass allow_fixed_dropout=True")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        reset_caches(modules, static_tensors)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self.output = step()

    def replay(self):
        self.graph.replay()
        return self.output
