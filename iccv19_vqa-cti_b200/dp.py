"""Batch-sharded data parallelism for the CTI hot path: one process per GPU, parameters
replicated, rows sharded, and exactly one collective -- a sum all-reduce of the parameter
gradients, bucketed and launched from autograd hooks so it overlaps the rest of backward.

This is the collective the reference's ``Trainer._all_reduce_and_rescale`` names but never issues
(reference src/MC/trainer.py:208-219: it only flattens, divides and clips).  NCCL over NVLink 5 /
NVSwitch on GPUs; the same code runs on gloo for the CPU tests.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world: int, group: int = 1) -> slice:
    """Contiguous row range of this rank.  ``group`` keeps rows that belong together on one rank
    (MC folds the 4 answer candidates of a question into the batch, reference src/MC/train.py:75-79,
    and scores them in groups of 4, src/MC/trainer.py:297)."""
    if n_rows % group:
        raise ValueError(f"{n_rows} rows do not divide into groups of {group}")
    n_groups = n_rows // group
    base, extra = divmod(n_groups, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return slice(lo * group, hi * group)


_ALIGN = 4           # every gradient view starts on a 16-byte boundary (kernels write them with 16-byte stores)


def _padded(n: int) -> int:
    return -(-n // _ALIGN) * _ALIGN


class _Bucket:
    __slots__ = ("params", "flat", "views", "pending", "work", "launched")

    def __init__(self, params: List[torch.nn.Parameter], flat: Optional[torch.Tensor] = None):
        self.params = params
        n = sum(_padded(p.numel()) for p in params)
        self.flat = flat if flat is not None else torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
        self.views, o = [], 0
        for p in params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += _padded(p.numel())
        self.pending = len(params)
        self.work = None
        self.launched = False              # its all-reduce of this step is already in flight (launch_bucket)


class GradAllReducer:
    """Overlapped, bucketed gradient all-reduce.

        reducer = GradAllReducer(model.parameters())
        loss.backward()          # buckets are reduced as soon as their last gradient lands
        reducer.finish()         # wait; every p.grad is now the cross-rank SUM (or mean) -- a view
                                 # of the bucket's flat buffer, so a fused clip / optimizer can use
                                 # reducer.flat_grads() directly

    Buckets are filled in reverse registration order (the order backward produces gradients)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 25 << 20,
                 process_group: Optional[dist.ProcessGroup] = None, average: bool = False,
                 param_groups: Optional[Sequence[Sequence[torch.nn.Parameter]]] = None):
        """param_groups: optional explicit buckets, in the order backward completes them -- bucket i holds exactly the
        parameters of param_groups[i] (``launch_bucket(i)`` all-reduces it as soon as its producer is done, while the rest
        of backward runs); the remaining parameters are bucketed by size behind them."""
        self.group = process_group
        self.average = average
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        ps = [p for p in params if p.requires_grad]
        groups: List[List[torch.nn.Parameter]] = []
        explicit = set()
        for grp in (param_groups or ()):
            grp = [p for p in grp if p.requires_grad]
            if any(p in explicit for p in grp):
                raise ValueError("GradAllReducer: a parameter appears in two param_groups")
            explicit.update(grp)
            groups.append(grp)
        self.n_explicit = len(groups)
        if any(all(p is not q for q in ps) for p in explicit):
            raise ValueError("GradAllReducer: param_groups must be drawn from params")
        cur, cur_bytes = [], 0
        for p in reversed(ps):
            if p in explicit:
                continue
            if cur and (cur_bytes + p.numel() * p.element_size() > bucket_bytes or p.dtype != cur[0].dtype):
                groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
        if cur:
            groups.append(cur)
        # the buckets are consecutive pieces of ONE buffer when every gradient has the same dtype (always, on this path):
        # the hook-free path can then reduce everything with a single collective
        self.slab: Optional[torch.Tensor] = None
        if ps and all(p.dtype == ps[0].dtype and p.device == ps[0].device for p in ps):
            self.slab = torch.zeros(sum(_padded(p.numel()) for grp in groups for p in grp), dtype=ps[0].dtype,
                                    device=ps[0].device)
        self.buckets: List[_Bucket] = []
        o = 0
        for grp in groups:
            n = sum(_padded(p.numel()) for p in grp)
            self.buckets.append(_Bucket(grp, None if self.slab is None else self.slab[o:o + n]))
            o += n
        self._in_place = set()                            # params whose producer writes the gradient into the bucket view itself
        self._src = {}                                    # param -> gradient tensor produced by backward (graph-owned under replay)
        self._of = {}
        self._handles = []
        for b in self.buckets:
            for p in b.params:
                self._of[p] = b
                self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # ------------------------------------------------------------------ #
    def _launch(self, b: _Bucket) -> None:
        torch._foreach_copy_(b.views, [p.grad for p in b.params])        # one fused copy per bucket
        for p, v in zip(b.params, b.views):
            p.grad = v
        if self.world > 1:
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if not getattr(self, "_enabled", True):
            return
        b = self._of[p]
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def finish(self) -> None:
        """Wait for every bucket; parameters that received no gradient this step count as zeros."""
        for b in self.buckets:
            if b.launched:                                                # launch_bucket(): already in flight, in place
                continue
            if b.pending != 0:                                            # some grads never arrived (unused params)
                for p, v in zip(b.params, b.views):
                    if p.grad is None:
                        v.zero_()
                        p.grad = v
                self._launch(b)
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()
                b.work = None
            if self.average and self.world > 1:
                b.flat.div_(self.world)
            b.pending = len(b.params)
            b.launched = False

    def launch_bucket(self, i: int) -> None:
        """Start the all-reduce of bucket ``i`` NOW, asynchronously: every gradient of the bucket has just been written
        into its view by its producer (``prepack.bind_grad_buffers(..., groups=...)`` calls this from the backward node
        that finishes a group of layers), so the transfer overlaps the rest of backward.  ``reduce_now()`` then handles the
        other buckets and waits for all of them.  Capturable in a CUDA graph (capture_error_mode="thread_local")."""
        b = self.buckets[i]
        if any(p not in self._in_place for p in b.params):
            raise RuntimeError("launch_bucket: only for buckets whose gradients are all written in place (mark_in_place)")
        if self.world > 1 and getattr(self, "_collectives", True):
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        b.launched = True

    def reduce_now(self) -> None:
        """All-reduce the gradients backward has just produced, without hooks: used after a CUDA-graph replay of
        forward + backward (the replay re-fills the same gradient tensors, so the autograd hooks never fire).
        One fused copy into the flat buffer, ONE all-reduce of the whole buffer (the buckets are pieces of it), and
        ``p.grad`` is left pointing at the reduced bucket views -- no copy back.  The tensors backward writes into are
        remembered, so the next call (after the next replay, which refills them) finds its sources even though
        ``p.grad`` now names the views (or was reset to None); an eager backward simply assigns fresh ``p.grad`` tensors,
        which take over.  ``forget_sources()`` drops the remembered tensors (after re-capturing a graph)."""
        if getattr(self, "_enabled", True):
            raise RuntimeError("reduce_now() is the hook-free path: call set_hooks_enabled(False) first (with the hooks on, "
                               "backward already filled and reduced the buckets -- use finish())")
        dsts, srcs = [], []
        for b in self.buckets:
            for p, v in zip(b.params, b.views):
                if p in self._in_place:                                # already in the bucket (prepack.bind_grad_buffers)
                    continue
                g = p.grad
                if g is not None and g.data_ptr() != v.data_ptr():
                    self._src[p] = g                                   # fresh from backward
                src = self._src.get(p)                                 # p.grad is None or the view: the remembered tensor
                if src is None:
                    v.zero_()                                         # no gradient this step counts as zeros
                else:
                    dsts.append(v)
                    srcs.append(src)
        if dsts:
            torch._foreach_copy_(dsts, srcs)
        if self.world > 1 and getattr(self, "_collectives", True):
            early = any(b.launched for b in self.buckets)
            if self.slab is not None and not early:
                dist.all_reduce(self.slab, op=dist.ReduceOp.SUM, group=self.group)
            else:
                # buckets launched during backward (launch_bucket) are in flight; the rest -- consecutive pieces of the
                # slab -- go out as one more collective, then everything is waited for
                rest = [b for b in self.buckets if not b.launched]
                if rest:
                    if self.slab is not None and all(x.flat.data_ptr() + x.flat.numel() * x.flat.element_size() == y.flat.data_ptr()
                                                     for x, y in zip(rest, rest[1:])):
                        o = (rest[0].flat.data_ptr() - self.slab.data_ptr()) // self.slab.element_size()
                        tail = self.slab[o:o + sum(b.flat.numel() for b in rest)]
                        rest[0].work = dist.all_reduce(tail, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                    else:
                        for b in rest:
                            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                for b in self.buckets:
                    if b.work is not None:
                        b.work.wait()
                        b.work = None
        for b in self.buckets:
            b.launched = False
            if self.average and self.world > 1:
                b.flat.div_(self.world)
            for p, v in zip(b.params, b.views):
                p.grad = v
            b.pending = len(b.params)

    def mark_in_place(self, params) -> None:
        """These parameters' gradients are written straight into their bucket views by whoever produces them (the deferred
        weight-norm backward): ``reduce_now`` neither copies nor zeroes them."""
        self._in_place.update(params)

    def forget_sources(self) -> None:
        self._src.clear()

    def set_hooks_enabled(self, enabled: bool) -> None:
        self._enabled = enabled

    def set_collectives_enabled(self, enabled: bool) -> None:
        """False: ``launch_bucket`` / ``reduce_now`` do their local work but issue no collective (a pass that only one
        rank runs, e.g. per-kernel profiling, must not wait for the others)."""
        self._collectives = enabled

    def flat_grads(self) -> List[torch.Tensor]:
        return [b.flat for b in self.buckets]

    def zero_grad(self) -> None:
        """Drop gradients so the next backward assigns fresh ones (the hooks re-point them at the buckets)."""
        for b in self.buckets:
            for p in b.params:
                p.grad = None

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


def clip_flat_grads_(flats: Sequence[torch.Tensor], max_norm: float, denom: float = 1.0) -> torch.Tensor:
    """``grad /= denom`` then global-L2-norm clip, on the reduced flat buffers (what the reference does
    on its own flat copy: src/MC/trainer.py:213-214, src/utils.py:323-328).  Returns the pre-clip norm."""
    if denom != 1.0:
        torch._foreach_div_(list(flats), denom)
    norm = torch.linalg.vector_norm(torch.stack(torch._foreach_norm(list(flats))))
    if max_norm > 0:
        coef = (max_norm / (norm + 1e-6)).clamp(max=1.0)
        torch._foreach_mul_(list(flats), coef)
    return norm
