"""Batch-sharded data parallelism for the CTI hot path: one process per GPU, parameters
replicated, rows sharded, and exactly one collective -- a sum all-reduce of the parameter
gradients, bucketed and launched from autograd hooks so it overlaps the rest of backward.

This is the collective the reference's ``Trainer._all_reduce_and_rescale`` names but never issues
(reference src/MC/trainer.py:208-219: it only flattens, divides and clips).  NCCL over NVLink 5 /
NVSwitch on GPUs; the same code runs on gloo for the CPU tests.

``transport="peer"`` moves the bytes on the copy engines over NVLink peer memory instead (``PeerRegion``,
csrc/peer.cu): no SM-resident transfer kernel, so the transfer overlaps backward without displacing the path's
persistent kernels.  NCCL then only carries the rendezvous (the exchange of the IPC handles).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib


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


class _DeviceSpan:
    """A raw device allocation presented through __cuda_array_interface__ (torch.as_tensor aliases it)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


_CREDIT = 512            # flag word: "the peer's staging buffer is free" (two-rank exchange of PeerRegion.all_reduce)


class _NoWork:
    def wait(self) -> None:
        pass


class _PeerWork:
    def __init__(self, region: "PeerRegion"):
        self.region = region

    def wait(self) -> None:
        self.region.stamp("main: before join")
        torch.cuda.current_stream().wait_stream(self.region.stream)
        self.region.stamp("main: joined")


class PeerRegion:
    """This rank's share of the peer-memory gradient sum: one device allocation -- flag block | slab of ``n_floats``
    gradients | two staging buffers of the same size (+ chunk rounding) -- exported to the other ranks of the node through CUDA IPC, plus their regions
    mapped here.  ``all_reduce(o, n)`` enqueues the sum of ``slab[o:o+n]`` over the ranks on the region's side stream:

        push my copy of chunk p into rank p's staging (copy engine)  ->  barrier  ->  chunk[me] += staged copies (rank
        order)  ->  push the reduced chunk into every peer's slab (copy engine)  ->  barrier

    (two ranks: push the whole range, barrier, both ranks add -- see ``all_reduce``).

    The caller orders the side stream after the producers (``stream.wait_stream``) and its consumers after the side
    stream (``_PeerWork.wait``).  All of it is capturable in a CUDA graph; a replayed barrier needs no host argument
    (its epoch lives in the flag block).  ``timeout_s``: a barrier whose peer does not arrive gives up and records the
    peer (``check()`` raises) instead of spinning forever."""

    def __init__(self, n_floats: int, device: torch.device, group: Optional[dist.ProcessGroup] = None,
                 timeout_s: float = 20.0, barrier: str = "memops"):
        """barrier: "memops" -- stream memory operations (no resident kernel: a rank waiting for a slower one holds no SM
        resources; no timeout) or "spin" -- a one-CTA kernel that spins on the flags and gives up after ``timeout_s``."""
        if device.type != "cuda":
            raise RuntimeError("PeerRegion: the peer-memory transport needs CUDA devices (use transport='nccl' / gloo)")
        self.lib = _lib.load()
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 16:
            raise RuntimeError("PeerRegion: at most 16 ranks (one node)")
        if barrier not in ("memops", "spin"):
            raise ValueError("PeerRegion: barrier must be 'memops' or 'spin'")
        self.barrier_kind = barrier
        self._slot = 0
        self.n = n_floats
        self.flag_bytes = 4096                                    # CTI_PEER_FLAG_BYTES
        self.timeout_s = timeout_s
        # Set-up is all-or-nothing ACROSS ranks: a rank whose allocation / export / import fails still takes part in the
        # exchanges below, and every rank raises if any did (the caller can then fall back to transport="nccl" everywhere).
        err = None
        ptr = ctypes.c_void_p()
        raw = None
        self.n_pad = n_floats + 64 * self.world                 # one staging buffer (chunk rounding: 4 floats per rank)
        with torch.cuda.device(device):
            try:
                _lib.check(self.lib.cti_peer_alloc(self.flag_bytes + 4 * (n_floats + 2 * self.n_pad), ctypes.byref(ptr)),
                           "cti_peer_alloc")
                handle = ctypes.create_string_buffer(64)
                _lib.check(self.lib.cti_peer_export(ptr, handle), "cti_peer_export")
                raw = handle.raw
            except RuntimeError as exc:
                err = exc
            handles = [None] * self.world
            dist.all_gather_object(handles, raw, group=group)
            self.bases: List[int] = []
            try:
                for r, h in enumerate(handles):
                    if h is None:
                        raise RuntimeError(f"PeerRegion: rank {r} could not export its region")
                    if r == self.rank:
                        self.bases.append(ptr.value)
                    else:
                        q = ctypes.c_void_p()
                        _lib.check(self.lib.cti_peer_import(ctypes.create_string_buffer(h, 64), ctypes.byref(q)),
                                   "cti_peer_import")
                        self.bases.append(q.value)
                # the two-rank exchange starts with one credit: the peer's staging is free
                if self.barrier_kind == "memops":
                    _lib.check(self.lib.cti_peer_flag_op(ptr.value, _CREDIT, 1, 0,
                                                         torch.cuda.current_stream(device).cuda_stream), "cti_peer_flag_op")
                torch.cuda.synchronize(device)
            except RuntimeError as exc:
                err = err or exc
            oks = [None] * self.world
            dist.all_gather_object(oks, err is None, group=group)
            if not all(oks):
                raise RuntimeError(f"PeerRegion: peer-memory set-up failed on rank(s) {[r for r, ok in enumerate(oks) if not ok]}"
                                   + (f" (here: {err})" if err is not None else ""))
            self.stream = torch.cuda.Stream(device)
            self.copy_streams = [torch.cuda.Stream(device) for _ in range(min(self.world - 1, 4))] if self.world > 2 else []
        self.trace: Optional[torch.Tensor] = None                 # debug: int64 time stamps (start_trace / stamp)
        self.trace_labels: List[str] = []
        self._blocks = (ctypes.c_void_p * self.world)(*self.bases)
        self._span = _DeviceSpan(ptr.value + self.flag_bytes, n_floats)
        self.slab = torch.as_tensor(self._span, device=device)
        dist.barrier(group=group)                                  # every region is mapped everywhere before first use

    def _slab(self, r: int, o: int) -> int:
        return self.bases[r] + self.flag_bytes + 4 * o

    def _staging(self, r: int, o: int) -> int:
        return self.bases[r] + self.flag_bytes + 4 * (self.n + o)

    def start_trace(self, slots: int = 128) -> None:
        """Debug: record %globaltimer at the protocol's phase boundaries (and wherever ``stamp`` is called) of the steps
        issued -- or captured -- from now on; ``read_trace()`` returns (label, ns since the first stamp)."""
        self.trace = torch.zeros(slots, dtype=torch.int64, device=self.slab.device)
        self.trace_labels = []

    def stamp(self, label: str, stream: Optional[torch.cuda.Stream] = None) -> None:
        if self.trace is None or len(self.trace_labels) >= self.trace.numel():
            return
        st = (stream or torch.cuda.current_stream()).cuda_stream
        _lib.check(self.lib.cti_peer_stamp(self.trace.data_ptr() + 8 * len(self.trace_labels), st), "cti_peer_stamp")
        self.trace_labels.append(label)

    def read_trace(self):
        t = self.trace.cpu().tolist()[:len(self.trace_labels)]
        t0 = min(x for x in t if x) if any(t) else 0
        return sorted(((lab, x - t0) for lab, x in zip(self.trace_labels, t)), key=lambda z: z[1])

    def barrier(self, slot: int = 0, second: bool = False) -> None:
        """Memory-operation barriers reset their flags by the waiter, so two consecutive barriers -- in ANY order graphs
        are replayed in -- must not share a slot: the first barrier of an exchange takes an even slot, the second
        (``second=True``) the odd one next to it.  (The two-rank exchange has one barrier; its credit orders the re-use.)"""
        if self.barrier_kind == "memops":
            if not second:
                self._slot = (self._slot + 2) % 8
            _lib.check(self.lib.cti_peer_barrier_memops(self._blocks, self.rank, self.world, self._slot + int(second),
                                                        self.stream.cuda_stream), "cti_peer_barrier_memops")
            return
        _lib.check(self.lib.cti_peer_barrier(self._blocks, self.rank, self.world, slot, self.timeout_s,
                                             self.stream.cuda_stream), "cti_peer_barrier")

    def _copies(self, copies) -> None:
        """Enqueue (dst, src, bytes) copies behind the side stream's work and order the side stream after them.  More
        than one copy (more than two ranks): spread over the copy streams, so several copy engines work at once (the pieces
        are 1 / world of a bucket -- for the last bucket of a step a latency-bound megabyte each)."""
        copies = [c for c in copies if c[2] > 0]
        if len(copies) <= 1 or not self.copy_streams:
            for d, s_, b in copies:
                _lib.check(self.lib.cti_peer_copy(d, s_, b, self.stream.cuda_stream), "cti_peer_copy")
            return
        start = torch.cuda.Event()
        start.record(self.stream)
        used = self.copy_streams[:min(len(copies), len(self.copy_streams))]
        for cs in used:
            cs.wait_event(start)
        for i, (d, s_, b) in enumerate(copies):
            _lib.check(self.lib.cti_peer_copy(d, s_, b, used[i % len(used)].cuda_stream), "cti_peer_copy")
        for cs in used:
            done = torch.cuda.Event()
            done.record(cs)
            self.stream.wait_event(done)

    def all_reduce(self, o: int, n: int) -> None:
        if o % 4 or n % 4 or o < 0 or o + n > self.n:
            raise ValueError("PeerRegion.all_reduce: the range must be 16-byte aligned and inside the slab")
        W, me, st, lib = self.world, self.rank, self.stream.cuda_stream, self.lib
        if W == 1 or n == 0:
            return
        tag = f"ar[{o}+{n}]"
        self.stamp(tag + " start", self.stream)
        if W == 2 and self.barrier_kind == "memops":
            # two ranks: the two-phase exchange moves n / 2 + n / 2 floats per direction; pushing the whole range at once
            # moves the same n, both ranks add the two copies in rank order (bit-identical sums), and one copy phase and
            # one barrier disappear.  The trailing barrier (the peer may not overwrite my staging before my sum has read it)
            # becomes a credit: after its sum each rank writes "staging free" into the peer's flag block, and waits for --
            # then takes -- its own credit before the next push.  The credit was usually written long before it is needed.
            peer = 1 - me
            # wait for "the peer's staging is free", take the credit: one batch
            _lib.check(lib.cti_peer_flag_ops(self.bases[me], (ctypes.c_int * 2)(_CREDIT, _CREDIT), (ctypes.c_uint32 * 2)(1, 0),
                                             (ctypes.c_int * 2)(1, 0), 2, st), "cti_peer_flag_ops")
            _lib.check(lib.cti_peer_copy(self._staging(peer, 0), self._slab(me, o), 4 * n, st), "cti_peer_copy")
            self.stamp(tag + " pushed", self.stream)
            self.barrier()
            self.stamp(tag + " barrier1", self.stream)
            _lib.check(lib.cti_sum_staged(self._slab(me, o), self._staging(me, 0), 1, me, n, self.n_pad, st), "cti_sum_staged")
            _lib.check(lib.cti_peer_flag_op(self.bases[peer], _CREDIT, 1, 0, st), "cti_peer_flag_op")   # my staging is free
            self.stamp(tag + " done", self.stream)
            return
        c = -(-(-(-n // W)) // 4) * 4                              # chunk floats: ceil(n / W) rounded up to 16 bytes

        def chunk(p):
            lo = min(p * c, n)
            return lo, min(lo + c, n) - lo
        my_lo, my_n = chunk(me)
        copies = []
        for k in range(1, W):                                      # my copy of chunk p -> staging slot of rank p
            p = (me + k) % W
            lo, m = chunk(p)
            slot = me if me < p else me - 1
            copies.append((self._staging(p, slot * c), self._slab(me, o + lo), 4 * m))
        self._copies(copies)
        self.stamp(tag + " pushed", self.stream)
        self.barrier()
        self.stamp(tag + " barrier1", self.stream)
        _lib.check(lib.cti_sum_staged(self._slab(me, o + my_lo), self._staging(me, 0), W - 1, me, my_n, c, st), "cti_sum_staged")
        # the reduced chunk -> every peer's slab
        self._copies([(self._slab((me + k) % W, o + my_lo), self._slab(me, o + my_lo), 4 * my_n) for k in range(1, W)])
        self.stamp(tag + " gathered", self.stream)
        self.barrier(second=True)
        self.stamp(tag + " done", self.stream)

    def all_reduce_fused(self, o: int, n: int, stream: Optional[torch.cuda.Stream] = None) -> None:
        """The same sum as ONE kernel on ``stream`` (default: the current stream) -- ``cti_peer_allreduce_fused``: stores
        over NVLink instead of copy-engine nodes, flag rounds inside the kernel.  For the LAST exchange of a step, issued
        where nothing competes for the SMs (every CTA of the kernel spins on the peers' flags): ~3x lower latency than the
        copy-engine protocol, whose strength is that it overlaps with compute.  Uses the second staging buffer, so it never
        meets the copy-engine exchanges still in flight on the side stream."""
        if o % 4 or n % 4 or o < 0 or o + n > self.n:
            raise ValueError("PeerRegion.all_reduce_fused: the range must be 16-byte aligned and inside the slab")
        if self.world == 1 or n == 0:
            return
        st = stream or torch.cuda.current_stream()
        W = self.world
        slabs = (ctypes.c_void_p * W)(*[self._slab(r, o) for r in range(W)])
        stages = (ctypes.c_void_p * W)(*[self._staging(r, self.n_pad) for r in range(W)])
        self.stamp(f"fused[{o}+{n}] start", st)
        _lib.check(self.lib.cti_peer_allreduce_fused(self._blocks, slabs, stages, self.rank, W, n, self.timeout_s,
                                                     st.cuda_stream), "cti_peer_allreduce_fused")
        self.stamp(f"fused[{o}+{n}] done", st)

    def close(self) -> None:
        """Unmap the peers' regions and free this rank's (after a synchronize on every rank: nothing may still be in
        flight).  The slab tensor must not be used afterwards."""
        if not self.bases:
            return
        torch.cuda.synchronize(self.slab.device)
        with torch.cuda.device(self.slab.device):
            for r, b in enumerate(self.bases):
                if r != self.rank:
                    _lib.check(self.lib.cti_peer_close(b), "cti_peer_close")
            _lib.check(self.lib.cti_peer_free(self.bases[self.rank]), "cti_peer_free")
        self.bases = []

    def check(self) -> None:
        """Raise if a barrier of this rank ever timed out (call after a synchronize)."""
        v = ctypes.c_int(0)
        _lib.check(self.lib.cti_peer_error(self.bases[self.rank], ctypes.byref(v)), "cti_peer_error")
        if v.value:
            raise RuntimeError(f"PeerRegion: rank {v.value - 1} did not arrive at a barrier within {self.timeout_s} s")


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
                 param_groups: Optional[Sequence[Sequence[torch.nn.Parameter]]] = None, transport: str = "nccl"):
        """transport: "nccl" -- ``dist.all_reduce`` of the process group's backend (NCCL on GPUs, gloo in the CPU tests);
        "peer" -- copy-engine transfers over NVLink peer memory (``PeerRegion``; one node, CUDA, fp32 gradients).

        param_groups: optional explicit buckets, in the order backward completes them -- bucket i holds exactly the
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
        self.peer: Optional[PeerRegion] = None
        self.fused_tail = not os.environ.get("CTI_PEER_NO_FUSED_TAIL")   # reduce_now: last exchange as one kernel
        if transport not in ("nccl", "peer"):
            raise ValueError("GradAllReducer: transport must be 'nccl' or 'peer'")
        if ps and all(p.dtype == ps[0].dtype and p.device == ps[0].device for p in ps):
            n_slab = sum(_padded(p.numel()) for grp in groups for p in grp)
            if transport == "peer" and self.world > 1:
                if ps[0].dtype != torch.float32:
                    raise RuntimeError("GradAllReducer: transport='peer' sums fp32 gradients")
                self.peer = PeerRegion(n_slab, ps[0].device, process_group,
                                       barrier=os.environ.get("CTI_PEER_BARRIER", "memops"))
                self.slab = self.peer.slab
            else:
                self.slab = torch.zeros(n_slab, dtype=ps[0].dtype, device=ps[0].device)
        elif transport == "peer":
            raise RuntimeError("GradAllReducer: transport='peer' needs every gradient in one dtype on one device")
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
            b.work = self._all_reduce_async(b.flat)

    def _all_reduce_async(self, flat: torch.Tensor, last: bool = False):
        """Start the sum of ``flat`` (a piece of the slab) over the ranks; returns a handle with ``wait()``.
        last: nothing of the step is left to overlap with (the caller waits right away)."""
        if self.peer is not None and last and self.fused_tail:
            self.peer.all_reduce_fused((flat.data_ptr() - self.slab.data_ptr()) // 4, flat.numel())
            return _NoWork()
        if self.peer is not None:
            self.peer.stamp("main: fork")
            self.peer.stream.wait_stream(torch.cuda.current_stream())         # after the producers of these gradients
            self.peer.all_reduce((flat.data_ptr() - self.slab.data_ptr()) // 4, flat.numel())
            return _PeerWork(self.peer)
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

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
            b.work = self._all_reduce_async(b.flat)
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
                self._all_reduce_async(self.slab, last=True).wait()
            else:
                # buckets launched during backward (launch_bucket) are in flight; the rest -- consecutive pieces of the
                # slab -- go out as one more collective, then everything is waited for
                rest = [b for b in self.buckets if not b.launched]
                if rest:
                    if self.slab is not None and all(x.flat.data_ptr() + x.flat.numel() * x.flat.element_size() == y.flat.data_ptr()
                                                     for x, y in zip(rest, rest[1:])):
                        o = (rest[0].flat.data_ptr() - self.slab.data_ptr()) // self.slab.element_size()
                        tail = self.slab[o:o + sum(b.flat.numel() for b in rest)]
                        rest[0].work = self._all_reduce_async(tail, last=True)
                    else:
                        for b in rest:
                            b.work = self._all_reduce_async(b.flat)
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
        """Detach the hooks.  (A peer region stays mapped: the gradients live in it; ``self.peer.close()`` after a
        barrier across the ranks releases it.)"""
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
