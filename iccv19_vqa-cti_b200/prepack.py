"""All weight-norm folds of a model in two kernel launches.

Every weight-normed layer keeps a bf16 pack ``W_eff = V g / ||V||_F`` that is rebuilt when its parameters change
(``WNLinear.packed``, ``TCNet``'s per-rank groups) -- lazily, one layer at a time: 2 launches per layer, 32 per
training step of the CTI hot path.  ``prepack(modules)`` rebuilds all of them at once (``cti_wn_pack_multi``) into
persistent buffers and primes the per-layer caches, so the forward pass that follows finds every pack ready.  Call it
once per step after the optimizer update (``FusedClipAdamax(..., modules=...)`` does), or at the top of a step that is
captured into a CUDA graph.  The packs are bit-identical to the lazily built ones.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.nn as nn

from . import _lib
from . import functions as F_
from . import kernels as K_
from .fc import WNLinear
from .tc import TCNet

_SEG, _BLK = 4096, 1024
_PLANS = {}


class _Plan:
    def __init__(self, roots: List[nn.Module]):
        self.singles: List[WNLinear] = []
        self.tcnets: List[TCNet] = []
        in_rank = set()
        for root in roots:
            for m in root.modules():
                if isinstance(m, TCNet) and hasattr(m, "v_net"):
                    self.tcnets.append(m)
                    for nets in (m.v_net, m.q_net, m.a_net):
                        for n in nets:
                            in_rank.update(id(x) for x in n.modules())
        for root in roots:
            for m in root.modules():
                # input widths that are not a multiple of 8 get zero-padded packs (functions.pack_layer): left to the
                # lazy per-layer path
                if isinstance(m, WNLinear) and id(m) not in in_rank and m.in_features % 8 == 0 \
                        and all(m is not s for s in self.singles):
                    self.singles.append(m)
        if not self.singles and not self.tcnets:
            raise RuntimeError("prepack: no weight-normed layers found")
        dev = (self.singles[0].weight_v if self.singles else self.tcnets[0].T_g).device
        if dev.type != "cuda":
            raise RuntimeError("prepack: the modules must live on a CUDA device (no CPU path)")
        self.device = dev
        v_ptrs, g_ptrs, w_ptrs, s_ptrs, elems = [], [], [], [], []
        self.single_packs, self.rank_packs = [], []
        for lin in self.singles:
            n, k = lin.weight_v.shape
            w = torch.zeros((-(-n // 8) * 8, k), dtype=torch.bfloat16, device=dev)      # zero rows pad odd widths
            ss = torch.zeros((1,), dtype=torch.float32, device=dev)
            self.single_packs.append(F_.Packed(w, ss))
            v_ptrs.append(lin.weight_v.data_ptr()); g_ptrs.append(lin.weight_g.data_ptr())
            w_ptrs.append(w.data_ptr()); s_ptrs.append(ss.data_ptr()); elems.append(n * k)
        for tc in self.tcnets:
            packs = []
            for nets in (tc.v_net, tc.q_net, tc.a_net):
                lins = [n.single()[0] for n in nets]
                d, h = lins[0].weight_v.shape
                w = torch.zeros((len(lins) * d, h), dtype=torch.bfloat16, device=dev)
                ss = torch.zeros((len(lins),), dtype=torch.float32, device=dev)
                packs.append(F_.Packed(w, ss))
                for r, lin in enumerate(lins):
                    v_ptrs.append(lin.weight_v.data_ptr()); g_ptrs.append(lin.weight_g.data_ptr())
                    w_ptrs.append(w.data_ptr() + 2 * r * d * h); s_ptrs.append(ss.data_ptr() + 4 * r)
                    elems.append(d * h)
            self.rank_packs.append(packs)
        if any(e % 4 for e in elems):
            raise RuntimeError("prepack: every layer must have a multiple of 4 elements")
        first_seg, n_seg, seg_entry, seg_index, blk_entry, blk_index = [], [], [], [], [], []
        for e, n in enumerate(elems):
            first_seg.append(len(seg_entry))
            ns = -(-n // _SEG)
            n_seg.append(ns)
            seg_entry += [e] * ns
            seg_index += list(range(ns))
            nb = -(-n // _BLK)
            blk_entry += [e] * nb
            blk_index += list(range(nb))
        i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=dev)
        i32 = lambda xs: torch.tensor(xs, dtype=torch.int32, device=dev)
        self.tables = (i64(v_ptrs), i64(g_ptrs), i64(w_ptrs), i64(s_ptrs), i64(elems), i32(first_seg), i32(n_seg),
                       i32(seg_entry), i32(seg_index), i32(blk_entry), i32(blk_index))
        self.n_segs, self.n_blks = len(seg_entry), len(blk_entry)
        self.partials = torch.empty((self.n_segs,), dtype=torch.float32, device=dev)
        self.total = sum(elems)
        self.ptr_key = tuple(v_ptrs)

    def still_valid(self) -> bool:                       # parameters re-allocated (.to(), load with assign) -> rebuild
        ptrs = [lin.weight_v.data_ptr() for lin in self.singles]
        for tc in self.tcnets:
            for nets in (tc.v_net, tc.q_net, tc.a_net):
                ptrs += [n.single()[0].weight_v.data_ptr() for n in nets]
        return tuple(ptrs) == self.ptr_key

    def run(self) -> None:
        t = self.tables
        K_._call("cti_wn_pack_multi", _lib.load().cti_wn_pack_multi,
                 (t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(), t[5].data_ptr(),
                  t[6].data_ptr(), t[7].data_ptr(), t[8].data_ptr(), self.n_segs, t[9].data_ptr(), t[10].data_ptr(),
                  self.n_blks, self.partials.data_ptr(), K_._stream()), kernels=2, nbytes=10.0 * self.total)
        for lin, pk in zip(self.singles, self.single_packs):
            lin._pack = ((lin.weight_v._version, lin.weight_g._version, lin.weight_v.data_ptr()), pk)
        for tc, packs in zip(self.tcnets, self.rank_packs):
            rank_params = tc.__dict__.get("_rank_params")
            if rank_params is None:
                rank_params = [p for nets in (tc.v_net, tc.q_net, tc.a_net) for p in nets.parameters()]
                tc.__dict__["_rank_params"] = rank_params
            key = tuple(p._version for p in rank_params) + (rank_params[0].data_ptr(),)
            tc._rank_pack = (key, list(packs), None)       # the stacked fp32 copies are rebuilt on demand (tc.py)


def prepack(modules) -> None:
    """Rebuild the bf16 weight packs of every weight-normed layer under ``modules`` (a module or an iterable of
    modules) in two launches and prime the layers' caches.  The packs live in persistent buffers: call it between the
    backward pass of one step and the forward pass of the next, never in between."""
    roots = [modules] if isinstance(modules, nn.Module) else list(modules)
    key = tuple(id(m) for m in roots)
    plan = _PLANS.get(key)
    if plan is None or not plan.still_valid():
        plan = _PLANS[key] = _Plan(roots)
    plan.run()
