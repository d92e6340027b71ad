"""All weight-norm folds of a model in two kernel launches.

Every weight-normed layer keeps a bf16 pack ``W_eff = V g / ||V||_F`` that is rebuilt when its parameters change
(``WNLinear.packed``, ``TCNet``'s per-rank groups) -- lazily, one layer at a time: 2 launches per layer, 32 per
training step of the CTI hot path.  ``prepack(modules)`` rebuilds all of them at once (``cti_wn_pack_multi``) into
persistent buffers and primes the per-layer caches, so the forward pass that follows finds every pack ready.  Call it
once per step after the optimizer update (``FusedClipAdamax(..., modules=...)`` does), or at the top of a step that is
captured into a CUDA graph.  The packs are bit-identical to the lazily built ones.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

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
        f32, bf16 = torch.float32, torch.bfloat16

        # ---- one entry per weight-normed Linear (a per-rank net is an entry that writes into its group's stacked pack)
        ent_v, ent_g, elems = [], [], []
        dw_elems = []                        # size of each persistent dW_eff buffer (one per single layer / per rank GROUP)
        for lin in self.singles:
            n, k = lin.weight_v.shape
            ent_v.append(lin.weight_v); ent_g.append(lin.weight_g); elems.append(n * k)
            dw_elems.append((-(-n // 8) * 8) * k)
        self.group_lins = []                 # [tcnet][3] -> list of the R WNLinears
        for tc in self.tcnets:
            per_tc = []
            for nets in (tc.v_net, tc.q_net, tc.a_net):
                lins = [n.single()[0] for n in nets]
                per_tc.append(lins)
                d, h = lins[0].weight_v.shape
                for lin in lins:
                    ent_v.append(lin.weight_v); ent_g.append(lin.weight_g); elems.append(d * h)
                dw_elems.append(len(lins) * d * h)
            self.group_lins.append(per_tc)
        if any(e % 4 for e in elems):
            raise RuntimeError("prepack: every layer must have a multiple of 4 elements")
        self.ent_v, self.ent_g, self.elems = ent_v, ent_g, elems
        self.n_entries = len(elems)

        # ---- persistent buffers: bf16 packs, squared norms, dW_eff accumulators (zeroed at the top of every step)
        self.dw_slab = torch.zeros(sum(-(-e // 4) * 4 for e in dw_elems), dtype=f32, device=dev)
        self.sumsq = torch.zeros(self.n_entries, dtype=f32, device=dev)
        self.single_packs, self.rank_packs = [], []
        self.rank_bias, self.rank_vproxy = [], []
        w_ptrs, s_ptrs, dw_ptrs = [], [], []
        o, e = 0, 0
        for lin in self.singles:
            n, k = lin.weight_v.shape
            npad = -(-n // 8) * 8
            w = torch.zeros((npad, k), dtype=bf16, device=dev)                           # zero rows pad odd widths
            dw = self.dw_slab[o:o + npad * k].view(npad, k)
            self.single_packs.append(F_.Packed(w, self.sumsq[e:e + 1], dw))
            w_ptrs.append(w.data_ptr()); s_ptrs.append(self.sumsq.data_ptr() + 4 * e); dw_ptrs.append(dw.data_ptr())
            o += -(-npad * k // 4) * 4
            e += 1
        for per_tc in self.group_lins:
            packs, biases, vps = [], [], []
            for lins in per_tc:
                R = len(lins)
                d, h = lins[0].weight_v.shape
                w = torch.zeros((R * d, h), dtype=bf16, device=dev)
                dw = self.dw_slab[o:o + R * d * h].view(R * d, h)
                packs.append(F_.Packed(w, self.sumsq[e:e + R], dw))
                biases.append(torch.zeros(R * d, dtype=f32, device=dev))
                vps.append(torch.empty((R * d, h), dtype=f32, device=dev))                # gradient carrier only: never read
                for r in range(R):
                    w_ptrs.append(w.data_ptr() + 2 * r * d * h); s_ptrs.append(self.sumsq.data_ptr() + 4 * (e + r))
                    dw_ptrs.append(dw.data_ptr() + 4 * r * d * h)
                o += R * d * h
                e += R
            self.rank_packs.append(packs)
            self.rank_bias.append(biases)
            self.rank_vproxy.append(vps)

        first_seg, n_seg, seg_entry, seg_index, blk_entry, blk_index = [], [], [], [], [], []
        for ei, n in enumerate(elems):
            first_seg.append(len(seg_entry))
            ns = -(-n // _SEG)
            n_seg.append(ns)
            seg_entry += [ei] * ns
            seg_index += list(range(ns))
            nb = -(-n // _BLK)
            blk_entry += [ei] * nb
            blk_index += list(range(nb))
        i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=dev)
        i32 = lambda xs: torch.tensor(xs, dtype=torch.int32, device=dev)
        self._s_ptrs, self._dw_ptrs = s_ptrs, dw_ptrs
        self.groups = None                   # set_grad_groups(): the backward of the fold split into groups of layers
        self.lazy = None                     # (proxies, opened groups) between prepack() and the end of the forward pass
        self.entry_group = None
        self.tables = (i64([t.data_ptr() for t in ent_v]), i64([t.data_ptr() for t in ent_g]), i64(w_ptrs), i64(s_ptrs),
                       i64(elems), i32(first_seg), i32(n_seg), i32(seg_entry), i32(seg_index), i32(blk_entry), i32(blk_index))
        self.dw_ptrs = i64(dw_ptrs)
        # dV / dg of one backward pass live in two fresh flat buffers: their pointer tables are offsets + base
        offs, oo = [], 0
        for n in elems:
            offs.append(oo)
            oo += n
        self.total = oo
        self.dv_off = i64([4 * x for x in offs])
        self.dg_off = i64([4 * x for x in range(self.n_entries)])
        self.dv_ptrs = torch.zeros(self.n_entries, dtype=torch.int64, device=dev)
        self.dg_ptrs = torch.zeros(self.n_entries, dtype=torch.int64, device=dev)
        self.dv_offs_host = offs
        self.n_segs, self.n_blks = len(seg_entry), len(blk_entry)
        self.partials = torch.empty((self.n_segs,), dtype=f32, device=dev)
        self.ptr_key = tuple(t.data_ptr() for t in ent_v)
        # flat argument list of the autograd node: [V, g] per entry, then the biases of the per-rank nets
        self.fn_inputs = [t for pair in zip(ent_v, ent_g) for t in pair]
        self.rank_bias_params = [[[lin.bias for lin in lins] for lins in per_tc] for per_tc in self.group_lins]
        self.fn_inputs += [b for per_tc in self.rank_bias_params for grp in per_tc for b in grp]
        self.targets = None

    def set_grad_targets(self, targets) -> None:
        """targets: {parameter: fp32 tensor of its shape} -- where the deferred backward writes dV / dg (e.g. the views of a
        gradient all-reduce bucket).  The backward then assigns ``p.grad = target`` itself: no copy into the bucket."""
        if targets is None:
            self.targets = None
            return
        missing = [1 for t in self.ent_v + self.ent_g if t not in targets]
        if missing:
            raise RuntimeError("prepack: every weight_v / weight_g of the plan needs a gradient target")
        self.targets = ([targets[v] for v in self.ent_v], [targets[g] for g in self.ent_g])
        i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=self.device)
        self.dv_ptrs_static = i64([t.data_ptr() for t in self.targets[0]])
        self.dg_ptrs_static = i64([t.data_ptr() for t in self.targets[1]])

    def entry_owners(self) -> List[WNLinear]:
        """The WNLinear behind every entry, in entry order (singles, then the per-rank nets of each TCNet)."""
        return list(self.singles) + [lin for per_tc in self.group_lins for lins in per_tc for lin in lins]

    def set_grad_groups(self, module_groups, callbacks) -> None:
        """Split the deferred weight-norm backward into groups of layers (each group = the layers under a list of modules,
        given in the order backward finishes them): group i gets its own autograd node, which runs the two-launch dV / dg
        pass for ITS layers as soon as they have all produced dW_eff, writes into the bound gradient targets and then calls
        ``callbacks[i]()`` -- e.g. the all-reduce of the bucket holding exactly those gradients, which then overlaps the
        rest of backward.  Needs ``set_grad_targets`` first.  ``module_groups = None`` returns to one node for everything."""
        if module_groups is None:
            self.groups = None
            self.lazy = None
            return
        if self.targets is None:
            raise RuntimeError("prepack: bind gradient targets before splitting the backward into groups")
        owners = self.entry_owners()
        taken = [None] * self.n_entries
        for gi, mods in enumerate(module_groups):
            inside = {id(m) for root in mods for m in root.modules()}
            for e, lin in enumerate(owners):
                if id(lin) in inside:
                    if taken[e] is not None:
                        raise RuntimeError("prepack: a layer belongs to two gradient groups")
                    taken[e] = gi
        if any(t is None for t in taken):
            raise RuntimeError("prepack: the gradient groups must cover every weight-normed layer of the plan")
        dev = self.device
        i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=dev)
        i32 = lambda xs: torch.tensor(xs, dtype=torch.int32, device=dev)
        n_single = len(self.singles)
        self.entry_group = taken
        self.groups = []
        for gi in range(len(module_groups)):
            ents = [e for e in range(self.n_entries) if taken[e] == gi]
            first_seg, n_seg, seg_entry, seg_index, blk_entry, blk_index = [], [], [], [], [], []
            for li, e in enumerate(ents):
                n = self.elems[e]
                first_seg.append(len(seg_entry))
                ns = -(-n // _SEG)
                n_seg.append(ns)
                seg_entry += [li] * ns
                seg_index += list(range(ns))
                nb = -(-n // _BLK)
                blk_entry += [li] * nb
                blk_index += list(range(nb))
            # proxies of the group, as positions in _PackAllFn's output tuple: single layers, then (V stand-in, bias stack)
            # of every per-rank group whose R entries belong to this gradient group
            outs, kinds = [], []
            for i in range(n_single):
                if taken[i] == gi:
                    outs.append(i); kinds.append(("single", i))
            e = n_single
            for ti, per_tc in enumerate(self.group_lins):
                for j, lins in enumerate(per_tc):
                    mine = {taken[e + r] for r in range(len(lins))}
                    if mine == {gi}:
                        base = n_single + 6 * ti + 2 * j
                        outs += [base, base + 1]
                        kinds += [("rank", ti, j, e, len(lins)), ("bias",)]
                    elif gi in mine:
                        raise RuntimeError("prepack: the per-rank nets of one modality must share a gradient group")
                    e += len(lins)
            self.groups.append({
                "entries": ents, "outs": outs, "kinds": kinds, "callback": callbacks[gi],
                "tables": (i64([self._dw_ptrs[e] for e in ents]), i64([self.ent_v[e].data_ptr() for e in ents]),
                           i64([self.ent_g[e].data_ptr() for e in ents]), i64([self._s_ptrs[e] for e in ents]),
                           i64([self.targets[0][e].data_ptr() for e in ents]), i64([self.targets[1][e].data_ptr() for e in ents]),
                           i64([self.elems[e] for e in ents]), i32(first_seg), i32(n_seg), i32(seg_entry), i32(seg_index),
                           i32(blk_entry), i32(blk_index)),
                "n_segs": len(seg_entry), "n_blks": len(blk_entry), "total": sum(self.elems[e] for e in ents),
                "partials": torch.empty((len(seg_entry),), dtype=torch.float32, device=dev)})

    def open_group(self, gi: int) -> None:
        """Create group gi's autograd node now -- called when the forward pass first asks for a proxy of the group
        (``WNLinear.v_in``, ``TCNet`` per-rank proxies).  The autograd engine runs ready nodes in the reverse of their
        creation order; a node created here, right before the group's first consumer, therefore runs as soon as the
        group's layers have all handed dW_eff back -- in the MIDDLE of backward, before the nodes of everything that
        came earlier in the forward pass.  Created inside ``prepack`` (before the whole forward) the nodes would run
        last, and the all-reduce they launch would overlap nothing."""
        if self.lazy is None or gi in self.lazy[1]:
            return
        proxies, opened = self.lazy
        opened.add(gi)
        grp = self.groups[gi]
        for pos, t in zip(grp["outs"], _GroupGradFn.apply(self, gi, *[proxies[o] for o in grp["outs"]])):
            proxies[pos] = t
        self.prime(proxies, only=gi)

    def grad_group(self, gi: int) -> None:
        """dV / dg of the layers of group gi (two launches), into the bound gradient targets."""
        g = self.groups[gi]
        t = g["tables"]
        K_._call("cti_wn_grad_multi", _lib.load().cti_wn_grad_multi,
                 (t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(), t[5].data_ptr(),
                  t[6].data_ptr(), t[7].data_ptr(), t[8].data_ptr(), t[9].data_ptr(), t[10].data_ptr(), g["n_segs"],
                  t[11].data_ptr(), t[12].data_ptr(), g["n_blks"], g["partials"].data_ptr(), K_._stream()),
                 kernels=2, nbytes=20.0 * g["total"])

    def still_valid(self) -> bool:                       # parameters re-allocated (.to(), load with assign) -> rebuild
        return tuple(t.data_ptr() for t in self.ent_v) == self.ptr_key

    # ------------------------------------------------------------------ #
    def pack(self) -> None:
        """The forward fold of every entry: two launches."""
        t = self.tables
        K_._call("cti_wn_pack_multi", _lib.load().cti_wn_pack_multi,
                 (t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(), t[5].data_ptr(),
                  t[6].data_ptr(), t[7].data_ptr(), t[8].data_ptr(), self.n_segs, t[9].data_ptr(), t[10].data_ptr(),
                  self.n_blks, self.partials.data_ptr(), K_._stream()), kernels=2, nbytes=10.0 * self.total)

    def grad(self, flat_dv: Optional[torch.Tensor], flat_dg: Optional[torch.Tensor]) -> None:
        """The backward of the fold for every entry: two launches, dW_eff read from the persistent accumulators; dV / dg
        go to the two flat buffers, or (both None) to the bound gradient targets."""
        if flat_dv is None:
            dv_ptrs, dg_ptrs = self.dv_ptrs_static, self.dg_ptrs_static
        else:
            torch.add(self.dv_off, flat_dv.data_ptr(), out=self.dv_ptrs)
            torch.add(self.dg_off, flat_dg.data_ptr(), out=self.dg_ptrs)
            dv_ptrs, dg_ptrs = self.dv_ptrs, self.dg_ptrs
        t = self.tables
        K_._call("cti_wn_grad_multi", _lib.load().cti_wn_grad_multi,
                 (self.dw_ptrs.data_ptr(), t[0].data_ptr(), t[1].data_ptr(), t[3].data_ptr(), dv_ptrs.data_ptr(),
                  dg_ptrs.data_ptr(), t[4].data_ptr(), t[5].data_ptr(), t[6].data_ptr(), t[7].data_ptr(), t[8].data_ptr(),
                  self.n_segs, t[9].data_ptr(), t[10].data_ptr(), self.n_blks, self.partials.data_ptr(), K_._stream()),
                 kernels=2, nbytes=20.0 * self.total)

    def prime(self, proxies, only: Optional[int] = None) -> None:
        """Point the layers' caches at the fresh packs (and, when ``proxies`` is given, at the proxies that defer their
        weight-norm backward).  only: just the layers of this gradient group (``open_group``)."""
        n_single = len(self.singles)
        pending = None                                     # gradient groups whose node does not exist yet (open_group)
        if proxies is not None and self.groups is not None and self.lazy is not None:
            pending = set(range(len(self.groups))) - self.lazy[1]
        e_rank = n_single
        for i, (lin, pk) in enumerate(zip(self.singles, self.single_packs)):
            if only is not None and self.entry_group[i] != only:
                continue
            key = (lin.weight_v._version, lin.weight_g._version, lin.weight_v.data_ptr())
            lin._pack = (key, pk)
            lin._vproxy = None if proxies is None else (key, proxies[i])
            lin._vlazy = None if pending is None or self.entry_group[i] not in pending else (self, self.entry_group[i])
        for ti, (tc, packs) in enumerate(zip(self.tcnets, self.rank_packs)):
            if only is not None:
                e0 = e_rank
                if self.entry_group[e0] != only:
                    e_rank += sum(len(lins) for lins in self.group_lins[ti])
                    continue
            rank_params = tc.__dict__.get("_rank_params")
            if rank_params is None:
                rank_params = [p for nets in (tc.v_net, tc.q_net, tc.a_net) for p in nets.parameters()]
                tc.__dict__["_rank_params"] = rank_params
            key = tuple(p._version for p in rank_params) + (rank_params[0].data_ptr(),)
            tc._rank_pack = (key, list(packs), None)       # the stacked fp32 copies are rebuilt on demand (tc.py)
            gi = self.entry_group[e_rank] if self.groups is not None else None
            e_rank += sum(len(lins) for lins in self.group_lins[ti])
            tc.__dict__["_rank_lazy"] = None if pending is None or gi not in pending else (self, gi)
            if proxies is None:
                tc.__dict__["_rank_proxy"] = None
            else:
                base = n_single + 6 * ti
                g_standin = self.sumsq[:1]                 # never read: dg comes from the deferred backward
                tc.__dict__["_rank_proxy"] = (key, [(proxies[base + 2 * j], g_standin, proxies[base + 2 * j + 1])
                                                    for j in range(3)])


class _PackAllFn(torch.autograd.Function):
    """One autograd node for the weight-norm re-parametrisation of EVERY layer of a model (reference src/fc.py:22,27:
    ``weight_norm(nn.Linear(...), dim=None)``).  forward: the two-launch fold into the bf16 packs; outputs are proxies --
    an alias of each single layer's ``weight_v``, and per group of R per-rank nets a stacked ``weight_v`` stand-in and the
    stacked bias.  The layers' autograd functions take the proxies in place of the parameters and return dW_eff / dbias for
    them; autograd calls backward here once every layer has done so: two launches produce all dV / dg."""

    @staticmethod
    def forward(ctx, plan, *params):
        plan.dw_slab.zero_()
        plan.pack()
        outs = [lin.weight_v.detach() for lin in plan.singles]
        with torch.no_grad():
            for per_tc, vps, biases in zip(plan.rank_bias_params, plan.rank_vproxy, plan.rank_bias):
                for grp, vp, bstack in zip(per_tc, vps, biases):
                    torch.cat([b.detach() for b in grp], 0, out=bstack)
                    outs += [vp.detach(), bstack.detach()]
        ctx.plan = plan
        ctx.set_materialize_grads(False)     # a proxy without a gradient arrives as None, not as a zero-filled tensor
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        plan = ctx.plan
        plan.lazy = None                     # the step's graph is done: do not keep its nodes alive through the proxies
        dev = plan.device
        n_single = len(plan.singles)
        used = [False] * plan.n_entries
        # incoming dW_eff normally IS the persistent accumulator (the layer wrote into it); anything else (a layer used
        # twice, so autograd summed two gradients) is copied in
        e = 0
        for i, pk in enumerate(plan.single_packs):
            g_ = grads[i]
            if g_ is not None:
                n = plan.singles[i].weight_v.shape[0]
                if g_.data_ptr() != pk.dw.data_ptr():
                    pk.dw[:n].copy_(g_)
                used[e] = True
            e += 1
        gi = n_single
        for packs in plan.rank_packs:
            for pk in packs:
                g_ = grads[gi]
                R = pk.sumsq.numel()
                if g_ is not None:
                    if g_.data_ptr() != pk.dw.data_ptr():
                        pk.dw.copy_(g_)
                    for r in range(R):
                        used[e + r] = True
                e += R
                gi += 2
        out = [None]
        if plan.groups is not None:
            out += [None, None] * plan.n_entries       # dV / dg were produced by the groups' own nodes (_GroupGradFn)
        elif plan.targets is not None:
            # gradients go straight into the bound targets (all-reduce bucket views); p.grad is assigned here, autograd
            # gets None for these parameters (an AccumulateGrad node would copy, or add a buffer to itself)
            plan.grad(None, None)
            for ei, (v, g) in enumerate(zip(plan.ent_v, plan.ent_g)):
                if used[ei]:
                    v.grad, g.grad = plan.targets[0][ei], plan.targets[1][ei]
                out += [None, None]
        else:
            flat_dv = torch.empty(plan.total, dtype=torch.float32, device=dev)
            flat_dg = torch.empty(plan.n_entries, dtype=torch.float32, device=dev)
            plan.grad(flat_dv, flat_dg)
            for ei, (v, off) in enumerate(zip(plan.ent_v, plan.dv_offs_host)):
                if used[ei]:
                    out += [flat_dv[off:off + v.numel()].view_as(v), flat_dg[ei]]
                else:
                    out += [None, None]
        gi = n_single
        for per_tc in plan.rank_bias_params:
            for grp in per_tc:
                db = grads[gi + 1]
                d = grp[0].numel()
                out += [None if db is None else db[r * d:(r + 1) * d] for r in range(len(grp))]
                gi += 2
        return tuple(out)


class _GroupGradFn(torch.autograd.Function):
    """Identity on the proxies of ONE group of layers (``_Plan.set_grad_groups``).  Its backward runs once all layers of
    the group have handed their dW_eff back -- in the middle of the backward pass, when the group is the last glimpse or
    the pooling of an earlier one -- finishes dV / dg for the group (two launches), assigns them as the parameters'
    gradients (they live in the all-reduce bucket) and fires the group's callback."""

    @staticmethod
    def forward(ctx, plan, gi, *proxies):
        ctx.plan, ctx.gi = plan, gi
        ctx.set_materialize_grads(False)
        return tuple(p.detach() for p in proxies)

    @staticmethod
    def backward(ctx, *grads):
        plan, grp = ctx.plan, plan_group(ctx)
        used = set()
        out = [None, None]
        for kind, g_ in zip(grp["kinds"], grads):
            if kind[0] == "bias":                   # stacked bias gradient of the per-rank nets: on to _PackAllFn.backward
                out.append(g_)
                continue
            out.append(None)
            if g_ is None:
                continue
            if kind[0] == "single":
                i = kind[1]
                pk = plan.single_packs[i]
                if g_.data_ptr() != pk.dw.data_ptr():
                    pk.dw[:plan.singles[i].weight_v.shape[0]].copy_(g_)
                used.add(i)
            else:
                _, ti, j, e0, R = kind
                pk = plan.rank_packs[ti][j]
                if g_.data_ptr() != pk.dw.data_ptr():
                    pk.dw.copy_(g_)
                used.update(range(e0, e0 + R))
        plan.grad_group(ctx.gi)
        for e in grp["entries"]:
            if e in used:
                plan.ent_v[e].grad, plan.ent_g[e].grad = plan.targets[0][e], plan.targets[1][e]
        if grp["callback"] is not None:
            grp["callback"]()
        return tuple(out)


def plan_group(ctx):
    return ctx.plan.groups[ctx.gi]


def prepack(modules) -> None:
    """Rebuild the bf16 weight packs of every weight-normed layer under ``modules`` (a module or an iterable of
    modules) in two launches and prime the layers' caches.  The packs live in persistent buffers: call it between the
    backward pass of one step and the forward pass of the next, never in between.

    With gradients enabled the call also DEFERS the weight-norm backward of those layers: they hand dW_eff to proxies
    of their parameters and one multi-tensor launch pair computes every dV / dg at the end of the backward pass
    (``_PackAllFn``) -- instead of two launches per layer and, for the 3 x R per-rank nets of a TCNet, a torch.cat of 96
    tensors in forward and its 96-way split in backward."""
    roots = [modules] if isinstance(modules, nn.Module) else list(modules)
    key = tuple(id(m) for m in roots)
    plan = _PLANS.get(key)
    if plan is None or not plan.still_valid():
        plan = _PLANS[key] = _Plan(roots)
    if torch.is_grad_enabled() and any(t.requires_grad for t in plan.fn_inputs):
        proxies = list(_PackAllFn.apply(plan, *plan.fn_inputs))
        # with gradient groups, each group's node (_GroupGradFn) is created lazily, when the forward pass first touches
        # one of the group's layers (open_group): see there
        plan.lazy = None if plan.groups is None else (proxies, set())
        plan.prime(proxies)
    else:
        plan.pack()
        plan.lazy = None
        plan.prime(None)


def weight_norm_param_groups(modules, groups) -> List[List[torch.nn.Parameter]]:
    """For ``dp.GradAllReducer(params, param_groups=...)``: per group of modules (``groups``: lists of modules under
    ``modules``, in the order backward finishes them) the weight_v / weight_g parameters of its weight-normed layers -- the
    gradients ``bind_grad_buffers(modules, reducer, groups=groups)`` produces group by group."""
    roots = [modules] if isinstance(modules, nn.Module) else list(modules)
    key = tuple(id(m) for m in roots)
    plan = _PLANS.get(key)
    if plan is None or not plan.still_valid():
        plan = _PLANS[key] = _Plan(roots)
    owners = plan.entry_owners()
    out = []
    for mods in groups:
        inside = {id(m) for root in mods for m in root.modules()}
        ents = [e for e, lin in enumerate(owners) if id(lin) in inside]
        out.append([plan.ent_v[e] for e in ents] + [plan.ent_g[e] for e in ents])
    return out


def bind_grad_buffers(modules, reducer, groups=None) -> None:
    """Let the deferred weight-norm backward write dV / dg of every layer under ``modules`` directly into the bucket views
    of ``reducer`` (a ``dp.GradAllReducer``): the gradient all-reduce then needs no copy of these tensors (97 % of the
    hot path's gradient bytes).  ``reducer = None`` unbinds.

    groups: lists of modules in the order backward finishes them (for the MC model: the last glimpse's pooling and
    projections, ..., the first glimpse's, the attention).  The reducer must have been built with
    ``param_groups=weight_norm_param_groups(modules, groups)``: bucket i then holds exactly group i's gradients, the
    weight-norm backward runs per group, and each group's all-reduce starts the moment the group is finished
    (``reducer.launch_bucket(i)``), overlapping the rest of backward."""
    roots = [modules] if isinstance(modules, nn.Module) else list(modules)
    key = tuple(id(m) for m in roots)
    plan = _PLANS.get(key)
    if plan is None or not plan.still_valid():
        plan = _PLANS[key] = _Plan(roots)
    if reducer is None:
        plan.set_grad_groups(None, None)
        plan.set_grad_targets(None)
        return
    targets = {p: v for b in reducer.buckets for p, v in zip(b.params, b.views)}
    plan.set_grad_targets(targets)
    reducer.mark_in_place(plan.ent_v + plan.ent_g)
    if groups is None:
        plan.set_grad_groups(None, None)
        return
    want = weight_norm_param_groups(roots, groups)
    callbacks = []
    for gi, ps in enumerate(want):
        b = reducer.buckets[gi] if gi < len(reducer.buckets) else None
        if b is None or len(b.params) != len(ps) or any(x is not y for x, y in zip(b.params, ps)):
            raise RuntimeError("bind_grad_buffers: build the reducer with param_groups=weight_norm_param_groups(modules, "
                               "groups) so that bucket i holds exactly the gradients of group i")
        # the last group finishes with backward itself: nothing left to overlap, so its bucket travels with the remaining
        # gradients in reduce_now()'s single transfer instead of paying for one of its own
        callbacks.append(None if gi == len(want) - 1 else (lambda i: (lambda: reducer.launch_bucket(i)))(gi))
    plan.set_grad_groups(groups, callbacks)
