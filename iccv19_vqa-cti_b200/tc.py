"""Drop-in for the reference's ``src/tc.py``: ``TCNet`` -- compact trilinear interaction
(reference src/tc.py:9-61, with the mode products of src/Tensor.py:3-19 folded into one kernel).

Constructor, attribute names, parameter names / shapes / registration order follow the reference,
so ``state_dict()`` round-trips with reference checkpoints.
"""
from __future__ import annotations


import torch
import torch.nn as nn

from . import functions as F_
from .fc import FCNet, cast_features, features_f32_2d


def _rows(v, q, a) -> int:
    """Batch rows of a call.  Extension of the reference interface: ``v`` may hold B / n images when each run of n
    consecutive rows of q / a asks about the same image -- the multiple-choice trainer clones every image once per
    answer candidate (reference src/MC/train.py:75-76); handing the un-cloned features in gives identical outputs
    (the clones' projections are identical) with the image-side GEMMs, casts and activations done once per image."""
    B = q.shape[0]
    if a.shape[0] != B or v.shape[0] < 1 or B % v.shape[0] != 0:
        raise RuntimeError(f"batch mismatch: v has {v.shape[0]} samples, q {q.shape[0]}, a {a.shape[0]} "
                           "(v must have as many rows as q and a, or a divisor of it)")
    return B


class TCNet(nn.Module):
    def __init__(self, v_dim, q_dim, a_dim, h_dim, h_out, rank, glimpse, act='ReLU', dropout=[.2, .5], k=1):
        super().__init__()
        self.v_dim = v_dim
        self.q_dim = q_dim
        self.a_dim = a_dim
        self.h_out = h_out
        self.rank = rank
        self.h_dim = h_dim * k
        self.hv_dim = int(h_dim / rank)
        self.hq_dim = int(h_dim / rank)
        self.ha_dim = int(h_dim / rank)
        self.glimpse = glimpse

        self.v_tucker = FCNet([v_dim, self.h_dim], act=act, dropout=dropout[1])
        self.q_tucker = FCNet([q_dim, self.h_dim], act=act, dropout=dropout[0])
        self.a_tucker = FCNet([a_dim, self.h_dim], act=act, dropout=dropout[0])
        if self.h_dim < 1024:                                       # reference src/tc.py:27
            # the reference builds a_tucker a second time in this branch (src/tc.py:28); doing the same keeps the RNG
            # consumption -- hence every initial weight under a given seed -- identical to the reference's
            self.a_tucker = FCNet([a_dim, self.h_dim], act=act, dropout=dropout[0])
            self.v_net = nn.ModuleList([FCNet([self.h_dim, self.hv_dim], act=act, dropout=dropout[1])
                                        for _ in range(rank)])
            self.q_net = nn.ModuleList([FCNet([self.h_dim, self.hq_dim], act=act, dropout=dropout[0])
                                        for _ in range(rank)])
            self.a_net = nn.ModuleList([FCNet([self.h_dim, self.ha_dim], act=act, dropout=dropout[0])
                                        for _ in range(rank)])
            if h_out > 1:
                self.ho_dim = int(h_out / rank)
                h_out = self.ho_dim
            self.T_g = nn.Parameter(torch.Tensor(1, rank, self.hv_dim, self.hq_dim, self.ha_dim, glimpse,
                                                 h_out).normal_())
        self.dropout = nn.Dropout(dropout[1])                       # unused, as in the reference (src/tc.py:38)
        self._rank_pack = None

    # ------------------------------------------------------------------ #
    def _rank_group(self, nets: nn.ModuleList):
        """One grouped layer out of the R per-rank FCNets: V (R*d, H), g (R,), b (R*d,)."""
        cache = self.__dict__.setdefault("_rank_lins", {})
        lins = cache.get(id(nets))
        if lins is None:
            lins = cache[id(nets)] = [n.single()[0] for n in nets]
        V = torch.cat([l.weight_v for l in lins], 0)
        g = torch.stack([l.weight_g for l in lins], 0)
        b = torch.cat([l.bias for l in lins], 0)
        return V, g, b

    def _check_trilinear(self):
        if not hasattr(self, 'T_g'):
            raise RuntimeError("TCNet.forward needs the per-rank nets and T_g (h_dim * k < 1024), "
                               "as in the reference (src/tc.py:27,47)")
        if self.T_g.shape[-1] != 1:
            raise RuntimeError("TCNet.forward: h_out > 1 is unusable in the reference too "
                               "(Tensor.ModeProduct cannot view the core)")
        if self.hv_dim != 16:
            raise RuntimeError(f"the sm_100a trilinear kernel is built for h_dim / rank == 16, got {self.hv_dim}")

    def _logits(self, v, q, a, rowmask_wanted: bool):
        self._check_trilinear()
        B, K = _rows(v, q, a), v.shape[1]
        Q, A = q.shape[1], a.shape[1]
        G = self.T_g.shape[5]
        v_bf16, rowmask = cast_features(v)
        lv, pv = self.v_tucker.single()
        lq, pq = self.q_tucker.single()
        la, pa = self.a_tucker.single()
        drops = None
        if self.training:
            # input dropout of the tucker nets and of the per-rank nets: one site per modality; functions.RANK_DROPOUT
            # decides whether the R per-rank nets draw independent masks from it (the reference's semantics, src/tc.py:29-31)
            sites = [F_.new_drop(p, True) for p in (pv, pq, pa, self.v_net[0].single()[1], self.q_net[0].single()[1],
                                                    self.a_net[0].single()[1])]
            if any(d is not None for d in sites):
                drops = (features_f32_2d(v) if sites[0] is not None else None, *sites)
        # weight packs are cached on the parameters' version counters (optimizer steps bump them)
        rank_params = self.__dict__.get("_rank_params")
        if rank_params is None:                      # module traversal is slow: list the per-rank parameters once
            rank_params = [p for nets in (self.v_net, self.q_net, self.a_net) for p in nets.parameters()]
            self.__dict__["_rank_params"] = rank_params
        key = tuple(p._version for p in rank_params) + (rank_params[0].data_ptr(),)
        cache = self._rank_pack                      # (key, [Packed] * 3, detached stacks or None)
        lz = self.__dict__.get("_rank_lazy")
        if lz is not None and torch.is_grad_enabled():
            lz[0].open_group(lz[1])                  # gradient group's node, created at its first use (prepack.open_group)
        proxy = self.__dict__.get("_rank_proxy")     # (key, [(V proxy, g stand-in, bias proxy)] * 3): set by prepack
        if (proxy is not None and proxy[0] == key and cache is not None and cache[0] == key
                and torch.is_grad_enabled()):
            # prepack deferred the weight-norm backward of the 3 x R per-rank nets: the stacked proxies stand for the 96
            # weight_v / bias parameters (no torch.cat of 96 tensors here, no 96-way split in backward)
            groups = proxy[1]
            rank_packs = cache[1]
        else:
            if cache is not None and cache[0] == key and cache[2] is not None and not torch.is_grad_enabled():
                groups = cache[2]
            else:
                groups = [self._rank_group(nets) for nets in (self.v_net, self.q_net, self.a_net)]
            if cache is None or cache[0] != key:
                cache = self._rank_pack = (key, [F_.pack_layer(V, g, self.rank) for V, g, _ in groups],
                                           [tuple(t.detach() for t in grp) for grp in groups])
            elif cache[2] is None:                   # packs came from prepack(): add the stacks
                cache = self._rank_pack = (key, cache[1], [tuple(t.detach() for t in grp) for grp in groups])
            rank_packs = [F_.Packed(pk.w, pk.sumsq) for pk in cache[1]]       # this path does its own weight-norm backward
        (Vvn, gvn, bvn), (Vqn, gqn, bqn), (Van, gan, ban) = groups
        packs = [lv.packed(), lq.packed(), la.packed()] + rank_packs
        dims = (B, K, Q, A, G, self.rank)
        return F_.TriLogitsFn.apply(dims, packs, drops, v_bf16, rowmask if rowmask_wanted else None, q, a, self.T_g,
                                    lv.v_in(), lv.weight_g, lv.bias, lq.v_in(), lq.weight_g, lq.bias,
                                    la.v_in(), la.weight_g, la.bias, Vvn, gvn, bvn, Vqn, gqn, bqn, Van, gan, ban)

    def forward(self, v, q, a):
        """v (B,K,v_dim), q (B,Q,q_dim), a (B,A,a_dim) -> trilinear logit map (B,K,Q,A,G).
        v may also hold B / n images when every run of n consecutive (q, a) rows belongs to one image (see _rows)."""
        return self._logits(v, q, a, False)

    def forward_with_weights(self, v, q, a, w):
        """Attention-weighted trilinear pooling: w (B,K,Q,A) -> joint embedding (B, h_dim)."""
        B, K = _rows(v, q, a), v.shape[1]
        Q, A = q.shape[1], a.shape[1]
        v_bf16, _ = cast_features(v)
        lv, pv = self.v_tucker.single()
        lq, pq = self.q_tucker.single()
        la, pa = self.a_tucker.single()
        drops = None
        if self.training:
            sites = [F_.new_drop(p, True) for p in (pv, pq, pa)]
            if any(d is not None for d in sites):
                drops = (features_f32_2d(v) if sites[0] is not None else None, *sites)
        packs = [lv.packed(), lq.packed(), la.packed()]
        dims = (B, K, Q, A, self.h_dim)
        return F_.PoolFn.apply(dims, packs, drops, v_bf16, q, a, w, lv.v_in(), lv.weight_g, lv.bias, lq.v_in(),
                               lq.weight_g, lq.bias, la.v_in(), la.weight_g, la.bias)
