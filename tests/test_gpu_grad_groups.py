"""The deferred weight-norm backward split into groups of layers (prepack.bind_grad_buffers(..., groups=...)): each group
finishes its dV / dg as soon as its layers are done and launches its bucket's all-reduce, overlapping the rest of
backward.  On one GPU the collective is a no-op; the gradients must equal the ungrouped run bit for bit (same kernels,
same summation order; bias column sums, finished with float atomics, to rounding) and must live in the reducer's buckets."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

import cti_b200  # noqa: E402
from cti_b200.dp import GradAllReducer  # noqa: E402

DEV = "cuda"


def _build():
    torch.manual_seed(3)
    G = 2
    att = cti_b200.TriAttention(2048, 1024, 1024, 512, 1, 32, G, 1)
    pools = [cti_b200.TCNet(2048, 1024, 1024, 512, 1, 32, 1, k=2) for _ in range(G)]
    q_prj = [cti_b200.FCNet([1024, 1024], '', .2) for _ in range(G)]
    a_prj = [cti_b200.FCNet([1024, 1024], '', .2) for _ in range(G)]
    mods = torch.nn.ModuleList([att, *pools, *q_prj, *a_prj]).to(DEV).eval()
    return mods, att, pools, q_prj, a_prj


def _step(mods, att, pools, q_prj, a_prj, v, q, a, cot, reducer):
    for p in mods.parameters():
        p.grad = None
    cti_b200.prepack(mods)
    qq, aa = q.detach().requires_grad_(True), a.detach().requires_grad_(True)
    p_att, _ = att(v, qq, aa)
    qe, ae = qq, aa
    for g in range(len(pools)):
        b = pools[g].forward_with_weights(v, qe, ae, p_att[:, :, :, :, g])
        qe = q_prj[g](b.unsqueeze(1)) + qe
        ae = a_prj[g](b.unsqueeze(1)) + ae
    ((qe.sum(1) + ae.sum(1)) * cot).sum().backward()
    if reducer is not None:
        reducer.reduce_now()
    return {n: p.grad.clone() for n, p in mods.named_parameters()}


def _same(got, ref, name):
    """dV / dg come from the deterministic multi-tensor weight-norm pass: bit-equal.  Bias gradients and T_g are column
    sums finished with float atomics (order varies from run to run): equal to rounding."""
    if name.endswith("weight_v") or name.endswith("weight_g"):
        assert torch.equal(got, ref), name
    else:
        assert ((got - ref).norm() / (ref.norm() + 1e-20)).item() < 1e-5, name


def test_grouped_weight_norm_backward_equals_ungrouped_and_fills_the_buckets():
    mods, att, pools, q_prj, a_prj = _build()
    g = torch.Generator().manual_seed(1)
    B = 8
    v = torch.relu(torch.randn(B, 50, 2048, generator=g)).to(DEV)
    q = torch.tanh(torch.randn(B, 12, 1024, generator=g)).to(DEV)
    a = torch.tanh(torch.randn(B, 6, 1024, generator=g)).to(DEV)
    cot = torch.randn(B, 1024, generator=g).to(DEV)
    ref = _step(mods, att, pools, q_prj, a_prj, v, q, a, cot, None)

    groups = [[pools[1], q_prj[1], a_prj[1]], [pools[0], q_prj[0], a_prj[0]], [att]]
    pg = cti_b200.weight_norm_param_groups(mods, groups)
    assert [len(x) for x in pg] == [10, 10, 6 + 2 * 96]          # (weight_v, weight_g) x layers of each group
    reducer = GradAllReducer(list(mods.parameters()), param_groups=pg)
    reducer.set_hooks_enabled(False)
    launched = []
    orig = reducer.launch_bucket
    reducer.launch_bucket = lambda i: (launched.append(i), orig(i))[1]
    try:
        cti_b200.bind_grad_buffers(mods, reducer, groups=groups)
        for _ in range(2):                                        # second step: buckets re-used
            launched.clear()
            got = _step(mods, att, pools, q_prj, a_prj, v, q, a, cot, reducer)
            assert launched == [0, 1]      # in the order backward finishes the groups; the last one rides with reduce_now
            for n, r in ref.items():
                _same(got[n], r, n)
            lo = reducer.slab.data_ptr()
            hi = lo + reducer.slab.numel() * 4
            assert all(lo <= p.grad.data_ptr() < hi for p in mods.parameters())
        # captured: the group nodes and the reduce are part of the graph
        out = {}

        def fb():
            out.update(_step(mods, att, pools, q_prj, a_prj, v, q, a, cot, reducer))
        reducer.forget_sources()
        gs = cti_b200.GraphedStep(fb, [mods], [v, q, a], capture_error_mode="thread_local")
        gs.replay()
        torch.cuda.synchronize()
        worst = max(((p.grad - ref[n]).norm() / (ref[n].norm() + 1e-12)).item() for n, p in mods.named_parameters())
        assert worst < 1e-5, worst
    finally:
        cti_b200.bind_grad_buffers(mods, None)
    again = _step(mods, att, pools, q_prj, a_prj, v, q, a, cot, None)
    for n, r in ref.items():
        _same(again[n], r, n)
