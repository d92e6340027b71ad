"""Fused glimpse loop (cti_b200.glimpse_joint) against the same modules called one by one, as
the unmodified reference model calls them (src/MC/base_model.py:143-150), and against the fp32 oracle.

The fused call runs the same GEMM / pooling kernels on the same operands; the only arithmetic that moves is fp32
(residual adds re-associated in registers, gradient sums taken inside GEMM epilogues), so the two paths must agree far
inside the bf16 tolerance: forward 1e-5 relative, gradients 2e-3 in norm (different fp32 summation orders of split-K /
reduce-add), while each stays within the usual tolerance of the oracle.
"""
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

import cti_b200  # noqa: E402
from cti_b200 import kernels as KS  # noqa: E402
from oracle import cti_oracle as O  # noqa: E402

DEV = "cuda"


def normrel(x, y):
    x, y = x.detach().float().cpu(), y.detach().float().cpu()
    return ((x - y).norm() / (y.norm() + 1e-12)).item()


def build(v_dim, hid, h_mm, rank, G, seed=3):
    torch.manual_seed(seed)
    att = cti_b200.TriAttention(v_dim, hid, hid, h_mm, 1, rank, G, 1)
    pools = [cti_b200.TCNet(v_dim, hid, hid, h_mm, 1, rank, 1, k=2) for _ in range(G)]
    q_prj = [cti_b200.FCNet([hid, hid], '', .2) for _ in range(G)]
    a_prj = [cti_b200.FCNet([hid, hid], '', .2) for _ in range(G)]
    mods = torch.nn.ModuleList([att, *pools, *q_prj, *a_prj]).to(DEV).eval()
    return mods, att, pools, q_prj, a_prj


def inputs(B, K, Q, A, v_dim, hid, seed=5, rep=1):
    g = torch.Generator().manual_seed(seed)
    Bv = B // rep
    v = torch.relu(torch.randn(Bv, K, v_dim, generator=g))
    nb = torch.randint(K // 3, K + 1, (Bv,), generator=g)
    v = v * (torch.arange(K)[None, :] < nb[:, None]).float()[:, :, None]
    q = torch.tanh(torch.randn(B, Q, hid, generator=g))
    a = torch.tanh(torch.randn(B, A, hid, generator=g))
    cot = torch.randn(B, hid, generator=g)
    return v.to(DEV), q.to(DEV), a.to(DEV), cot.to(DEV)


def run(mods, att, pools, q_prj, a_prj, v, q, a, cot, fused, deferred):
    params = list(mods.parameters())
    for p in params:
        p.grad = None
    if deferred:
        cti_b200.prepack(mods)
    q = q.detach().requires_grad_(True)
    a = a.detach().requires_grad_(True)
    p_att, _ = att(v, q, a)
    if fused:
        joint = cti_b200.glimpse_joint(pools, q_prj, a_prj, v, q, a, p_att)
    else:
        qe, ae = q, a
        for g in range(len(pools)):
            b_emb = pools[g].forward_with_weights(v, qe, ae, p_att[:, :, :, :, g])
            qe = q_prj[g](b_emb.unsqueeze(1)) + qe
            ae = a_prj[g](b_emb.unsqueeze(1)) + ae
        joint = qe.sum(1) + ae.sum(1)
    (joint * cot).sum().backward()
    grads = {n: p.grad.clone() for n, p in mods.named_parameters() if p.grad is not None}
    return joint.detach(), q.grad.clone(), a.grad.clone(), grads


@pytest.mark.parametrize("B,K,Q,A,G,rep,deferred", [(8, 50, 12, 6, 2, 1, False), (8, 50, 12, 6, 2, 1, True),
                                                   (12, 37, 9, 3, 2, 4, True), (160, 50, 12, 6, 2, 1, True),
                                                   (6, 20, 7, 3, 3, 1, True)])
def test_fused_glimpse_loop_equals_module_calls(B, K, Q, A, G, rep, deferred):
    v_dim, hid, h_mm, rank = 2048, 1024, 512, 32
    mods, att, pools, q_prj, a_prj = build(v_dim, hid, h_mm, rank, G)
    v, q, a, cot = inputs(B, K, Q, A, v_dim, hid, rep=rep)
    j0, dq0, da0, g0 = run(mods, att, pools, q_prj, a_prj, v, q, a, cot, False, deferred)
    j1, dq1, da1, g1 = run(mods, att, pools, q_prj, a_prj, v, q, a, cot, True, deferred)
    assert normrel(j1, j0) < 1e-5, normrel(j1, j0)
    assert normrel(dq1, dq0) < 2e-3 and normrel(da1, da0) < 2e-3, (normrel(dq1, dq0), normrel(da1, da0))
    assert set(g0) == set(g1)
    # per parameter class (the 96 scalar weight_g of the per-rank nets are sums with cancellation: as single numbers a few
    # of them move by percents under any re-ordering of fp32 sums, as a vector they do not)
    classes = {}
    for n in sorted(g0):
        classes.setdefault(re.sub(r"_net\.\d+\.", "_net.*.", n), []).append(n)
    worst = max((normrel(torch.cat([g1[n].reshape(-1) for n in ns]), torch.cat([g0[n].reshape(-1) for n in ns])), c)
                for c, ns in classes.items())
    assert worst[0] < 5e-3, worst
    flat0 = torch.cat([g0[n].reshape(-1) for n in sorted(g0)])
    flat1 = torch.cat([g1[n].reshape(-1) for n in sorted(g0)])
    assert normrel(flat1, flat0) < 1e-3


def test_fused_glimpse_loop_against_oracle():
    """Same check the module path gets: joint embedding and the flat gradient against the fp32 oracle."""
    B, K, Q, A, G = 16, 50, 12, 6, 2
    mods, att, pools, q_prj, a_prj = build(2048, 1024, 512, 32, G, seed=11)
    v, q, a, cot = inputs(B, K, Q, A, 2048, 1024, seed=13)
    j1, dq1, da1, g1 = run(mods, att, pools, q_prj, a_prj, v, q, a, cot, True, True)
    # oracle on the same weights
    sd = {k: t.detach().cpu() for k, t in mods.state_dict().items()}
    names = ["v_att"] + [f"t_net.{g}" for g in range(G)] + [f"q_prj.{g}" for g in range(G)] + [f"a_prj.{g}" for g in range(G)]
    params = {}
    for k, t in sd.items():
        idx, rest = k.split(".", 1)
        params[f"{names[int(idx)]}.{rest}"] = t.clone().requires_grad_(True)
    vq, qq, aq = v.cpu(), q.cpu().requires_grad_(True), a.cpu().requires_grad_(True)
    joint, _, _ = O.cti_hot_path(vq, qq, aq, params, G)
    (joint * cot.cpu()).sum().backward()
    assert normrel(j1, joint) < 2e-3
    assert normrel(dq1, qq.grad) < 3e-2 and normrel(da1, aq.grad) < 3e-2


def test_glimpse_kernels_match_torch():
    g = torch.Generator().manual_seed(1)
    B, Tq, Ta, D = 37, 12, 6, 1024
    q = torch.randn(B, Tq, D, generator=g).to(DEV)
    a = torch.randn(B, Ta, D, generator=g).to(DEV)
    rq = [torch.randn(B, D, generator=g).to(DEV) for _ in range(3)]
    ra = [torch.randn(B, D, generator=g).to(DEV) for _ in range(3)]
    for n in (0, 1, 3):
        qe, ae = q, a
        for i in range(n):
            qe = qe + rq[i][:, None]
            ae = ae + ra[i][:, None]
        oq, oa = KS.glimpse_residual_cast(q, rq[:n], a, ra[:n])
        assert torch.equal(oq.view(B, Tq, D), qe.to(torch.bfloat16)) and torch.equal(oa.view(B, Ta, D), ae.to(torch.bfloat16))
        js = KS.glimpse_token_sum(q, rq[:n], a, ra[:n])
        assert normrel(js, qe.sum(1) + ae.sum(1)) < 1e-6
        oq16, _ = KS.glimpse_residual_cast(q.to(torch.bfloat16), rq[:n], None, [])
        ref = q.to(torch.bfloat16).float()
        for i in range(n):
            ref = ref + rq[i][:, None]
        assert torch.equal(oq16.view(B, Tq, D), ref.to(torch.bfloat16))
    only_q = KS.glimpse_token_sum(q, [], None, [])
    assert normrel(only_q, q.sum(1)) < 1e-6
    x = torch.randn(B, D, generator=g).to(DEV)
    bq, ba = KS.glimpse_bcast_rows(x, Tq, Ta)
    assert torch.equal(bq, x[:, None].expand(B, Tq, D)) and torch.equal(ba, x[:, None].expand(B, Ta, D))


def test_pool_bwd_strided_attention_gradient():
    B, K, Q, A, C, G = 9, 50, 12, 6, 1024, 2
    g = torch.Generator().manual_seed(2)
    bf = torch.bfloat16
    v = torch.relu(torch.randn(B * K, C, generator=g)).to(bf).to(DEV)
    q = torch.relu(torch.randn(B * Q, C, generator=g)).to(bf).to(DEV)
    a = torch.relu(torch.randn(B * A, C, generator=g)).to(bf).to(DEV)
    w = torch.softmax(torch.randn(B, K * Q * A, generator=g), 1).view(B, K, Q, A).to(DEV)
    dout = torch.randn(B, C, generator=g).to(DEV)
    ref = KS.tri_pool_bwd(v, q, a, w, w.stride(0), dout, B, K, Q, A, C)
    buf = torch.full((B, G, K * Q * A), float("nan"), device=DEV)
    out = KS.tri_pool_bwd(v, q, a, w, w.stride(0), dout, B, K, Q, A, C, dw_out=buf[:, 1])
    assert torch.equal(buf[:, 1].reshape(B, K, Q, A), ref[6])
    assert torch.isnan(buf[:, 0]).all()
    for x, y in zip(out[:3], ref[:3]):
        assert torch.equal(x, y)
